"""python scripts/prof_subset.py make  -> writes the fast-only and full-algebra subsets of the 20K GA mix (classified at 4096
points by the build in use) to gpurun_out/subset_{real,full}.txt;  python scripts/prof_subset.py time n1 n2 .. -> times both."""
import sys
import numpy as np
sys.path.insert(0, ".")
import closyr_b200 as cb
ml = 40
def pts(n):
    xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
    return xs, 2.0 + 1.5 * xs / np.pi + np.cos(xs)
ctx = cb.Context([0])
if sys.argv[1] == "make":
    b = cb.PopulationBuilder(); texts = b.generate(888, 20000, ml, 0, want_texts=True)
    ctx.set_points(*pts(4096)); s, f = ctx.score_batch(b.view(), ml)
    for name, sel in (("real", (f & 4) == 0), ("full", (f & 4) != 0)):
        keep = [t for t, k in zip(texts, sel) if k]
        open("gpurun_out/subset_%s.txt" % name, "w").write("\n".join((keep * (20000 // len(keep) + 1))[:20000]))
else:
    for name in ("real", "full"):
        texts = open("gpurun_out/subset_%s.txt" % name).read().split("\n")
        b = cb.PopulationBuilder(); b.add_many(texts)
        for n in [int(a) for a in sys.argv[2:]]:
            ctx.set_points(*pts(n)); ctx.upload_population(b.view())
            for _ in range(3): ms = ctx.score_resident(ml)
            print("%-5s subset %d x %6d: %8.3f ms" % (name, len(texts), n, ms), flush=True)

"""ms per scoring call of the seeded GA-mix population at several point counts: python scripts/prof_n.py pop n1 n2 ..."""
import sys
import numpy as np
sys.path.insert(0, ".")
import closyr_b200 as cb

pop = int(sys.argv[1])
ml = 40
b = cb.PopulationBuilder()
b.generate(888, pop, ml, 0)
ctx = cb.Context([0])
ref = None
for n in [int(a) for a in sys.argv[2:]]:
    xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
    ys = 2.0 + 1.5 * xs / np.pi + np.cos(xs)
    ctx.set_points(xs, ys)
    ctx.upload_population(b.view())
    for _ in range(4):
        ms = ctx.score_resident(ml)
    s, f = ctx.download_scores()
    print("%7d x %7d: %8.3f ms   checksum %r" % (pop, n, ms, float(np.sum(np.asarray(s)[np.isfinite(s)]))), flush=True)

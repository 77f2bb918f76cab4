"""One scoring call of a seeded population, for ncu captures:
   python scripts/prof_one.py [pop] [points] [max_leafs] [mix] [all|real|full]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import closyr_b200 as cb

pop = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ml = int(sys.argv[3]) if len(sys.argv) > 3 else 40
mix = int(sys.argv[4]) if len(sys.argv) > 4 else 0
subset = sys.argv[5] if len(sys.argv) > 5 else "all"
xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
ys = 2.0 + 1.5 * xs / np.pi + np.cos(xs)
b = cb.PopulationBuilder()
texts = b.generate(888, pop, ml, mix, want_texts=True)
ctx = cb.Context([0])
ctx.set_points(xs, ys)
if subset != "all":
    s, f = ctx.score_batch(b.view(), ml)
    keep = [t for t, fl in zip(texts, f) if bool(fl & 4) == (subset == "full")]
    texts = (keep * (pop // max(1, len(keep)) + 1))[:pop]
    b = cb.PopulationBuilder()
    b.add_many(texts)
ctx.upload_population(b.view())
for _ in range(3):
    ms = ctx.score_resident(ml)
s, f = ctx.download_scores()
print("subset %s pop %d points %d: %.3f ms, full-semantics fraction %.3f" % (subset, pop, n, ms, ((f & 4) != 0).mean()))

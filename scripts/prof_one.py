"""One scoring call of a seeded population, for ncu captures: python scripts/prof_one.py [pop] [points] [max_leafs] [mix]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import closyr_b200 as cb

pop = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ml = int(sys.argv[3]) if len(sys.argv) > 3 else 40
mix = int(sys.argv[4]) if len(sys.argv) > 4 else 0
xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
ys = 2.0 + 1.5 * xs / np.pi + np.cos(xs)
b = cb.PopulationBuilder()
b.generate(888, pop, ml, mix)
ctx = cb.Context([0])
ctx.set_points(xs, ys)
ctx.upload_population(b.view())
for _ in range(3):
    ms = ctx.score_resident(ml)
s, f = ctx.download_scores()
print("pop %d points %d: %.3f ms, full-semantics fraction %.3f, min-score %.3f" % (pop, n, ms, ((f & 4) != 0).mean(), ((f & 2) != 0).mean()))

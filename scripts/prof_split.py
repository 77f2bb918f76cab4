"""Times the fast-path-only and the full-algebra subsets of a seeded population separately."""
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
texts = b.generate(888, pop, ml, mix, want_texts=True)
ctx = cb.Context([0])
ctx.set_points(xs, ys)

def timeit(texts, label):
    bb = cb.PopulationBuilder()
    bb.add_many(texts)
    ctx.upload_population(bb.view())
    for _ in range(3):
        ms = ctx.score_resident(ml)
    s, f = ctx.download_scores()
    print("%-28s %7d expr: %8.3f ms  %.3e expr-pts/s" % (label, len(texts), ms, len(texts) * n / ms * 1e3))
    return s, f

s, f = timeit(texts, "whole population")
real = [t for t, fl in zip(texts, f) if not (fl & 4)]
full = [t for t, fl in zip(texts, f) if (fl & 4)]
k = max(1, pop // max(1, len(real)))
timeit((real * (k + 1))[:pop], "fast-path-only subset")
k = max(1, pop // max(1, len(full)))
timeit((full * (k + 1))[:pop], "needs-full-algebra subset")

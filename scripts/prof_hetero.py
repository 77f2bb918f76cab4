"""What does heterogeneity cost?  K sampled expressions: sum of the times of K homogeneous populations (R copies of one
expression each) against ONE population holding the same K x R programs shuffled."""
import sys
import numpy as np
sys.path.insert(0, ".")
import closyr_b200 as cb

K = int(sys.argv[1]) if len(sys.argv) > 1 else 64
R = int(sys.argv[2]) if len(sys.argv) > 2 else 2960
ml = 40
b = cb.PopulationBuilder()
texts = b.generate(888, 100000, ml, 0, want_texts=True)
rng = np.random.default_rng(5)
pick = [texts[i] for i in rng.choice(len(texts), K, replace=False)]
ctx = cb.Context([0])

def timeit(tx):
    bb = cb.PopulationBuilder()
    bb.add_many(tx)
    ctx.upload_population(bb.view())
    for _ in range(3):
        ms = ctx.score_resident(ml)
    return ms

for n in (256, 1024, 4096, 16384):
    xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
    ys = 2.0 + 1.5 * xs / np.pi + np.cos(xs)
    ctx.set_points(xs, ys)
    each = np.array([timeit([t] * R) for t in pick])
    mix = [t for t in pick for _ in range(R)]
    o = rng.permutation(len(mix))
    t_mix = timeit([mix[i] for i in o])
    t_blk = timeit(mix)
    print("n=%6d  K=%d R=%d  sum of homogeneous %.3f ms (min %.3f med %.3f max %.3f per population)  shuffled mix %.3f ms  blocked mix %.3f ms"
          % (n, K, R, each.sum(), each.min(), np.median(each), each.max(), t_mix, t_blk), flush=True)

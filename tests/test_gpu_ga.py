"""The stand-alone GA loop (row N4) with the product scorer: every generation's children go through
closyr_score_batch on the GPU; the scores the GA works with must be the reference formula's (checked with the
oracle on the final population), for BASELINE configs[0] (README run) and a configs[1]-shaped run."""
import numpy as np
import pytest

import closyr_b200 as cb
from closyr_b200 import ga as G
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def synth_points(n):
    xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
    return xs, 2.0 + 1.5 * xs / np.pi + np.cos(xs)


def _check_population_scores(ga, xs, ys, max_leafs):
    exprs, scores = zip(*[ga.expr(i) for i in range(len(ga))])
    assert all(s is not None for s in scores)
    want = O.score_batch_conditioned(list(exprs), xs, ys, max_leafs, O.SUM_CUDA_TREE)
    bad = []
    for e, got, (w, well) in zip(exprs, scores, zip(*want)):
        if got == w or abs(got - w) <= 1e-9 * abs(w):
            continue
        if not well or O.score_discrepancy_explained(e, xs, ys, max_leafs, got, w, O.SUM_CUDA_TREE):
            continue
        bad.append((e, got, w))
    assert not bad, bad[:3]


def test_readme_run_on_gpu():       # configs[0]: pop 25, 5 iterations, xs 0..6, ys 1,2,30,4,5,6,10 (README.md:55)
    xs = np.arange(7, dtype=np.float64)
    ys = np.array([1, 2, 30, 4, 5, 6, 10], dtype=np.float64)
    with cb.Context([0]) as ctx:
        ctx.set_points(xs, ys)
        before = ctx.kernel_launches
        ga = G.GA(G.gpu_scorer(ctx, 40), 25, 40, seed=888, threads=1)
        best = []
        done = ga.run(5, on_iteration=lambda i, st, g: best.append(g.best()[0]))
        assert ctx.kernel_launches > before          # the scores came from the kernels
    assert done == 5 and len(ga) == 25
    assert all(b1 >= b0 for b0, b1 in zip(best, best[1:]))
    assert best[-1] > -38.57143242857143
    _check_population_scores(ga, xs, ys, 40)


def test_config2_shaped_run_on_gpu():   # configs[1]: pop 1000, max-leafs 40, 1024 points (20 of the 200 iterations)
    xs, ys = synth_points(1024)
    with cb.Context([0]) as ctx:
        ctx.set_points(xs, ys)
        ga = G.GA(G.gpu_scorer(ctx, 40), 1000, 40, seed=888, threads=0)
        best = []
        ga.run(20, on_iteration=lambda i, st, g: best.append(g.best()[0]))
    assert all(b1 >= b0 for b0, b1 in zip(best, best[1:]))
    assert best[-1] > best[0]
    _check_population_scores(ga, xs, ys, 40)

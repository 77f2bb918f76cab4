"""Sharded scoring on the GPU: the point-sharded building blocks on one device (ranks emulated one after the
other, as the profiling guide asks), and the in-process multi-device context when the box has several GPUs.
Scores must be BIT-IDENTICAL to single-device scoring for every split (fixed reduction tree)."""
import numpy as np
import pytest

import closyr_b200 as cb
from closyr_b200 import _lib as L
from closyr_b200 import sharder

pytestmark = pytest.mark.gpu


def synth_points(n):
    xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
    return xs, 2.0 + 1.5 * xs / np.pi + np.cos(xs)


def population(n_expr, seed=21):
    b = cb.PopulationBuilder()
    b.generate(seed, n_expr, 40, 0)
    b.add_many(["Plus(Rational(-1,2),E)", "Indeterminate", "Power(Plus(-1,x),Rational(1,2))", "x"])
    return b


@pytest.mark.parametrize("n_points,world", [(70000, 2), (70000, 4), (70000, 8), (5000, 8), (100, 4)])
def test_point_sharded_partials_combine_to_identical_scores(n_points, world):
    import torch
    xs, ys = synth_points(n_points)
    b = population(300)
    with cb.Context([0]) as ctx:
        ctx.set_points(xs, ys)
        want_s, want_f = ctx.score_batch(b.view(), 40)
        ne = ctx.n_expr
        plan = sharder.plan_point_shards(n_points, world)
        allp = torch.zeros(world * ne * 3, dtype=torch.float64, device="cuda:0")
        torch.cuda.synchronize()
        ctx.upload_population(b.view())
        parts = 0
        for r, (lo, hi) in enumerate(plan):                     # one emulated rank after the other
            ctx.score_partials(40, lo, hi, allp.data_ptr() + r * ne * 3 * 8)
            parts += 1 if hi > lo else 0
        ctx.finalize_partials(allp.data_ptr(), max(1, parts), 40)
        got_s, got_f = ctx.download_scores()
    assert got_s.tolist() == want_s.tolist()
    assert ((got_f ^ want_f) & ~np.uint8(L.R_FULL_SEMANTICS)).max() == 0


def test_single_rank_point_sharded_helper_matches():
    import torch
    xs, ys = synth_points(9000)
    b = population(200)
    with cb.Context([0]) as ctx:
        ctx.set_points(xs, ys)
        want, _ = ctx.score_batch(b.view(), 40)
        ctx.upload_population(b.view())
        got, _ = sharder.score_point_sharded(ctx, 40, 9000, 0, 1, None, torch.device("cuda:0"))
    assert got.tolist() == want.tolist()


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("shard", [L.SHARD_POPULATION, L.SHARD_POINTS])
def test_in_process_multi_device_context(shard):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    xs, ys = synth_points(20000)
    b = population(5000)
    with cb.Context([0]) as one:
        one.set_points(xs, ys)
        want, wf = one.score_batch(b.view(), 40)
    with cb.Context(list(range(n)), shard=shard) as many:
        assert many.device_count == n
        many.set_points(xs, ys)
        got, gf = many.score_batch(b.view(), 40)
        order = many.rank(50)
    assert got.tolist() == want.tolist()
    assert order.tolist() == np.argsort(want, kind="stable")[::-1][:50].tolist()


def test_multi_process_point_sharding():
    """One process per GPU (torchrun): the NCCL all-gather of the partial sums is the library's own."""
    import os
    import subprocess
    import sys
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mp_point_shard_worker.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", "29731", worker],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "MP_POINT_SHARD OK" in out.stdout

"""Multi-GPU host logic on CPU: shard plans, the fixed-tree combine model, and a world_size-2 gloo run of the
population-sharded and the point-sharded flows (the device scoring step is played by the oracle here)."""
import os
import socket

import numpy as np
import pytest

import closyr_b200 as cb
from closyr_b200 import sharder
from oracle import oracle as O


def synth_points(n):
    xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
    return xs, 2.0 + 1.5 * xs / np.pi + np.cos(xs)


def test_population_plan_covers_every_expression_once_and_balances_code():
    b = cb.PopulationBuilder()
    b.generate(888, 5000, 40, 0)
    off = b.arrays()["code_off"]
    for world in (1, 2, 3, 4, 8):
        cuts = sharder.plan_population_shards(off, world)
        assert cuts[0] == 0 and cuts[-1] == 5000 and len(cuts) == world + 1
        assert all(a <= c for a, c in zip(cuts, cuts[1:]))
        work = [int(off[c1] - off[c0]) for c0, c1 in zip(cuts, cuts[1:])]
        assert max(work) - min(work) <= 2 * int(np.diff(off).max())
    assert sharder.plan_population_shards(np.zeros(1, np.uint32), 4) == [0, 0, 0, 0, 0]


@pytest.mark.parametrize("n_points", [1, 127, 128, 129, 4096, 65536, 100000])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_point_plan_is_aligned_power_of_two_spans(n_points, world):
    plan = sharder.plan_point_shards(n_points, world)
    nch = (n_points + 127) // 128
    assert len(plan) == world
    covered = 0
    span = None
    for lo, hi in plan:
        assert lo == covered or lo == nch
        covered = max(covered, hi)
        if hi > lo:
            span = span or (plan[0][1] - plan[0][0] if plan[0][1] - plan[0][0] & (plan[0][1] - plan[0][0] - 1) == 0 else None)
    assert covered == nch
    width = max(hi - lo for lo, hi in plan)
    p2 = 1
    while p2 < width:
        p2 <<= 1
    for r, (lo, hi) in enumerate(plan):
        if hi > lo:
            assert lo == r * p2                    # aligned: every rank's span is a subtree of the chunk tree


def tree_partial(res, c0, c1):
    """(sum, max) of chunks [c0, c1) in the device's reduction shape."""
    return O.tree_partial(res, c0, c1)


def test_combine_reference_reproduces_the_single_device_tree():
    rng = np.random.default_rng(5)
    n = 5000
    res = rng.uniform(0, 8, n)
    whole, _ = tree_partial(res, 0, (n + 127) // 128)
    for world in (2, 4, 8):
        plan = sharder.plan_point_shards(n, world)
        parts = [[list(tree_partial(res, lo, hi)) + [0.0]] for lo, hi in plan if hi > lo]
        total, mx, nr = sharder.combine_partials_reference(np.array(parts))
        assert total[0] == whole and mx[0] == res.max() and nr[0] == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, texts, n_points, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xs, ys = synth_points(n_points)
    b = cb.PopulationBuilder()
    b.add_many(texts)
    off = b.arrays()["code_off"]
    # population sharding: every rank scores its block, all ranks gather the generation's scores
    cuts = sharder.plan_population_shards(off, world)
    view = sharder.shard_view(b.view(), cuts[rank], cuts[rank + 1])
    assert view.n_expr == cuts[rank + 1] - cuts[rank]
    local, _ = O.score_batch(texts[cuts[rank]:cuts[rank + 1]], xs, ys, 40)
    full = sharder.gather_scores(local, cuts, rank, world, dist)
    # point sharding: partial (sum, max, non-real) triples per rank, all-gather, fixed-tree combine
    import torch
    plan = sharder.plan_point_shards(n_points, world)
    lo, hi = plan[rank]
    mine = np.zeros((len(texts), 3))
    for i, t in enumerate(texts[:8]):
        f = O.eval_vec_pheno(t, xs)
        res = np.array([O.compute_residual(y, a) for y, a in zip(ys, f)])
        mine[i, 0], mine[i, 1] = tree_partial(res, lo, hi)
    gathered = [torch.zeros(len(texts), 3, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(mine))
    total, mx, _ = sharder.combine_partials_reference(np.stack([g.numpy() for g in gathered]))
    if rank == 0:
        out.put((full, total[:8], mx[:8]))
    dist.destroy_process_group()


def test_world_size_2_gloo_population_and_point_sharding():
    import torch.multiprocessing as mp
    n_points = 1000
    b = cb.PopulationBuilder()
    texts = b.generate(11, 64, 40, 0, want_texts=True)
    texts[:8] = ["x", "Plus(Rational(-1,2),Power(x,2))", "Times(x,x,x)", "Plus(1,Times(Rational(1,2),x))",
                 "Power(x,-1)", "Times(2,x)", "Plus(x,Times(-1,Power(x,2)))", "Power(Plus(1,x),Rational(1,2))"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, texts, n_points, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, total, mx = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    xs, ys = synth_points(n_points)
    want, _ = O.score_batch(texts, xs, ys, 40)
    assert full.tolist() == want.tolist()
    # the combined point-sharded partials give the same score as the single-device tree order, bit for bit
    for i, t in enumerate(texts[:8]):
        s = -1.0 * ((2.0 * (total[i] / n_points)) + mx[i])
        leafs = O.leaf_count(t)
        pen = min(0.1, (0.0000001 * leafs) * leafs)
        assert s - abs(s) * pen == O.score_fn(t, xs, ys, 40, O.SUM_CUDA_TREE)

"""Worker of tests/test_gpu_multi.py::test_multi_process_point_sharding (one process per GPU, launched by
torch.distributed.run): every rank holds the whole population, joins the library's NCCL communicator and scores with
the points split across the ranks; rank 0 compares with single-device scoring -- bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import closyr_b200 as cb  # noqa: E402
from closyr_b200 import _lib as L  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 70000
    xs = (np.pi / 15.0) * 49.0 * np.arange(n) / (n - 1)
    ys = 2.0 + 1.5 * xs / np.pi + np.cos(xs)
    b = cb.PopulationBuilder()
    b.generate(21, 3000, 40, 0)
    b.add_many(["Plus(Rational(-1,2),E)", "Indeterminate", "Power(Plus(-1,x),Rational(1,2))", "x"])
    with cb.Context([local], shard=L.SHARD_POINTS) as ctx:
        ctx.init_point_sharding(rank, world, dist, torch.device("cuda", local))
        ctx.set_points(xs, ys)
        got, gf = ctx.score_batch(b.view(), 40)               # collective
        order = ctx.rank(50)
    ok = True
    if rank == 0:
        with cb.Context([local]) as one:
            one.set_points(xs, ys)
            want, wf = one.score_batch(b.view(), 40)
        ok = got.tolist() == want.tolist() and order.tolist() == np.argsort(want, kind="stable")[::-1][:50].tolist()
        print("MP_POINT_SHARD %s world %d" % ("OK" if ok else "MISMATCH", world))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

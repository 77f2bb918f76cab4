"""End-to-end GA runs without a JVM: host GA loop (closyr_b200/csrc/ga.cpp)
around the GPU scorer.  Prints one JSON line per BASELINE config it is asked for.

    python scripts/ga_run.py c1        # README run: pop 25, 5 iterations, 7 points
    python scripts/ga_run.py c2        # pop 1000, 200 iterations, max-leafs 40, 1024 points
    python scripts/ga_run.py c3 [gens] # pop 100K, max-leafs 40, 4096 points, 50 generations

ms_per_generation is wall clock of one `solver-step`: shuffle + compete (mutation / crossover on the host) +
compile of the children + closyr_score_batch (H2D, kernels, D2H) + commit.  Early stop as in the reference
(symbolic_regression.clj:409-414) is disabled with --no-early-stop so the iteration count is the config's.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import closyr_b200 as cb          # noqa: E402
from closyr_b200 import ga as G   # noqa: E402


def synth_points(n):
    xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
    return xs, 2.0 + 1.5 * xs / np.pi + np.cos(xs)


CONFIGS = {
    "c1": dict(pop=25, iters=5, max_leafs=40, points="readme"),
    "c2": dict(pop=1000, iters=200, max_leafs=40, points=1024),
    "c3": dict(pop=100000, iters=50, max_leafs=40, points=4096),
}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    cfg = dict(CONFIGS[name])
    if len(sys.argv) > 2 and sys.argv[2].isdigit():
        cfg["iters"] = int(sys.argv[2])
    early = "--no-early-stop" not in sys.argv
    if cfg["points"] == "readme":
        xs = np.arange(7, dtype=np.float64)
        ys = np.array([1, 2, 30, 4, 5, 6, 10], dtype=np.float64)
    else:
        xs, ys = synth_points(cfg["points"])
    with cb.Context([0]) as ctx:
        ctx.set_points(xs, ys)
        ga = G.GA(G.gpu_scorer(ctx, cfg["max_leafs"]), cfg["pop"], cfg["max_leafs"], seed=888, threads=0)
        ga.score_pending()                              # generation 0: the whole population
        launches0 = ctx.kernel_launches
        per_gen = []
        t_all = time.perf_counter()
        done = 0
        for _ in range(cfg["iters"]):
            t0 = time.perf_counter()
            st = ga.evolve()
            ga.score_pending()
            per_gen.append((time.perf_counter() - t0) * 1e3)
            done += 1
            if early and st["early_stop"]:
                break
        total_ms = (time.perf_counter() - t_all) * 1e3
        launches = ctx.kernel_launches - launches0
        score, leafs, expr = ga.best()
    t = ga.timing
    print(json.dumps({
        "config": name, "population": cfg["pop"], "points": len(xs), "max_leafs": cfg["max_leafs"],
        "iterations_done": done, "iterations_asked": cfg["iters"],
        "ms_per_generation": total_ms / max(1, done), "ms_per_generation_median": float(np.median(per_gen)),
        "host_modify_ms_per_generation": t["modify_ms"] / max(1, done),
        "host_compile_ms_per_generation": t["compile_ms"] / max(1, done),
        "score_call_ms_per_generation": t["score_ms"] / max(1, done),
        "children_scored_per_generation": (t["scored"] - cfg["pop"]) / max(1, done),
        "gpu_launches": launches, "best_score": score, "best_leafs": leafs, "best_expr": cb.to_fullform(expr)[:200],
        "objective": "2 + 1.5*x/Pi + Cos(x)" if cfg["points"] != "readme" else "README points",
    }))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- measures the scoring hot path on B200 (contract in the task statement).

A step = scoring ONE generation (every expression of the population against every point).
Default workload = BASELINE.json configs[2] ("population 100K, max-leafs 40, 4096 points, population-sharded
at 1/2/4/8 B200"), the configuration the metric "expr-point evals/sec and ms/generation at 1/2/4/8 B200" is
quoted on; the population is split across the GPUs (strong scaling of a fixed generation).

    python bench.py --gpus N --steps K --warmup W                    # this framework
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU restatement on the host cores

Launch shapes:
  * torchrun (WORLD_SIZE set): one process per GPU; population sharding needs no collective, point sharding
    (--shard points) uses the library's own NCCL all-gather (Context.init_point_sharding);
  * plain `python bench.py --gpus N` with N > 1: ONE process drives N devices through one context -- the shape a
    JVM caller has (src/closyr/symbolic_regression.clj:474-478 builds one score closure per run).
On 8 GPUs the line also carries `c4_8gpu`: BASELINE.json configs[3] (1M expressions x 65536 points, max-leafs 80,
transcendental-heavy), population- and point-sharded, 3 generations each.

Prints one JSON line (rank 0).
"""
import argparse
import json
import lzma
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (population, points, max_leafs, mix, description)
    "c2": (1000, 1024, 40, 0, "configs[1]: pop 1000 x 1024 points, max-leafs 40"),
    "c3": (100000, 4096, 40, 0, "configs[2]: pop 100K x 4096 points, max-leafs 40, GA mix"),
    "c4": (1000000, 65536, 80, 1, "configs[3]: pop 1M x 65536 points, max-leafs 80, transcendental-heavy"),
}
SEED = 888   # echoes the reference's PRNG seed (src/closyr/util/prng.clj:59)

# FP64-pipe instructions on the fast path of CUDA 12.9 libdevice's own routine for each operation, read from
# sm_100a SASS (profiles/r01_libdevice_fp64_weights.md).  FROZEN: this table defines the algorithmic work of
# an (expression, point) pair; it does not change when the kernels do.
W_EPILOGUE = 4            # y - f, min with max-resid, sum, max
OP_WEIGHT = {
    5: 1, 6: 1, 7: 1, 8: 1,        # ADDX ADDC MULX MULC
    10: 5, 11: 1, 12: 8, 13: 13,   # RECIP SQR SQRT RSQRT
    15: 90, 16: 15, 17: 29,        # POWC EXP LOG
    18: 17, 19: 17, 20: 28,        # SIN COS TAN
    21: 23, 22: 24,                # ASIN ACOS
    25: 1, 26: 1, 27: 90,          # ADDS MULS POWS
    32: 1, 33: 1, 34: 1, 35: 1, 36: 1, 37: 1, 38: 1, 39: 1,   # ADDS0..3 MULS0..3
}


def synth_points(n):
    """the reference's example span; ys = the integration-test objective."""
    xs = (np.pi / 15.0) * 49.0 * np.arange(n, dtype=np.float64) / max(1, n - 1)
    ys = 2.0 + 1.5 * xs / np.pi + np.cos(xs)
    return xs, ys


def config_of(workload, shard):
    """The `config` object of the JSON line: identical for both arms."""
    pop_n, n_points, max_leafs, mix, desc = WORKLOADS[workload]
    return {"workload": desc, "population": pop_n, "points": n_points, "max_leafs": max_leafs, "sharding": shard,
            "generator_seed": SEED, "l2": "flushed between timed iterations (256 MiB memset)",
            "step": "score one full generation"}


def algorithmic_work(arrays):
    """Per expression: (weighted FP64 slots per point, instructions per point)."""
    code = arrays["code"].astype(np.int64)
    op = code & 0xFF
    w = np.zeros(256, dtype=np.int64)
    for k, v in OP_WEIGHT.items():
        w[k] = v
    slots = w[op]
    powi = op == 14
    if powi.any():
        n = ((code[powi] >> 16) & 0xFFFF).astype(np.int64)
        n = np.where(n >= 32768, n - 65536, n)
        a = np.abs(n)
        bits = np.floor(np.log2(np.maximum(a, 1))).astype(np.int64)
        pop = np.array([bin(int(v)).count("1") for v in a], dtype=np.int64)
        slots[powi] = bits + pop - 1 + np.where(n < 0, 5, 0)
    # node evaluations: one per instruction plus one per fused leaf load / slot store (ISA v1.1 modifiers), so
    # the count means the same as it did for v1 programs
    a = (code >> 8) & 0xFF
    mod = ((op == 1) | (op == 2) | ((op >= 9) & (op <= 23)))
    insn = (op != 0).astype(np.int64) + (mod & ((a & 2) != 0)) + (mod & (op >= 9) & ((a & 1) != 0))
    off = arrays["code_off"].astype(np.int64)
    starts = off[:-1]
    per_slots = np.add.reduceat(slots, starts) if len(starts) else np.zeros(0, np.int64)
    per_insn = np.add.reduceat(insn, starts) if len(starts) else np.zeros(0, np.int64)
    return per_slots + W_EPILOGUE, per_insn


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------------------------------------
# the reference arm
# ---------------------------------------------------------------------------------------------------------
def population_texts(workload, count):
    """FullForm texts of the workload's seeded population WITHOUT loading the product library into this process:
    the committed cache (baseline/populations, written by scripts/dump_population.py) when it holds the workload,
    otherwise a child process runs the generator and this process only reads its file."""
    pop_n, n_points, max_leafs, mix, desc = WORKLOADS[workload]
    cache = os.path.join(ROOT, "baseline", "populations", "%s_seed%d.txt.xz" % (workload, SEED))
    if os.path.exists(cache):
        with lzma.open(cache, "rt") as f:
            texts = f.read().split("\n")
        texts = [t for t in texts if t]
        if len(texts) >= count:
            return texts[:count], "committed cache " + os.path.relpath(cache, ROOT)
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "pop.txt")
        subprocess.check_call([sys.executable, os.path.join(ROOT, "scripts", "dump_population.py"), workload, out, str(count)],
                              stdout=subprocess.DEVNULL)
        with open(out) as f:
            texts = [t for t in f.read().split("\n") if t]
    return texts, "generated by a child process (scripts/dump_population.py)"


def run_reference(args, emit):
    """The reference arm: the CPU restatement of closyr's scoring path (oracle/oracle.c) on every host core, the
    WHOLE population of the workload per step (a bounded sample only where a step would not fit the run).
    The real reference (Clojure + Symja on a JVM) cannot run in this image: no JVM, no jars, no network.
    Nothing of the product is imported here: the only native code this process loads is oracle/."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import oracle as O

    pop_n, n_points, max_leafs, mix, desc = WORKLOADS[args.workload]
    # one step of the full workload in CPU time: ~1.3e-8 s per expression-point and core
    sample = pop_n if pop_n * n_points <= 1.0e9 else min(pop_n, args.cpu_sample)
    xs, ys = synth_points(n_points)
    texts, source = population_texts(args.workload, sample)
    cores = O.max_threads()
    for _ in range(args.warmup):
        O.score_batch(texts[: max(1, sample // 20)], xs, ys, max_leafs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.score_batch(texts, xs, ys, max_leafs)
    dt = time.perf_counter() - t0
    value = sample * n_points * args.steps / dt
    whole = sample == pop_n
    sample_desc = ("all %d expressions x all %d points per step" % (pop_n, n_points)) if whole else \
        ("first %d of %d expressions x all %d points per step" % (sample, pop_n, n_points))
    line = {
        "impl": "reference", "metric": "expr-point evals/sec", "value": value, "unit": "expr-points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps if whole else 1e3 * dt / args.steps * (pop_n / sample),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args.workload, args.shard),
        "note": "CPU restatement of closyr's scorer (C, tree walk over complex values, one thread per host core), not "
                "Symja/JVM; population texts: " + source +
                ("" if whole else "; ms_per_step is extrapolated linearly from the sample to the full generation"),
        "cpu_baseline": {"value": value, "unit": "expr-points/s", "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": "expr-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------
# the product arm
# ---------------------------------------------------------------------------------------------------------
def parity_spot_check(O, texts, xs, ys, max_leafs, scores):
    """GPU scores of the first expressions against the oracle: 1e-9 relative wherever fp64 determines the score (the
    oracle agrees with itself when every libm result moves by +-1 ulp), how far the others are, and whether any of it
    could change what the GA does with the scores (compete's `>=`, src/closyr/ga.clj:43; the best expression)."""
    m = len(texts)
    ref, stable = O.score_batch_conditioned(texts, xs, ys, max_leafs)
    got = np.asarray(scores[:m], dtype=np.float64)
    err = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
    flips = (got == O.MIN_SCORE) != (ref == O.MIN_SCORE)
    unstable = ~stable

    def hist(mask):
        e = err[mask]
        return {"<=1e-9": int((e <= 1e-9).sum()), "<=1e-6": int(((e > 1e-9) & (e <= 1e-6)).sum()),
                "<=1e-3": int(((e > 1e-6) & (e <= 1e-3)).sum()), ">1e-3": int((e > 1e-3).sum())}

    # compete pairs: neighbours (the reference pairs a shuffled population; any fixed pairing is as good a sample)
    a, b = np.arange(0, m - 1, 2), np.arange(1, m, 2)
    pair_diff = int(((got[a] >= got[b]) != (ref[a] >= ref[b])).sum())
    dev = [int(i) for i in np.nonzero(stable & (err > 1e-9))[0][:50]]
    unexplained = sum(0 if O.score_discrepancy_explained(texts[i], xs, ys, max_leafs, float(got[i]), float(ref[i])) else 1
                      for i in dev)
    return {
        "expressions": int(m), "fp64_stable_fraction": float(stable.mean()),
        "stable_within_1e-9": float((err[stable] <= 1e-9).mean()) if stable.any() else None,
        "all_within_1e-9": float((err <= 1e-9).mean()),
        "best_index_equal": bool(int(np.argmax(got)) == int(np.argmax(ref))),
        "stable_deviations_unexplained": int(unexplained),
        "rel_err_histogram_stable": hist(stable), "rel_err_histogram_unstable": hist(unstable),
        "max_rel_err_stable": float(err[stable].max()) if stable.any() else None,
        "max_rel_err_unstable": float(err[unstable].max()) if unstable.any() else None,
        "min_score_sentinel_flips": int(flips.sum()),
        "compete_pairs": int(len(a)), "compete_pairs_with_different_outcome": pair_diff,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="closyr_b200", choices=["closyr_b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--shard", default="population", choices=["population", "points"])
    ap.add_argument("--cpu-sample", type=int, default=20000, help="expressions scored by the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="on 8 GPUs: skip the configs[3] measurement")
    ap.add_argument("--population", type=int, default=0, help="override the workload's population (configs[4] sweep)")
    ap.add_argument("--points", type=int, default=0, help="override the workload's point count (configs[4] sweep)")
    args = ap.parse_args()
    # ONE JSON line on stdout: native libraries print there too (NCCL announces its version when NCCL_DEBUG is set), so
    # file descriptor 1 points at stderr while the run lasts and the line goes to the real stdout at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    if args.population or args.points:
        p0, n0, ml, mx, d0 = WORKLOADS[args.workload]
        p1, n1 = args.population or p0, args.points or n0
        WORKLOADS[args.workload] = (p1, n1, ml, mx, "configs[4] sweep point (%s shape): pop %d x %d points, max-leafs %d"
                                    % (args.workload, p1, n1, ml))
    if args.impl == "reference":
        run_reference(args, emit)
        return

    import torch
    import closyr_b200 as cb
    from closyr_b200 import _lib as L
    from closyr_b200 import sharder

    rank, world, local = dist_env()
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # one process, several devices: the JVM's shape
    in_process = world == 1 and args.gpus > 1
    devices = list(range(args.gpus)) if in_process else [local]
    n_gpus = args.gpus if in_process else world
    dev = torch.device("cuda", devices[0])
    torch.cuda.set_device(dev)

    def barrier():
        for d in devices:
            torch.cuda.synchronize(torch.device("cuda", d))
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        if dist is None:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak_cache = {}

    def measure(workload, shard, steps, warmup, want_clocks):
        """One workload through one context: device-timed generations, end-to-end generations, roofline."""
        pop_n, n_points, max_leafs, mix, desc = WORKLOADS[workload]
        xs, ys = synth_points(n_points)
        builder = cb.PopulationBuilder(pinned=True)
        builder.generate(SEED, pop_n, max_leafs, mix)
        full = builder.view()
        arrays = builder.arrays()
        slots, insn = algorithmic_work(arrays)
        by_points = shard == "points" and n_gpus > 1
        if in_process or by_points:
            lo, hi = 0, pop_n                        # the context shards (in process) / every rank holds everything
        else:
            cuts = sharder.plan_population_shards(arrays["code_off"], world)
            lo, hi = cuts[rank], cuts[rank + 1]
        mine = sharder.shard_view(full, lo, hi)
        bytes_in = int(4 * (arrays["code_off"][hi] - arrays["code_off"][lo]) + 8 * (arrays["const_off"][hi] - arrays["const_off"][lo])
                       + (hi - lo) * 16 + 8)
        bytes_out = int((hi - lo) * 9)
        if in_process and by_points:
            bytes_in *= n_gpus                       # every device receives the whole population
        ctx = cb.Context(devices, shard=L.SHARD_POINTS if by_points else L.SHARD_POPULATION)
        if by_points and not in_process:
            ctx.init_point_sharding(rank, world, dist, dev)
        ctx.set_points(xs, ys)
        if "dfma" not in peak_cache:
            peak_cache["dfma"] = ctx.fp64_peak_probe()     # register-resident DFMA chains on this GPU, this run
        peak_dfma = peak_cache["dfma"]

        # ---- device-resident throughput ("value"): inputs already in HBM, CUDA events on the launching streams
        ctx.upload_population(mine)
        for _ in range(warmup):
            ctx.score_resident(max_leafs)
        sampler = ClockSampler(devices[0]) if want_clocks else None
        barrier()
        launches0 = ctx.kernel_launches
        if sampler:
            sampler.start()
        dev_ms = 0.0
        t_wall0 = time.perf_counter()
        for _ in range(steps):
            ctx.l2_flush()                           # L2 flushed between timed iterations (not inside the event pair)
            if by_points and dist is not None:
                barrier()                            # a collective step: every rank enters it together
            dev_ms += ctx.score_resident(max_leafs)
        barrier()
        t_wall = time.perf_counter() - t_wall0
        clocks = sampler.stop() if sampler else None
        launches = ctx.kernel_launches - launches0
        ms_per_step = max_over_ranks(dev_ms / steps)
        value = pop_n * n_points / (ms_per_step * 1e-3)

        # ---- end to end through the C ABI with host buffers: H2D of the generation + kernels + D2H of the scores
        scores = flags = None
        for _ in range(max(1, warmup // 2)):
            scores, flags = ctx.score_batch(mine, max_leafs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            scores, flags = ctx.score_batch(mine, max_leafs)
        barrier()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / steps)

        # ---- roofline of the dominant kernel (score_kernel): weighted FP64 slots of the whole generation / device time
        # vs the measured DFMA rate of all the GPUs
        total_slots = float(slots.sum()) * n_points
        total_insn = float(insn.sum()) * n_points
        achieved = total_slots / (ms_per_step * 1e-3)
        roof = {
            "bound": "fp64", "achieved": 2 * achieved / 1e12, "peak": 2 * peak_dfma * n_gpus / 1e12, "unit": "TFLOP/s",
            "frac": achieved / (peak_dfma * n_gpus), "traffic": None,
            "note": "non-tensor FP64 pipe: achieved = frozen libdevice-weighted FP64 instruction slots of the bytecode "
                    "x points / score_kernel time (CUDA events), x2 flop; peak = register-resident DFMA chain measured "
                    "in this run on this GPU x n_gpus (MEASURED_PEAKS.json has no FP64 entry); theoretical 148 SM x 64 "
                    "DFMA/clk x 1.965 GHz = 37.2 TFLOP/s per GPU",
            "node_evals_per_s": total_insn / (ms_per_step * 1e-3),
            "hbm_points_stream_GBps": (16.0 * n_points + bytes_in) / (ms_per_step * 1e-3) / 1e9,
        }
        try:
            # DRAM traffic of the dominant kernel from the committed ncu --set full capture of this command, per launch;
            # only reported for the workload it was captured on
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            wl = tr["workload"]
            if (wl["population"], wl["points"], wl["max_leafs"], wl["n_gpus"], wl["sharding"]) == (pop_n, n_points, max_leafs, n_gpus, shard):
                roof["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
                roof["traffic_unit"] = "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)"
                roof["algorithmic_bytes_per_launch"] = bytes_in + bytes_out + 16 * n_points
        except Exception:
            pass
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            roof["hbm_peak_GBps_measured"] = peaks.get("hbm_gbs")
        except Exception:
            pass
        # the host GA's next step on the scores: the best 100 expressions, ranked on the device (row N1)
        rank_ms = None
        if want_clocks:
            ctx.rank(100)
            t0 = time.perf_counter()
            for _ in range(3):
                ctx.rank(100)
            rank_ms = 1e3 * (time.perf_counter() - t0) / 3
        out = {
            "ms_per_step": ms_per_step, "value": value, "rank_top100_ms": rank_ms,
            "e2e": {"value": pop_n * n_points / e2e_s, "unit": "expr-points/s", "h2d_bytes_per_step": bytes_in,
                    "d2h_bytes_per_step": bytes_out, "ms_per_step": 1e3 * e2e_s},
            "gpu_launches": int(launches), "roofline": roof, "clocks": clocks,
            "wall_ms_per_step_incl_flush": 1e3 * t_wall / steps,
            "_scores": scores, "_lo": lo, "_hi": hi, "_points": (xs, ys), "_max_leafs": max_leafs, "_mix": mix, "_pop": pop_n,
        }
        ctx.close()
        return out

    def cpu_side(workload, m_out, n_check, sample):
        """cpu_baseline and parity_spot_check of a measured workload (rank 0)."""
        from oracle import oracle as O
        pop_n, n_points, max_leafs, mix, desc = WORKLOADS[workload]
        xs, ys = m_out["_points"]
        sample = min(pop_n, sample)
        b2 = cb.PopulationBuilder()
        texts = b2.generate(SEED, sample, max_leafs, mix, want_texts=True)
        res = {}
        if sample > 0:
            O.score_batch(texts[: max(1, sample // 20)], xs, ys, max_leafs)          # warm: threads, page faults
            t0 = time.perf_counter()
            want, cores = O.score_batch(texts, xs, ys, max_leafs)
            dt = time.perf_counter() - t0
            res["cpu_baseline"] = {"value": sample * n_points / dt, "unit": "expr-points/s", "cores": cores, "kind": "port",
                                   "sample": "first %d of %d expressions x all %d points, one warm pass" % (sample, pop_n, n_points)}
        if m_out["_lo"] == 0 and m_out["_scores"] is not None and n_check > 0:
            m = min(n_check, sample, m_out["_hi"])
            res["parity_spot_check"] = parity_spot_check(O, texts[:m], xs, ys, max_leafs, m_out["_scores"])
        return res

    shard = args.shard
    m = measure(args.workload, shard, args.steps, args.warmup, True)
    line = {
        "metric": "expr-point evals/sec", "value": m["value"], "unit": "expr-points/s", "n_gpus": n_gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args.workload, shard),
        "launch": "one process driving %d devices" % n_gpus if in_process else "one process per GPU (torchrun)" if world > 1 else "one process, one GPU",
        "ms_per_generation": m["ms_per_step"],
        "e2e": m["e2e"], "gpu_launches": m["gpu_launches"], "roofline": m["roofline"], "clocks": m["clocks"],
        "wall_ms_per_step_incl_flush": m["wall_ms_per_step_incl_flush"],
        "rank_top100_ms": m["rank_top100_ms"],
    }
    if rank == 0 and not args.no_cpu_baseline:
        line.update(cpu_side(args.workload, m, 2000, args.cpu_sample))

    # ---- configs[3] on the whole box: the configuration the 70 % target is quoted on
    if n_gpus == 8 and args.workload == "c3" and not args.no_c4 and not (args.population or args.points):
        c4 = {"config": config_of("c4", "population"), "steps": 3, "warmup": 1}
        for sh in ("population", "points"):
            mm = measure("c4", sh, 3, 1, False)
            c4[sh] = {"ms_per_generation": mm["ms_per_step"], "value": mm["value"], "frac": mm["roofline"]["frac"],
                      "e2e": mm["e2e"], "gpu_launches": mm["gpu_launches"]}
            if rank == 0 and not args.no_cpu_baseline and sh == "population":
                c4.update(cpu_side("c4", mm, 300, 300))
        line["c4_8gpu"] = c4

    if rank == 0:
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Multi-GPU sharder for one-process-per-GPU launches (north star (4)).

Expressions are independent, so the primary mode splits the POPULATION across ranks with no data-path
collective: every rank scores its contiguous block and the per-generation scores are gathered back to the
host GA.  For very large point counts the POINTS are split instead: every rank evaluates all expressions on
an index-aligned power-of-two span of 128-point chunks, the per-expression partial (sum, max, non-real count)
triples are exchanged with one NCCL all-gather over NVLink, and every rank combines them in the same fixed
binary tree, so a score is bit-identical for 1, 2, 4 or 8 GPUs (a plain all-reduce would not be: its
summation order depends on the ring).  The single-process path (one JVM driving 8 GPUs) lives inside the C
library (closyr_ctx_create with several devices); this module is the `torch.distributed` launcher's side.

Reference hook: the population that `ga/evolve` scores with `pmap` (src/closyr/ga.clj:58-88).
"""
import ctypes as C

import numpy as np

from . import _lib as L

CHUNK_POINTS = 128


def plan_population_shards(code_off, world):
    """Contiguous expression blocks balanced by program length (a proxy for work), identical to the policy
    of closyr_upload_population.  Returns world+1 cut indices."""
    code_off = np.asarray(code_off, dtype=np.int64)
    n = len(code_off) - 1
    cuts = [0]
    total = int(code_off[n] - code_off[0]) if n > 0 else 0
    for r in range(1, world):
        want = total * r // world
        e = int(np.searchsorted(code_off - code_off[0], want, side="left")) if n > 0 else 0
        cuts.append(max(cuts[-1], min(n, e)))
    cuts.append(n)
    return cuts


def plan_point_shards(n_points, world):
    """Aligned power-of-two spans of 128-point chunks, one per rank: [(chunk_begin, chunk_end)] * world.
    Trailing ranks may get an empty span; the reduction tree over ranks is the top of the chunk tree."""
    nch = (n_points + CHUNK_POINTS - 1) // CHUNK_POINTS
    per = (nch + world - 1) // world
    span = 1
    while span < max(1, per):
        span <<= 1
    out = []
    for r in range(world):
        lo = min(nch, r * span)
        out.append((lo, min(nch, lo + span)))
    return out


def shard_view(pop, lo, hi):
    """A closyr_population struct describing expressions [lo, hi) of `pop` (no copy; offsets stay absolute)."""
    v = L.Population()
    v.n_expr = hi - lo
    v.code = pop.code
    v.consts = pop.consts
    u32 = C.sizeof(C.c_uint32)
    for name in ("code_off", "const_off", "meta", "aux"):
        base = C.cast(getattr(pop, name), C.c_void_p).value or 0
        setattr(v, name, C.cast(C.c_void_p(base + lo * u32), C.POINTER(C.c_uint32)))
    return v


def gather_scores(local_scores, cuts, rank, world, dist=None, device=None):
    """All ranks receive the whole generation's scores (float64[n_expr]) in population order."""
    n = cuts[-1]
    if world == 1 or dist is None:
        return np.asarray(local_scores, dtype=np.float64).copy()
    import torch

    sizes = [cuts[r + 1] - cuts[r] for r in range(world)]
    width = max(sizes) if sizes else 0
    buf = torch.zeros(width, dtype=torch.float64, device=device)
    buf[: sizes[rank]] = torch.as_tensor(np.asarray(local_scores, dtype=np.float64), device=device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    full = np.empty(n, dtype=np.float64)
    for r in range(world):
        full[cuts[r]:cuts[r + 1]] = out[r][: sizes[r]].cpu().numpy()
    return full


def combine_partials_reference(parts):
    """Host model of the device combine (closyr_finalize_partials): parts is [n_parts][n_expr][3]; the sums are
    combined by an index-aligned binary tree over the parts, absent parts skipped.  Used by CPU tests."""
    parts = np.asarray(parts, dtype=np.float64)
    n_parts = parts.shape[0]
    lvl = {}
    for t in range(n_parts):
        s = parts[t, :, 0].copy()
        k, c = 0, t
        while c & 1:
            s = lvl[k] + s
            c >>= 1
            k += 1
        lvl[k] = s
    total = None
    k = 0
    while (n_parts >> k) != 0:
        if (n_parts >> k) & 1:
            total = lvl[k] if total is None else lvl[k] + total
        k += 1
    return total, parts[:, :, 1].max(axis=0), np.floor(parts[:, :, 2]).sum(axis=0)


def score_point_sharded(ctx, max_leafs, n_points, rank, world, dist, torch_device):
    """Point-sharded scoring of the resident population of `ctx` (every rank holds the whole population).
    world > 1: `ctx` was created with shard=SHARD_POINTS and joined with Context.init_point_sharding; the
    library computes this rank's partials, all-gathers them with NCCL and combines them (collective call).
    world == 1 (tests): the same building blocks on one device, partials -> fixed-tree combine.
    Returns (scores, flags) on every rank."""
    if world > 1:
        ctx.score_resident(max_leafs)
        return ctx.download_scores()
    import torch

    plan = plan_point_shards(n_points, world)
    ne = ctx.n_expr
    mine = torch.zeros(ne * 3, dtype=torch.float64, device=torch_device)
    lo, hi = plan[rank]
    torch.cuda.synchronize(torch_device)        # the library launches on its own (non-blocking) stream
    ctx.score_partials(max_leafs, lo, hi, mine.data_ptr())
    torch.cuda.synchronize(torch_device)
    ctx.finalize_partials(mine.data_ptr(), 1, max_leafs)
    return ctx.download_scores()

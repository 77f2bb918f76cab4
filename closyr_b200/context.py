"""Context: thin Python handle on closyr_ctx (streams, device buffers, resident xs/ys)."""
import ctypes as C

import numpy as np

from . import _lib as L


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Context:
    """One context drives 1..8 GPUs of one box from one process (the shape a JVM caller needs).

    devices: list of CUDA device indices (default [0]).  shard: L.SHARD_POPULATION or L.SHARD_POINTS.
    """

    def __init__(self, devices=None, shard=L.SHARD_POPULATION):
        devices = [0] if devices is None else list(devices)
        arr = (C.c_int * len(devices))(*devices)
        self._h = C.c_void_p()
        L.check(L.lib().closyr_ctx_create(arr, len(devices), shard, C.byref(self._h)))
        self.n_points = None
        self.n_expr = None

    def close(self):
        if self._h:
            L.lib().closyr_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def device_count(self):
        return L.lib().closyr_ctx_device_count(self._h)

    def set_points(self, xs, ys):
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        ys = np.ascontiguousarray(ys, dtype=np.float64)
        if xs.shape != ys.shape or xs.ndim != 1:
            raise ValueError("xs and ys must be 1-d and of equal length")
        L.check(L.lib().closyr_set_points(self._h, _dptr(xs), _dptr(ys), len(xs)), self._h)
        self.n_points = len(xs)

    def score_batch(self, pop, max_leafs=L.DEFAULT_MAX_LEAFS):
        """Host buffers in, host buffers out: (scores float64[n], flags uint8[n])."""
        n = pop.n_expr
        scores = np.empty(n, dtype=np.float64)
        flags = np.empty(n, dtype=np.uint8)
        L.check(L.lib().closyr_score_batch(self._h, C.byref(pop), max_leafs, _dptr(scores),
                                           flags.ctypes.data_as(C.POINTER(C.c_uint8))), self._h)
        self.n_expr = n
        return scores, flags

    def upload_population(self, pop):
        L.check(L.lib().closyr_upload_population(self._h, C.byref(pop)), self._h)
        self.n_expr = pop.n_expr

    def score_resident(self, max_leafs=L.DEFAULT_MAX_LEAFS):
        """Scores the resident population; returns the device time in ms (CUDA events on the ctx streams)."""
        ms = C.c_float()
        L.check(L.lib().closyr_score_resident(self._h, max_leafs, C.byref(ms)), self._h)
        return ms.value

    def download_scores(self, scores=None, flags=None):
        n = self.n_expr
        scores = np.empty(n, dtype=np.float64) if scores is None else scores
        flags = np.empty(n, dtype=np.uint8) if flags is None else flags
        L.check(L.lib().closyr_download_scores(self._h, _dptr(scores), flags.ctypes.data_as(C.POINTER(C.c_uint8))),
                self._h)
        return scores, flags

    def eval_batch(self, pop, xs):
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        n, ne = len(xs), pop.n_expr
        f = np.empty((ne, n), dtype=np.float64)
        status = np.empty(ne, dtype=np.uint8)
        L.check(L.lib().closyr_eval_batch(self._h, C.byref(pop), _dptr(xs), n, _dptr(f),
                                          status.ctypes.data_as(C.POINTER(C.c_uint8))), self._h)
        return f, status

    def score_from_values(self, f_of_xs, ys, leafs):
        f = np.ascontiguousarray(f_of_xs, dtype=np.float64)
        y = np.ascontiguousarray(ys, dtype=np.float64)
        out = np.empty(3, dtype=np.float64)
        L.check(L.lib().closyr_score_from_values(self._h, _dptr(f), _dptr(y), len(f), int(leafs), _dptr(out)), self._h)
        return float(out[0]), float(out[1]), float(out[2])

    def rank(self, k=None):
        k = self.n_expr if k is None else min(k, self.n_expr)
        out = np.empty(k, dtype=np.int64)
        L.check(L.lib().closyr_rank_resident(self._h, k, out.ctypes.data_as(C.POINTER(C.c_int64))), self._h)
        return out

    def score_partials(self, max_leafs, chunk_begin, chunk_end, dev_ptr):
        L.check(L.lib().closyr_score_partials_resident(self._h, max_leafs, chunk_begin, chunk_end, C.c_void_p(dev_ptr)),
                self._h)

    def finalize_partials(self, dev_ptr, n_parts, max_leafs):
        L.check(L.lib().closyr_finalize_partials(self._h, C.c_void_p(dev_ptr), n_parts, max_leafs), self._h)

    def init_point_sharding(self, rank, world, dist=None, torch_device=None):
        """One process per GPU: makes this context (ONE device, shard=SHARD_POINTS) rank `rank` of `world`.  Rank 0
        creates the NCCL id inside the library, `dist` (torch.distributed, any backend) carries its 128 bytes to the
        other ranks, every rank joins; afterwards score_resident / score_batch are collective calls and the
        all-gather of the partial sums is issued by the library on its own stream."""
        ident = (C.c_uint8 * L.COMM_ID_BYTES)()
        if rank == 0:
            L.check(L.lib().closyr_comm_unique_id(ident), self._h)
        if world > 1:
            import torch
            t = torch.tensor(list(ident), dtype=torch.uint8, device=torch_device if dist.get_backend() == "nccl" else "cpu")
            dist.broadcast(t, src=0)
            ident = (C.c_uint8 * L.COMM_ID_BYTES)(*t.cpu().tolist())
        L.check(L.lib().closyr_comm_init_rank(self._h, world, rank, ident), self._h)

    def l2_flush(self):
        L.check(L.lib().closyr_l2_flush(self._h), self._h)

    def fp64_peak_probe(self):
        out = C.c_double()
        L.check(L.lib().closyr_fp64_peak_probe(self._h, C.byref(out)), self._h)
        return out.value

    @property
    def kernel_launches(self):
        return L.lib().closyr_kernel_launch_count(self._h)

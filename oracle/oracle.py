"""ctypes loader for the CPU oracle (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (closyr_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

MIN_SCORE = -100000000.0   # ops.clj:24-26
MAX_RESID = 1000000.0      # ops.clj:34-36
SUM_LEFT_FOLD = 0          # the reference's (reduce + 0.0 coll), ops.clj:61-63
SUM_CUDA_TREE = 1          # the CUDA path's fixed reduction shape


def build(force=False):
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        L.oracle_leaf_count.argtypes = [C.c_char_p]
        L.oracle_leaf_count.restype = C.c_int
        L.oracle_eval.argtypes = [C.c_char_p, dp, C.c_int64, dp]
        L.oracle_eval.restype = C.c_int
        L.oracle_eval_raw.argtypes = [C.c_char_p, dp, C.c_int64, dp, dp, C.POINTER(C.c_int)]
        L.oracle_eval_raw.restype = C.c_int
        L.oracle_compute_residual.argtypes = [C.c_double, C.c_double]
        L.oracle_compute_residual.restype = C.c_double
        L.oracle_score_from_f.argtypes = [dp, dp, C.c_int64, C.c_int, C.c_int]
        L.oracle_score_from_f.restype = C.c_double
        L.oracle_score.argtypes = [C.c_char_p, dp, dp, C.c_int64, C.c_int, C.c_int]
        L.oracle_score.restype = C.c_double
        L.oracle_score_batch.argtypes = [C.POINTER(C.c_char_p), C.c_int64, dp, dp, C.c_int64, C.c_int, C.c_int,
                                         C.c_int, dp]
        L.oracle_score_batch.restype = C.c_int
        L.oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def leaf_count(expr):
    return lib().oracle_leaf_count(expr.encode())


def eval_vec_pheno(expr, xs):
    """eval.clj:99-112.  Returns the f(xs) vector, or None where the reference returns nil; raises
    ArithmeticError where the reference throws (constant expression that is not a real number)."""
    xs, pxs = _d(xs)
    out = np.empty(len(xs), dtype=np.float64)
    st = lib().oracle_eval(expr.encode(), pxs, len(xs), out.ctypes.data_as(C.POINTER(C.c_double)))
    if st == 0:
        return out
    if st == 1:
        return None
    if st == 2:
        raise ArithmeticError("constant expression is not a real number: " + expr)
    raise ValueError("cannot parse FullForm expression: " + expr)


def eval_raw(expr, xs):
    xs, pxs = _d(xs)
    n = len(xs)
    re = np.empty(n)
    im = np.empty(n)
    tags = np.empty(n, dtype=np.int32)
    st = lib().oracle_eval_raw(expr.encode(), pxs, n, re.ctypes.data_as(C.POINTER(C.c_double)),
                               im.ctypes.data_as(C.POINTER(C.c_double)), tags.ctypes.data_as(C.POINTER(C.c_int)))
    if st != 0:
        raise ValueError("cannot parse FullForm expression: " + expr)
    return re, im, tags


def compute_residual(expected, actual):
    return lib().oracle_compute_residual(float(expected), float(actual))


def score_from_f(f, ys, leafs, sum_mode=SUM_LEFT_FOLD):
    f, pf = _d(f)
    ys, pys = _d(ys)
    return lib().oracle_score_from_f(pf, pys, len(f), int(leafs), sum_mode)


def score_fn(expr, xs, ys, max_leafs=40, sum_mode=SUM_LEFT_FOLD):
    """ops.clj:119-136 on a FullForm string."""
    xs, pxs = _d(xs)
    ys, pys = _d(ys)
    return lib().oracle_score(expr.encode(), pxs, pys, len(xs), int(max_leafs), sum_mode)


def score_batch(exprs, xs, ys, max_leafs=40, sum_mode=SUM_LEFT_FOLD, nthreads=0):
    """Scores a population with one thread per host core (pmap-style chunks, ga.clj:49-55).
    Returns (scores, threads_used)."""
    xs, pxs = _d(xs)
    ys, pys = _d(ys)
    n = len(exprs)
    arr = (C.c_char_p * n)(*[e.encode() if isinstance(e, str) else e for e in exprs])
    out = np.empty(n, dtype=np.float64)
    used = lib().oracle_score_batch(arr, n, pxs, pys, len(xs), int(max_leafs), sum_mode, int(nthreads),
                                    out.ctypes.data_as(C.POINTER(C.c_double)))
    return out, used


def max_threads():
    return lib().oracle_max_threads()

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
        L.oracle_tree_partial.argtypes = [dp, C.c_int64, C.c_int64, C.c_int64, dp, dp]
        L.oracle_tree_partial.restype = None
        L.oracle_set_jitter.argtypes = [C.c_int]
        L.oracle_set_jitter.restype = None
        L.oracle_set_jitter_ulps.argtypes = [C.c_int]
        L.oracle_set_jitter_ulps.restype = None
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


def tree_partial(resid, chunk_begin, chunk_end):
    """(sum, max) of residual chunks [chunk_begin, chunk_end) in the CUDA reduction shape (aligned power-of-two span)."""
    r, pr = _d(resid)
    s, m = C.c_double(), C.c_double()
    lib().oracle_tree_partial(pr, len(r), chunk_begin, chunk_end, C.byref(s), C.byref(m))
    return s.value, m.value


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


# Envelope of the explanation probes: the math library under test may be up to 2 ulp off (CUDA's documented bound for
# sin cos tan asin acos pow; this repo's own routines state <= 2 ulp), the oracle's glibc about 0.5 ulp.
EXPLAIN_ULPS = 2


def score_batch_conditioned(exprs, xs, ys, max_leafs=40, sum_mode=SUM_LEFT_FOLD, tol=1e-10, probes=8):
    """Scores plus a mask of the expressions whose score is determined by fp64 arithmetic: the oracle is run
    again `probes` times with every transcendental result moved by a random +-1 ulp (other, equally valid
    math libraries); where its own scores move by more than `tol` relative, no implementation can be held
    to 1e-9 (chaotic expressions such as Cos(E^(x^2)), exact cancellations that decide a branch)."""
    base, _ = score_batch(exprs, xs, ys, max_leafs, sum_mode)
    ok = np.ones(len(base), dtype=bool)
    try:
        for seed in range(1, probes + 1):
            lib().oracle_set_jitter(seed)
            other, _ = score_batch(exprs, xs, ys, max_leafs, sum_mode)
            ok &= np.abs(other - base) <= tol * np.maximum(1.0, np.abs(base))
    finally:
        lib().oracle_set_jitter(0)
    return base, ok


def eval_conditioned(expr, xs, tol=1e-10, probes=16):
    """eval_vec_pheno plus a per-point mask of the values the oracle reproduces under libm jitter."""
    base = eval_vec_pheno(expr, xs)
    if base is None:
        return None, None
    ok = np.ones(len(base), dtype=bool)
    try:
        for seed in range(1, probes + 1):
            lib().oracle_set_jitter(seed)
            other = eval_vec_pheno(expr, xs)
            with np.errstate(invalid="ignore"):
                same = (other == base) | (np.abs(other - base) <= tol * np.abs(base))
            ok &= same
    finally:
        lib().oracle_set_jitter(0)
    return base, ok


def score_discrepancy_explained(expr, xs, ys, max_leafs, got, want, sum_mode=SUM_LEFT_FOLD, max_probes=256):
    """True when the oracle itself moves by at least half of |got - want| under some +-1 ulp perturbation of
    its math library: the expression sits on a knife edge (domain boundary, branch cut, chaotic argument)
    where two correct implementations legitimately differ.  A systematic error of the implementation under
    test is not reproducible this way, so it still fails."""
    need = 0.5 * abs(got - want)
    try:
        lib().oracle_set_jitter_ulps(EXPLAIN_ULPS)
        for seed in range(1, max_probes + 1):
            lib().oracle_set_jitter(seed)
            if abs(score_fn(expr, xs, ys, max_leafs, sum_mode) - want) >= need:
                return True
    finally:
        lib().oracle_set_jitter(0)
        lib().oracle_set_jitter_ulps(1)
    return False


def eval_discrepancy_explained(expr, xs, got, want, max_probes=256):
    """Per-point version of score_discrepancy_explained for eval vectors (inf = not a real number)."""
    got = np.asarray(got, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        bad = ~((got == want) | (np.abs(got - want) <= 1e-9 * np.abs(want)))
    todo = set(np.nonzero(bad)[0].tolist())
    try:
        lib().oracle_set_jitter_ulps(EXPLAIN_ULPS)
        for seed in range(1, max_probes + 1):
            if not todo:
                break
            lib().oracle_set_jitter(seed)
            other = eval_vec_pheno(expr, xs)
            for k in list(todo):
                if np.isinf(got[k]) != np.isinf(want[k]):
                    hit = np.isinf(other[k]) == np.isinf(got[k])
                else:
                    hit = np.isinf(other[k]) != np.isinf(want[k]) or abs(other[k] - want[k]) >= 0.5 * abs(got[k] - want[k])
                if hit:
                    todo.discard(k)
    finally:
        lib().oracle_set_jitter(0)
        lib().oracle_set_jitter_ulps(1)
    return not todo, int(bad.sum())

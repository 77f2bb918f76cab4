"""Stand-alone GA loop around the GPU scorer.

Mirrors the reference's solver loop so that "ms per generation" and "200 iterations" mean what they mean
there, without a JVM:

    ga/initialize + ga/evolve               src/closyr/ga.clj:11-88
    SolverStateController solver-step       src/closyr/symbolic_regression.clj:474-510 (next-iters: 409-414)
    ops/mutation-fn, ops/crossover-fn       src/closyr/ops.clj:143-188
    the 96 operators                        src/closyr/ops/initialize.clj:35-665

The expression rewrites and the evolve step live in the C library (closyr_b200/csrc/ga.cpp); this module wires
them to a scorer.  The scorer is any callable `(closyr_population, n) -> float64[n]`; the product scorer is
`Context.score_batch` (the sm_100a kernels).  There is no CPU scorer in this package: tests plug in the oracle.
"""
import ctypes as C
import time

import numpy as np

from . import _lib as L


def rand_ints(seed, n, count):
    """`(prng/set-random-seed! seed)` followed by `count` calls of `(prng/rand-int n)` (prng.clj:16-41)."""
    out = (C.c_int32 * count)()
    L.check(L.lib().closyr_jrandom_rand_ints(seed, n, count, out))
    return list(out)


def shuffled_indices(seed, n):
    """`(prng/shuffle (range n))` with a freshly seeded generator (prng.clj:49-55)."""
    out = (C.c_int32 * n)()
    L.check(L.lib().closyr_jrandom_shuffle(seed, n, out))
    return list(out)


def mutations():
    """[(op, label)] in the order of `ops-init/initial-mutations`."""
    lib = L.lib()
    out = []
    for i in range(lib.closyr_mutation_count()):
        op, label = C.c_char_p(), C.c_char_p()
        L.check(lib.closyr_mutation_info(i, C.byref(op), C.byref(label)))
        out.append((op.value.decode(), label.value.decode()))
    return out


def _take_string(p):
    s = C.cast(p, C.c_char_p).value.decode()
    L.lib().closyr_free(p)
    return s


def mutate_expr(expr, index, seed=None):
    """`(ops-modify/modify mutation pheno)` on expression text; seed None = every probabilistic test passes."""
    out = C.c_void_p()
    L.check(L.lib().closyr_mutate_expr(expr.encode(), index, 0 if seed is None else 1, seed or 0, C.byref(out)))
    return _take_string(out)


def crossover_expr(e1, e2, max_leafs=L.DEFAULT_MAX_LEAFS, seed=888):
    out = C.c_void_p()
    L.check(L.lib().closyr_crossover_expr(e1.encode(), e2.encode(), max_leafs, seed, C.byref(out)))
    return _take_string(out)


class GA:
    """The GA state of one run: `ga/initialize` (population of `x`) plus `ga/evolve`."""

    def __init__(self, scorer, pop_count, max_leafs=L.DEFAULT_MAX_LEAFS, seed=888, threads=0):
        self._h = C.c_void_p()
        L.check(L.lib().closyr_ga_create(pop_count, max_leafs, seed, threads, C.byref(self._h)))
        self.scorer = scorer
        self.max_leafs = max_leafs
        self.generation = 0
        self.timing = {"modify_ms": 0.0, "compile_ms": 0.0, "score_ms": 0.0, "scored": 0}

    def close(self):
        if self._h:
            L.lib().closyr_ga_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return L.lib().closyr_ga_population_count(self._h)

    def set_expr(self, index, expr):
        L.check(L.lib().closyr_ga_set_expr(self._h, index, expr.encode()))

    def expr(self, index):
        out, score, scored = C.c_void_p(), C.c_double(), C.c_int32()
        L.check(L.lib().closyr_ga_expr(self._h, index, C.byref(out), C.byref(score), C.byref(scored)))
        return _take_string(out), (score.value if scored.value else None)

    def score_pending(self):
        """`(pmap (partial with-score score-fn))` (ga.clj:25-29, 62-65) as one batch; returns the batch size."""
        t0 = time.perf_counter()
        view, n = L.Population(), C.c_int64()
        L.check(L.lib().closyr_ga_pending(self._h, C.byref(view), C.byref(n)))
        t1 = time.perf_counter()
        if n.value:
            scores = np.ascontiguousarray(self.scorer(view, n.value), dtype=np.float64)
            if scores.shape != (n.value,):
                raise ValueError("scorer returned %r scores for %d phenotypes" % (scores.shape, n.value))
            L.check(L.lib().closyr_ga_commit_scores(self._h, scores.ctypes.data_as(C.POINTER(C.c_double)), n.value))
        else:
            L.check(L.lib().closyr_ga_commit_scores(self._h, None, 0))
        t2 = time.perf_counter()
        self.timing["compile_ms"] += (t1 - t0) * 1e3
        self.timing["score_ms"] += (t2 - t1) * 1e3
        self.timing["scored"] += n.value
        return n.value

    def pending_exprs(self):
        """FullForm text of the phenotypes that have no score yet (for an external checker)."""
        view, n = L.Population(), C.c_int64()
        L.check(L.lib().closyr_ga_pending(self._h, C.byref(view), C.byref(n)))
        idx = (C.c_int64 * max(1, n.value))()
        L.check(L.lib().closyr_ga_pending_indices(self._h, idx, n.value))
        return [self.expr(idx[i])[0] for i in range(n.value)]

    def evolve(self):
        """One `ga/evolve` (ga.clj:58-88): scores whoever is unscored, then shuffle + compete; the children
        are scored by the next call (or by `score_pending`).  Returns the stats of the step."""
        self.score_pending()
        st = L.GAStats()
        L.check(L.lib().closyr_ga_evolve(self._h, C.byref(st)))
        self.timing["modify_ms"] += st.ms_modify
        self.generation += 1
        return {k: getattr(st, k) for k, _ in L.GAStats._fields_ if k != "reserved"}

    def best(self):
        """First element of `ops/sort-population` (ops.clj:191-201): (score, leafs, FullForm text)."""
        idx, score, leafs, out = C.c_int64(), C.c_double(), C.c_int32(), C.c_void_p()
        L.check(L.lib().closyr_ga_best(self._h, C.byref(idx), C.byref(score), C.byref(leafs), C.byref(out)))
        return score.value, leafs.value, _take_string(out)

    def run(self, iters, on_iteration=None):
        """`solver-step` until the iterations run out or a pair of phenotypes is a near-exact solution
        (`next-iters`, symbolic_regression.clj:409-414).  Returns the number of iterations done."""
        done = 0
        for _ in range(iters):
            st = self.evolve()
            self.score_pending()            # the reference scores each child inside compete (ga.clj:41-44)
            done += 1
            if on_iteration is not None:
                on_iteration(done, st, self)
            if st["early_stop"]:
                break
        return done


def gpu_scorer(ctx, max_leafs=L.DEFAULT_MAX_LEAFS):
    """The product scorer: one `closyr_score_batch` call per batch (host buffers in, scores out)."""
    def score(view, n):
        return ctx.score_batch(view, max_leafs)[0]
    return score

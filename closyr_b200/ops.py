"""Host-side mirror of the reference's operator interface for the scoring path.

Same names, argument meaning and error behaviour as closyr.ops / closyr.ops.eval / closyr.ops.common, so the
parity tests read like the reference's own tests; every function below runs on the GPU through the C ABI
(include/closyr_b200.h).  Nothing here evaluates expressions on the CPU.

  ->phenotype                               src/closyr/ops/common.clj:127-146   -> to_phenotype
  ->run-args                                src/closyr/symbolic_regression.clj:355-383 -> to_run_args
  ops/score-fn                              src/closyr/ops.clj:119-136          -> score_fn (+ score_batch)
  ops/compute-score-from-actuals-and-expecteds  src/closyr/ops.clj:98-116       -> compute_score_from_actuals_and_expecteds
  ops/compute-residual                      src/closyr/ops.clj:78-90            -> compute_residual
  ops-eval/eval-vec-pheno                   src/closyr/ops/eval.clj:99-112      -> eval_vec_pheno
  ops-eval/eval-vec-pheno-oversample        src/closyr/ops/eval.clj:115-175     -> eval_vec_pheno_oversample
  ops-common/extend-xs                      src/closyr/ops/common.clj:262-291   -> extend_xs
  sort-population                           src/closyr/ops.clj:195-201          -> sort_population
"""
import re
import uuid

import numpy as np

from . import _lib as L
from .context import Context
from .population import PopulationBuilder, to_fullform

min_score = L.MIN_SCORE
max_resid = L.MAX_RESID
default_max_leafs = L.DEFAULT_MAX_LEAFS

_BARE_SYMBOL = re.compile(r"^\s*[A-Za-z$][A-Za-z0-9$_]*\s*$")


def to_phenotype(expr, sym="x"):
    """`->phenotype`: NIL, or a bare built-in symbol that is not a real number (Indeterminate, Infinity, E,
    Sin ...) is replaced by x*1 = x (common.clj:127-132); everything else is canonicalised."""
    if expr is None or not str(expr).strip() or (_BARE_SYMBOL.match(str(expr)) and str(expr).strip() != sym):
        expr = sym
    return {"sym": sym, "id": str(uuid.uuid4()), "expr": to_fullform(str(expr))}


def to_run_args(input_xs_vec, input_ys_vec, ctx=None, devices=None):
    """`->run-args`: marshals xs/ys once; they stay resident on the device for every later call."""
    xs = np.ascontiguousarray(input_xs_vec, dtype=np.float64)
    ys = np.ascontiguousarray(input_ys_vec, dtype=np.float64)
    ctx = ctx or Context(devices)
    ctx.set_points(xs, ys)
    return {"input-xs-vec": xs, "input-ys-vec": ys, "input-xs-count": len(xs), "ctx": ctx}


def score_batch(run_args, run_config, phenos):
    """One generation in one call: scores every phenotype; returns (scores, flags)."""
    b = PopulationBuilder()
    try:
        b.add_many([p["expr"] for p in phenos])
        return run_args["ctx"].score_batch(b.view(), int(run_config.get("max-leafs", default_max_leafs)))
    finally:
        b.close()


def score_fn(run_args, run_config, pheno):
    """`ops/score-fn`: any failure maps to min-score, never an exception (ops.clj:134-136)."""
    try:
        scores, _ = score_batch(run_args, run_config, [pheno])
        return float(scores[0])
    except L.ClosyrError as e:
        if e.code == L.E_PARSE:
            return min_score
        raise


def eval_vec_pheno(pheno, run_args, xs=None):
    """`ops-eval/eval-vec-pheno`: f(xs) as doubles (+Inf where the element is not a real number), None when the
    whole result is Indeterminate; raises where the reference throws (non-real constant expression)."""
    xs = run_args["input-xs-vec"] if xs is None else np.ascontiguousarray(xs, dtype=np.float64)
    b = PopulationBuilder()
    try:
        b.add(pheno["expr"])
        f, status = run_args["ctx"].eval_batch(b.view(), xs)
    finally:
        b.close()
    if status[0] == 1:
        return None
    if status[0] == 2:
        raise ArithmeticError("constant expression is not a real number: " + pheno["expr"])
    return f[0].tolist()


def compute_score_from_actuals_and_expecteds(pheno, f_of_xs, input_ys_vec, leafs, ctx):
    return ctx.score_from_values(f_of_xs, input_ys_vec, leafs)[0]


def compute_residual(expected, actual, ctx):
    return ctx.score_from_values([actual], [expected], 0)[2]


def extend_xs(input_xs_vec):
    """`extend-xs`: 35 % more points on either side of the domain (common.clj:262-291)."""
    xs = [float(v) for v in input_xs_vec]
    x_min, x_max = xs[0], xs[-1]
    x_range_sz = x_max - x_min
    pct = 0.35
    extra_pts = pct * len(xs)
    step = (pct * x_range_sz) / extra_pts
    k = int(np.ceil(extra_pts))              # (range extra-pts) yields ceil(extra-pts) indices
    x_head = [x_min - (i + 1) * step for i in range(k)][::-1]
    x_tail = [x_max + (i + 1) * step for i in range(k)]
    return {"xs": x_head + xs + x_tail, "x-head": x_head, "x-tail": x_tail}


def eval_vec_pheno_oversample(pheno, run_args, ext_info):
    """`eval-vec-pheno-oversample`: the curve on the extended domain, head/tail clamped to
    [min-10, max+10] of the middle section unless infinite (eval.clj:115-175)."""
    middle = eval_vec_pheno(pheno, run_args)
    max_y, min_y = max(middle), min(middle)

    def clamp(y):
        return y if np.isinf(y) else min(max_y + 10.0, max(y, min_y - 10.0))

    head = [clamp(y) for y in eval_vec_pheno(pheno, run_args, ext_info["x-head"])] if ext_info["x-head"] else []
    tail = [clamp(y) for y in eval_vec_pheno(pheno, run_args, ext_info["x-tail"])] if ext_info["x-tail"] else []
    return {"xs": ext_info["xs"], "ys": head + middle + tail}


def sort_population(run_args, k=None):
    """Indices of the scored (resident) population in the reference's report order
    (reverse (sort-by :score pop)), ops.clj:195-201, computed on the device."""
    return run_args["ctx"].rank(k)

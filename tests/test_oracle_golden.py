"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for the
scoring hot path.  File:line citations are relative to /root/reference.
Expressions are written in Symja FullForm, i.e. the canonical tree `->phenotype` produces
(src/closyr/ops/common.clj:135-146)."""
import math

import numpy as np
import pytest

from oracle import oracle as O

INF = float("inf")

X2_MINUS_HALF = "Plus(Rational(-1,2),Power(x,2))"        # x*x - 1/2  -> -1/2+x^2
X_MINUS_HALF = "Plus(Rational(-1,2),x)"                  # x - 1/2    -> -1/2+x
X_SIN_X_MINUS_HALF = "Plus(Rational(-1,2),Times(x,Sin(x)))"
SQRT_X_MINUS_1 = "Power(Plus(-1,x),Rational(1,2))"       # Sqrt(x-1)  -> Sqrt(-1+x)
E_MINUS_HALF = "Plus(Rational(-1,2),E)"


# ---------------------------------------------------------------- score-fn (test/closyr/ops_test.clj)
def test_score_fn_simple():  # ops_test.clj:18-28
    assert O.leaf_count(X2_MINUS_HALF) == 7
    assert O.score_fn(X2_MINUS_HALF, [0.5, 1.0, 2.0], [0, 1, 2], 40) == -3.0000147


def test_score_fn_simple_cuda_tree_order_is_identical_for_tiny_n():
    assert O.score_fn(X2_MINUS_HALF, [0.5, 1.0, 2.0], [0, 1, 2], 40, O.SUM_CUDA_TREE) == -3.0000147


def test_score_fn_too_big():  # ops_test.clj:30-39
    assert O.score_fn(X2_MINUS_HALF, [0.5, 1.0, 2.0], [0, 1, 2], 0) == O.MIN_SCORE


def test_score_fn_failing_eval_is_min_score():  # ops_test.clj:41-65 (nil / throwing eval)
    assert O.score_fn("Indeterminate", [0.5, 1.0, 2.0], [0, 1, 2], 40) == O.MIN_SCORE
    assert O.score_fn("Complex(0,1)", [0.5, 1.0, 2.0], [0, 1, 2], 40) == O.MIN_SCORE
    assert O.score_fn("Plus(((", [0.5, 1.0, 2.0], [0, 1, 2], 40) == O.MIN_SCORE


def test_compute_score_from_actuals_and_expecteds():  # ops_test.clj:68-76
    assert O.score_from_f([0.5], [1.0], 10) == -1.500015


def test_compute_score_without_length_deduction():  # ops_test.clj:103-112 (deduction forced to 0)
    assert O.score_from_f([0.5], [1.0], 0) == -1.5


def test_compute_residuals():  # ops_test.clj:192-203
    assert O.compute_residual(10.0, 5.0) == 5.0
    assert O.compute_residual(INF, 5.0) == O.MAX_RESID
    assert O.compute_residual(INF, INF) == O.MAX_RESID
    assert O.compute_residual(1.0, float("nan")) == O.MAX_RESID


def test_readme_initial_population_score():  # README.md:55 data, all-x population (derived)
    xs = [0, 1, 2, 3, 4, 5, 6]
    ys = [1, 2, 30, 4, 5, 6, 10]
    assert O.score_fn("x", xs, ys, 40) == -38.57143242857143


# ---------------------------------------------------------------- eval-vec-pheno (test/closyr/ops_eval_test.clj)
def test_eval_x_minus_half():  # ops_eval_test.clj:60-66, 153-159
    assert O.eval_vec_pheno(X_MINUS_HALF, [0.5, 1.0]).tolist() == [0.0, 0.5]


def test_eval_y_equals_x():  # ops_eval_test.clj:82-88 (x*1 canonicalises to x), :238-242
    assert O.eval_vec_pheno("x", [0.5, 1.0]).tolist() == [0.5, 1.0]
    assert O.eval_vec_pheno("x", [0.0, 0.5, 1.0]).tolist() == [0.0, 0.5, 1.0]


def test_eval_infinite_x():  # ops_eval_test.clj:161-166
    assert O.eval_vec_pheno(X_MINUS_HALF, [INF]).tolist() == [INF]


def test_eval_x_minus_infinity():  # ops_eval_test.clj:168-173; x-Infinity canonicalises to -Infinity
    assert O.eval_vec_pheno("DirectedInfinity(-1)", [0.0]).tolist() == [-INF]
    assert O.eval_vec_pheno("Plus(x,DirectedInfinity(-1))", [0.0]).tolist() == [-INF]


def test_eval_constant_e_minus_half():  # ops_eval_test.clj:176-180
    assert O.eval_vec_pheno(E_MINUS_HALF, [0.5]).tolist() == [2.218281828459045]


def test_eval_constant_half():  # ops_eval_test.clj:182-186 (1 - 1/2 -> 1/2)
    assert O.eval_vec_pheno("Rational(1,2)", [0.5]).tolist() == [0.5]


def test_eval_single_point():  # ops_eval_test.clj:189-193
    assert O.eval_vec_pheno(X_MINUS_HALF, [0.5]).tolist() == [0.0]


def test_eval_x_sin_x():  # ops_eval_test.clj:195-199
    assert O.eval_vec_pheno(X_SIN_X_MINUS_HALF, [0.0, 0.5, 1.0]).tolist() == \
        [-0.5, -0.2602872306978985, 0.3414709848078965]


def test_eval_complex_results_are_infinite():  # ops_eval_test.clj:201-206
    assert O.eval_vec_pheno(SQRT_X_MINUS_1, [0.0, 0.5, 1.0]).tolist() == [INF, INF, 0.0]


@pytest.mark.parametrize("expr,expected", [
    ("1", [1.0, 1.0, 1.0]),          # ops_eval_test.clj:208-212
    ("-1", [-1.0, -1.0, -1.0]),      # :220-224
    ("0.123", [0.123, 0.123, 0.123]),  # :226-230
    ("0", [0.0, 0.0, 0.0]),          # :232-236 (-1+1 -> 0)
])
def test_eval_constants(expr, expected):
    assert O.eval_vec_pheno(expr, [0.0, 0.5, 1.0]).tolist() == expected


def test_eval_whole_result_indeterminate_is_nil():  # ops_eval_test.clj:98-104
    assert O.eval_vec_pheno("Indeterminate", [0.5]) is None


def test_eval_complex_constant_throws():  # ops_common_test.clj:31-32 + eval.clj:83-96
    with pytest.raises(ArithmeticError):
        O.eval_vec_pheno("Complex(0,1)", [0.5, 1.0, 2.0])


@pytest.mark.parametrize("n", [20, 40, 80, 160, 320])
def test_eval_benchmark_vectors_exact(n):  # ops_eval_test.clj:245-301
    xs = np.arange(n, dtype=np.float64)
    assert O.eval_vec_pheno(X2_MINUS_HALF, xs).tolist() == [x * x - 0.5 for x in xs.tolist()]


def test_constant_ast_quirk_decodes_per_argument():  # eval.clj:109 (A13): size-1 == n
    assert O.eval_vec_pheno(E_MINUS_HALF, [0.5, 1.0]).tolist() == [-0.5, INF]


# ---------------------------------------------------------------- leafCount (ops.clj:127; A12)
@pytest.mark.parametrize("expr,leafs", [
    ("x", 1), ("Rational(1,2)", 3), ("Times(-1,x)", 3), ("Sin(x)", 2),
    ("Plus(Rational(-1,2),Power(x,2))", 7), ("Complex(0,1)", 3),
])
def test_leaf_count(expr, leafs):
    assert O.leaf_count(expr) == leafs


def test_score_batch_matches_single_and_uses_threads():
    xs = np.linspace(0, 10, 50)
    ys = 2.0 + np.cos(xs)
    exprs = [X2_MINUS_HALF, "x", X_SIN_X_MINUS_HALF, SQRT_X_MINUS_1, "Indeterminate"] * 7
    got, used = O.score_batch(exprs, xs, ys, 40)
    assert used >= 1
    for e, s in zip(exprs, got):
        assert s == O.score_fn(e, xs, ys, 40)
    assert math.isfinite(got[0])

"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(closyr_b200/libclosyr_b200.so) and is compared with the CPU oracle on the same seeded inputs.

Tolerances: scores of transcendental expressions 1e-9 relative (BASELINE.json north star); everything that
is integer, index or pure IEEE arithmetic (+, *, /, sqrt, integer powers, the reduction tree, the score
epilogue, flags, ranking) is bit-exact.
"""
import math

import numpy as np
import pytest

import closyr_b200 as cb
from closyr_b200 import _lib as L
from closyr_b200 import ops
from oracle import oracle as O

pytestmark = pytest.mark.gpu

INF = float("inf")
REL_TOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    c = cb.Context([0])
    yield c
    c.close()


def synth_points(n):
    """xs over the reference's example span, ys = the integration-test objective."""
    xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
    ys = 2.0 + 1.5 * xs / np.pi + np.cos(xs)
    return xs, ys


def gpu_scores(ctx, texts, xs, ys, max_leafs=40):
    b = cb.PopulationBuilder()
    b.add_many(texts)
    ctx.set_points(xs, ys)
    s, f = ctx.score_batch(b.view(), max_leafs)
    b.close()
    return s, f


def assert_scores_close(got, want, texts=None):
    err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    i = int(np.argmax(err))
    assert err[i] <= REL_TOL, "expr %d %s: gpu %r oracle %r" % (i, texts[i] if texts else "", got[i], want[i])


PARITY_REPORT = []      # one record per population checked (printed at the end of the session with `pytest -s`)


def assert_population_scores_match(texts, xs, ys, max_leafs, got, probes=8, explain_limit=400):
    """The bar: 1e-9 relative on every score that fp64 arithmetic determines.
      * `ok` = the oracle's own score is stable when every result of its math library moves by +-1 ulp;
      * a deviation above 1e-9 on a STABLE expression that such a perturbation (2 ulp: the documented bound of the CUDA
        routines) cannot reproduce is a real discrepancy: none is allowed;
      * a deviation anywhere else must be reproducible the same way (knife edges: a point on a domain boundary or a
        branch cut, chaotic arguments such as Cos(E^(x^2))): no unexplained deviation is allowed there either;
      * explained deviations among the stable expressions are capped at 0.5 % of the population.
    The stable fraction is what it is (reported, not a gate): on a transcendental-heavy population on extreme points
    most expressions are chaotic, and the bar for those is "explained"."""
    want, ok = O.score_batch_conditioned(texts, xs, ys, max_leafs, probes=probes)
    err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    bad = np.nonzero(err > REL_TOL)[0]
    check = bad if len(bad) <= explain_limit else np.random.default_rng(0).choice(bad, explain_limit, replace=False)
    unexplained = [int(i) for i in check
                   if not O.score_discrepancy_explained(texts[i], xs, ys, max_leafs, got[i], want[i])]
    rec = {"expressions": len(texts), "points": len(xs), "stable_fraction": float(ok.mean()),
           "within_1e-9": float((err <= REL_TOL).mean()), "deviating": int(len(bad)), "deviating_stable": int(ok[bad].sum()),
           "explain_checked": int(len(check)), "unexplained": len(unexplained),
           "max_rel_err_stable": float(err[ok].max()) if ok.any() else None}
    PARITY_REPORT.append(rec)
    assert not unexplained, "unexplained deviations, first: expr %d %s: gpu %r oracle %r" % (
        unexplained[0], texts[unexplained[0]], got[unexplained[0]], want[unexplained[0]])
    assert ok[bad].sum() <= max(2, 0.005 * len(texts)), "too many knife-edge cases among the stable expressions: %r" % rec
    return want, ok


# ------------------------------------------------------------------ the reference's own known-answer tests
def run_args(ctx, xs, ys):
    return ops.to_run_args(xs, ys, ctx=ctx)


def test_score_fn_simple(ctx):  # test/closyr/ops_test.clj:18-28
    ra = run_args(ctx, [0.5, 1.0, 2.0], [0, 1, 2])
    assert ops.score_fn(ra, {"max-leafs": ops.default_max_leafs}, ops.to_phenotype("x*x-1/2")) == -3.0000147


def test_score_fn_too_big(ctx):  # ops_test.clj:30-39
    ra = run_args(ctx, [0.5, 1.0, 2.0], [0, 1, 2])
    assert ops.score_fn(ra, {"max-leafs": 0}, ops.to_phenotype("x*x-1/2")) == ops.min_score


def test_score_fn_failing_eval(ctx):  # ops_test.clj:41-65: nil / throwing eval => min-score
    ra = run_args(ctx, [0.5, 1.0, 2.0], [0, 1, 2])
    cfg = {"max-leafs": 40}
    assert ops.score_fn(ra, cfg, {"expr": "Indeterminate"}) == ops.min_score      # eval returns nil
    assert ops.score_fn(ra, cfg, {"expr": "Complex(0,1)"}) == ops.min_score       # toNumber throws
    assert ops.score_fn(ra, cfg, {"expr": "Plus((("}) == ops.min_score            # any exception


def test_compute_score_from_actuals_and_expecteds(ctx):  # ops_test.clj:68-76, 103-112
    p = ops.to_phenotype("Sin(x)+1/2")
    assert ops.compute_score_from_actuals_and_expecteds(p, [0.5], [1.0], 10, ctx) == -1.500015
    assert ops.compute_score_from_actuals_and_expecteds(p, [0.5], [1.0], 0, ctx) == -1.5


def test_compute_residuals(ctx):  # ops_test.clj:192-203
    assert ops.compute_residual(10.0, 5.0, ctx) == 5.0
    assert ops.compute_residual(INF, 5.0, ctx) == ops.max_resid
    assert ops.compute_residual(INF, INF, ctx) == ops.max_resid
    assert ops.compute_residual(1.0, float("nan"), ctx) == ops.max_resid


def test_readme_initial_population(ctx):  # README.md:55 (config 1 data), all-x population
    xs, ys = [0, 1, 2, 3, 4, 5, 6], [1, 2, 30, 4, 5, 6, 10]
    ra = run_args(ctx, xs, ys)
    scores, _ = ops.score_batch(ra, {"max-leafs": 40}, [ops.to_phenotype("x*1")] * 25)
    assert scores.tolist() == [-38.57143242857143] * 25


@pytest.mark.parametrize("expr,xs,want", [
    ("x-1/2", [0.5, 1.0], [0.0, 0.5]),                       # ops_eval_test.clj:60-66
    ("x*1", [0.5, 1.0], [0.5, 1.0]),                         # :82-88
    ("x-1/2", [INF], [INF]),                                 # :161-166
    ("x-Infinity", [0.0], [-INF]),                           # :168-173
    ("E-1/2", [0.5], [2.218281828459045]),                   # :176-180
    ("1-1/2", [0.5], [0.5]),                                 # :182-186
    ("x-1/2", [0.5], [0.0]),                                 # :189-193
    ("x*Sin(x)-1/2", [0.0, 0.5, 1.0], [-0.5, -0.2602872306978985, 0.3414709848078965]),   # :195-199
    ("Sqrt(x-1)", [0.0, 0.5, 1.0], [INF, INF, 0.0]),         # :201-206 complex => +Inf
    ("1", [0.0, 0.5, 1.0], [1.0, 1.0, 1.0]),                 # :208-212
    ("Infinity", [0.0, 0.5, 1.0], [0.0, 0.5, 1.0]),          # :214-218 (->phenotype replaces it by x)
    ("-1", [0.0, 0.5, 1.0], [-1.0, -1.0, -1.0]),             # :220-224
    ("0.123", [0.0, 0.5, 1.0], [0.123, 0.123, 0.123]),       # :226-230
    ("-1+1", [0.0, 0.5, 1.0], [0.0, 0.0, 0.0]),              # :232-236
    ("x", [0.0, 0.5, 1.0], [0.0, 0.5, 1.0]),                 # :238-242
    ("Indeterminate", [0.5], [0.5]),                         # :90-96 (->phenotype replaces it by x)
])
def test_eval_vec_pheno_known_answers(ctx, expr, xs, want):
    ra = run_args(ctx, xs, [0.0] * len(xs))
    got = ops.eval_vec_pheno(ops.to_phenotype(expr), ra)
    if expr == "x*Sin(x)-1/2":
        assert got[0] == want[0]
        np.testing.assert_allclose(got, want, rtol=1e-15)
    else:
        assert got == want


def test_eval_nil_and_throw(ctx):  # ops_eval_test.clj:98-112; eval.clj:83-96
    ra = run_args(ctx, [0.5], [0.0])
    assert ops.eval_vec_pheno({"expr": "Indeterminate"}, ra) is None
    with pytest.raises(ArithmeticError):
        ops.eval_vec_pheno({"expr": "Complex(0,1)"}, ra)


@pytest.mark.parametrize("n", [20, 40, 80, 160, 320])
def test_eval_benchmark_vectors_exact(ctx, n):  # ops_eval_test.clj:245-301
    xs = np.arange(n, dtype=np.float64)
    ra = run_args(ctx, xs, xs)
    assert ops.eval_vec_pheno(ops.to_phenotype("x*x-1/2"), ra) == [x * x - 0.5 for x in xs.tolist()]


def test_constant_ast_quirk(ctx):  # eval.clj:109
    ra = run_args(ctx, [0.5, 1.0], [0.0, 0.0])
    assert ops.eval_vec_pheno(ops.to_phenotype("E-1/2"), ra) == [-0.5, INF]
    assert ops.score_fn(ra, {"max-leafs": 40}, ops.to_phenotype("E-1/2")) == O.score_fn("Plus(Rational(-1,2),E)", [0.5, 1.0], [0, 0], 40)


# ------------------------------------------------------------------ seeded random populations vs the oracle
@pytest.mark.parametrize("n_points", [7, 100, 128, 129, 1000, 4099])
@pytest.mark.parametrize("mix,max_leafs", [(0, 40), (1, 80)])
def test_random_population_matches_oracle(ctx, n_points, mix, max_leafs):
    xs, ys = synth_points(n_points)
    b = cb.PopulationBuilder()
    texts = b.generate(888 + n_points, 400, max_leafs, mix, want_texts=True)
    ctx.set_points(xs, ys)
    got, flags = ctx.score_batch(b.view(), max_leafs)
    # `ok` = the score is determined by fp64 arithmetic (the oracle agrees with itself when every libm result
    # moves by one ulp); chaotic expressions such as Cos(E^(x^2)) cannot be held to 1e-9 by any implementation
    want, ok = assert_population_scores_match(texts, xs, ys, max_leafs, got)
    assert np.isfinite(got).all() and (got <= 0).all()
    assert int(np.argmax(got)) == int(np.argmax(want))          # identical best-expression selection
    assert ((flags & L.R_MIN_SCORE) != 0).tolist() == (want == O.MIN_SCORE).tolist()
    # device ranking == the host's (reverse (sort-by :score ...)) on the same scores, bit for bit
    order = ctx.rank()
    host = np.argsort(got, kind="stable")[::-1]
    assert order.tolist() == host.tolist()


def arithmetic_population(rng, count):
    """Expressions that only use correctly rounded IEEE operations: the GPU must match bit for bit."""
    leaves = ["x", "x", "x", "1/2", "1/10", "2", "10", "-1", "1.1", "0.9", "3/7"]

    def gen(d):
        if d == 0 or rng.random() < 0.25:
            return leaves[rng.integers(len(leaves))]
        k = rng.integers(7)
        if k == 0:
            return "Plus(%s,%s)" % (gen(d - 1), gen(d - 1))
        if k == 1:
            return "Times(%s,%s)" % (gen(d - 1), gen(d - 1))
        if k == 2:
            return "Plus(%s,%s,%s)" % (gen(d - 1), gen(d - 1), gen(d - 1))
        if k == 3:
            return "Power(%s,%d)" % (gen(d - 1), [2, -1, -2, 3, 5, -3][rng.integers(6)])
        if k == 4:
            return "Power(%s,Rational(%d,2))" % (gen(d - 1), [1, -1][rng.integers(2)])
        if k == 5:
            return "Times(%s,%s,%s)" % (gen(d - 1), gen(d - 1), gen(d - 1))
        return "Abs(%s)" % gen(d - 1)

    return [gen(5) for _ in range(count)]


@pytest.mark.parametrize("n_points", [3, 127, 128, 1025, 5000])
def test_arithmetic_population_is_bit_exact(ctx, n_points):
    rng = np.random.default_rng(12345 + n_points)
    texts = [cb.to_fullform(t) for t in arithmetic_population(rng, 300)]
    xs = rng.uniform(-3.0, 8.0, n_points)
    xs[0] = 0.0
    ys = rng.uniform(-5.0, 5.0, n_points)
    got, _ = gpu_scores(ctx, texts, xs, ys, 200)
    want, _ = O.score_batch(texts, xs, ys, 200, sum_mode=O.SUM_CUDA_TREE)
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, "expr %s gpu %r oracle %r" % (texts[bad[0]], got[bad[0]], want[bad[0]])


def check_eval_against_oracle(texts, xs, f, status):
    for i, t in enumerate(texts):
        try:
            want, ok = O.eval_conditioned(t, xs)
        except ArithmeticError:
            assert status[i] == 2, t
            continue
        if want is None:
            assert status[i] == 1, t
            continue
        assert status[i] == 0, t
        explained, n_bad = O.eval_discrepancy_explained(t, xs, f[i], want)
        assert explained, "%s: %d points differ from the oracle beyond what libm jitter explains" % (t, n_bad)
        with np.errstate(invalid="ignore"):
            same = (f[i] == want) | (np.abs(f[i] - want) <= 1e-9 * np.abs(want))
        assert (ok & ~same).sum() <= max(1, len(xs) // 100), t


def test_eval_batch_matches_oracle(ctx):
    xs, _ = synth_points(300)
    xs = np.concatenate([[-2.0, -0.5, 0.0], xs])
    b = cb.PopulationBuilder()
    texts = b.generate(4242, 200, 40, 0, want_texts=True)
    f, status = ctx.eval_batch(b.view(), xs)
    check_eval_against_oracle(texts, xs, f, status)


SEMANTIC_CASES = [
    # poles: 1/0.0 is ComplexInfinity
    "Power(x,-1)", "Plus(1,Power(x,-1))", "Power(Plus(1,Power(x,-1)),-1)", "Times(x,Power(x,-1))",
    "Power(E,Times(-1,Power(x,-2)))", "Plus(Power(x,-1),Times(-1,Power(x,-1)))", "Sin(Power(x,-1))",
    "Log(x)", "Power(E,Log(x))", "Power(E,Times(2,Log(x)))", "Log(Power(x,-1))", "Abs(Power(x,-1))",
    # complex excursions that come back to the real line (A9)
    "Cos(Power(x,Rational(1,2)))", "Power(Power(Plus(-1,x),Rational(1,2)),2)",
    "Times(Power(Plus(-1,x),Rational(1,2)),Power(Plus(-2,x),Rational(1,2)))",
    "Times(Power(Plus(-1,x),Rational(1,2)),Power(Plus(-2,x),Rational(-1,2)))",
    "Abs(Power(x,Rational(1,2)))", "Abs(Log(x))", "Power(Log(x),2)", "Cos(ArcCos(x))", "Sin(ArcSin(x))",
    "Abs(ArcSin(x))", "Abs(ArcCos(x))", "Power(x,Rational(3,2))", "Power(x,0.5)", "Power(x,1.5)", "Power(x,-2.5)",
    "Tan(Power(x,Rational(1,2)))", "Power(E,Power(x,Rational(1,2)))", "Log(Log(x))", "Power(Log(x),-1)",
    "Power(Plus(-1,x),Rational(-1,2))", "ArcSin(Times(2,x))", "ArcCos(Times(2,x))", "Power(ArcSin(x),2)",
    "Times(Power(x,Rational(1,2)),Power(x,Rational(1,2)))", "Power(Sin(x),Rational(1,2))", "Power(x,x)",
    "Power(2,x)", "Power(-2,x)", "Power(x,Sin(x))", "Power(Cos(x),x)", "Power(Sin(x),Cos(x))",
    # literals that are not real numbers
    "Plus(x,Complex(0,1))", "Times(x,Complex(0,1))", "Power(Times(x,Complex(0,1)),2)", "Plus(x,DirectedInfinity())",
    "Times(x,DirectedInfinity())", "Plus(x,Indeterminate)", "Plus(x,DirectedInfinity(1))", "Times(x,DirectedInfinity(-1))",
    "Power(E,Times(x,DirectedInfinity(-1)))", "unknownfn(x)", "Plus(x,y)",
    # constants
    "Plus(Rational(-1,2),E)", "Sin(Rational(1,2))", "Power(-2,Rational(1,2))", "Log(-1)", "Power(0,-1)", "Pi",
    "DirectedInfinity(1)", "DirectedInfinity()", "Indeterminate", "Power(2,Rational(1,2))",
    # overflow / big exponents
    "Power(E,Power(E,Power(E,x)))", "Power(x,40)", "Power(x,-40)", "Power(Times(1000000,x),32)", "Power(x,100)",
]


def test_semantic_cases_match_oracle(ctx):
    xs = np.array([-3.0, -1.5, -1.0, -0.5, -0.0, 0.0, 0.25, 0.5, 1.0, 1.5, 2.0, 3.0, 7.5, 100.0, 1e6, 1e300])
    ys = np.linspace(-1.0, 2.0, len(xs))
    got, flags = gpu_scores(ctx, SEMANTIC_CASES, xs, ys, 40)
    want, _ = O.score_batch(SEMANTIC_CASES, xs, ys, 40)
    assert_scores_close(got, want, SEMANTIC_CASES)
    ctx.set_points(xs, ys)
    b = cb.PopulationBuilder()
    b.add_many(SEMANTIC_CASES)
    f, status = ctx.eval_batch(b.view(), xs)
    check_eval_against_oracle(SEMANTIC_CASES, xs, f, status)


def balanced_tree(depth, leaf):
    """A complete binary tree of alternating Plus / Times: both operands of every node need the same number of spill
    slots, so the compiler runs out of the four register slots at depth 5 (ISA: CL_ST / CL_ADDS / CL_MULS with a >= 4)."""
    k = [0]

    def gen(d):
        if d == 0:
            k[0] += 1
            return leaf(k[0])
        return "%s(%s,%s)" % ("Plus" if d % 2 else "Times", gen(d - 1), gen(d - 1))

    return gen(depth)


def test_programs_with_deep_spill_slots(ctx):
    """Spill slots beyond the four fast ones (local memory in the complete-algebra interpreter, no fast pass)."""
    texts = [
        balanced_tree(6, lambda i: "%s(Plus(%d,x))" % (["Sin", "Cos"][i % 2], i)),                      # real everywhere
        balanced_tree(6, lambda i: "Power(Plus(%d,Times(-1,x)),Rational(1,2))" % (i % 9)),              # complex beyond x = i
        balanced_tree(5, lambda i: ["Log(x)", "Power(x,-1)", "ArcSin(x)", "Sin(x)"][i % 4]),            # poles, branch cuts
    ]
    b = cb.PopulationBuilder()
    b.add_many(texts)
    assert ((b.arrays()["meta"] >> 24) > 4).all()
    xs, ys = synth_points(700)
    got, flags = gpu_scores(ctx, texts, xs, ys, 1000)
    want, _ = O.score_batch(texts, xs, ys, 1000)
    assert_scores_close(got, want, texts)
    f, status = ctx.eval_batch(b.view(), xs[:200])
    check_eval_against_oracle(texts, xs[:200], f, status)


def test_scores_are_identical_for_every_tiling(ctx):
    """The reduction tree is fixed: a small population (many tiles per expression, partials + finalize kernel)
    and the same expressions inside a large population (one tile per expression) give bit-identical scores."""
    n = 70000
    xs, ys = synth_points(n)
    b = cb.PopulationBuilder()
    texts = b.generate(99, 24, 40, 0, want_texts=True)
    ctx.set_points(xs, ys)
    small, _ = ctx.score_batch(b.view(), 40)
    big = cb.PopulationBuilder()
    big.add_many(texts)
    big.generate(100, 20000, 10, 0)
    large, _ = ctx.score_batch(big.view(), 40)
    assert small.tolist() == large[:24].tolist()
    assert_population_scores_match(texts, xs, ys, 40, small)


def test_resident_api_and_launch_count(ctx):
    xs, ys = synth_points(512)
    b = cb.PopulationBuilder()
    b.generate(5, 3000, 40, 0)
    ctx.set_points(xs, ys)
    before = ctx.kernel_launches
    ctx.upload_population(b.view())
    ms = ctx.score_resident(40)
    s1, f1 = ctx.download_scores()
    s2, f2 = ctx.score_batch(b.view(), 40)
    assert ms > 0.0
    assert ctx.kernel_launches > before
    assert s1.tolist() == s2.tolist() and f1.tolist() == f2.tolist()


def test_empty_inputs(ctx):
    b = cb.PopulationBuilder()
    ctx.set_points([1.0, 2.0], [1.0, 2.0])
    s, f = ctx.score_batch(b.view(), 40)
    assert len(s) == 0
    b.add("x")
    ctx.set_points([], [])
    s, f = ctx.score_batch(b.view(), 40)
    assert s.tolist() == [ops.min_score]          # (reduce max []) throws in the reference => min-score


def test_duplicates_score_bit_equal_at_baseline_size(ctx):
    """Config 3 shape (pop 100K x 4096 points): equal expressions get bit-equal scores wherever they sit, every
    score is a finite number <= 0, and the device ranking is sorted."""
    xs, ys = synth_points(4096)
    b = cb.PopulationBuilder()
    b.generate(888, 50000, 40, 0)
    b.generate(888, 50000, 40, 0)                 # the same 50K again
    ctx.set_points(xs, ys)
    s, f = ctx.score_batch(b.view(), 40)
    assert s[:50000].tolist() == s[50000:].tolist()
    assert np.isfinite(s).all() and (s <= 0).all()
    order = ctx.rank(1000)
    assert (np.diff(s[order]) <= 0).all()
    assert s[order[0]] == s.max()


# ------------------------------------------------------------------ the reference's own expression corpus
def csv_style_points(n):
    """SURVEY.md section 8d, second objective: piecewise data in the style of resources/csvs/test_inputs_*.csv
    (a dozen hand-placed knots over a wide x range), linearly interpolated to n points."""
    kx = np.array([0.0, 2.0, 5.0, 9.0, 14.0, 20.0, 27.0, 35.0, 44.0, 54.0])
    ky = np.array([1.0, 1.5, 3.0, 0.0, -2.0, -12.0, -20.0, -24.0, -11.0, 4.0])
    xs = np.linspace(kx[0], kx[-1], n)
    return xs, np.interp(xs, kx, ky)


def golden_corpus_population():
    """SURVEY.md section 8d, second population source: the 96 golden results of ops_modify_test.clj:463-751 (plus the
    strings of ops_test / ops_eval_test), their pairwise crossovers (modify.clj:142-196) and one more mutation of
    each (initialize.clj:35-665) -- Symja-canonical shapes a real run produces."""
    import json
    import os
    from closyr_b200 import ga as G
    forms = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "symja_canonical_forms.json")))
    base = [m["expr"] for m in forms["mutations"]] + [m["expr"] for m in forms["extra"]]
    texts = [cb.to_fullform(t) for t in base]
    n = len(base)
    for k in range(400):
        i, j = (7 * k + 3) % n, (13 * k + 5) % n
        texts.append(G.crossover_expr(base[i], base[j], 60, seed=1000 + k))
    for k in range(2 * n):
        texts.append(G.mutate_expr(base[k % n], (5 * k + 1) % 96, seed=2000 + k))
    return texts


@pytest.mark.parametrize("points", ["synthetic", "csv-style"])
def test_golden_corpus_with_crossovers_and_mutations(ctx, points):
    xs, ys = synth_points(1024) if points == "synthetic" else csv_style_points(1024)
    texts = golden_corpus_population()
    assert len(texts) > 700
    got, flags = gpu_scores(ctx, texts, xs, ys, 40)
    want, ok = assert_population_scores_match(texts, xs, ys, 40, got)
    assert ((flags & L.R_TOO_BIG) != 0).tolist() == [O.leaf_count(t) > 40 for t in texts]
    stable = np.nonzero(ok)[0]
    assert int(stable[np.argmax(got[stable])]) == int(stable[np.argmax(want[stable])])     # same best expression


@pytest.mark.parametrize("seed,mix", [(31, 0), (32, 1), (33, 0)])
def test_generated_expressions_on_extreme_points(ctx, seed, mix):
    """Random expressions on the points where IEEE arithmetic and the value algebra part ways: signed zeros,
    denormals, huge magnitudes and both infinities (ops_eval_test.clj:161-173 pins x = +Inf and x - Infinity)."""
    xs = np.array([-np.inf, -1e308, -1e154, -1e6, -3.0, -1.0, -0.5, -1e-300, -5e-324, -0.0, 0.0, 5e-324, 1e-300,
                   0.5, 1.0, 2.0, 710.0, 1e6, 1e154, 1e308, np.inf])
    b = cb.PopulationBuilder()
    texts = b.generate(seed, 300, 30, mix, want_texts=True)
    f, status = ctx.eval_batch(b.view(), xs)
    check_eval_against_oracle(texts, xs, f, status)
    ys = np.linspace(-2.0, 2.0, len(xs))
    got, _ = gpu_scores(ctx, texts, xs, ys, 40)
    assert_population_scores_match(texts, xs, ys, 40, got)


def test_whole_configs2_population_against_the_oracle(ctx):
    """BASELINE.json configs[2] at full size: every one of the 100K expressions x 4096 points against the oracle
    (the bench only spot-checks the first 2000)."""
    import bench
    pop_n, n_points, max_leafs, mix, _ = bench.WORKLOADS["c3"]
    xs, ys = bench.synth_points(n_points)
    b = cb.PopulationBuilder()
    texts = b.generate(bench.SEED, pop_n, max_leafs, mix, want_texts=True)
    ctx.set_points(xs, ys)
    got, flags = ctx.score_batch(b.view(), max_leafs)
    want, ok = assert_population_scores_match(texts, xs, ys, max_leafs, got, probes=3, explain_limit=300)
    # what the GA does with the scores: the same best expression, the same outcome of compete's `>=` on neighbours
    # (src/closyr/ga.clj:43) wherever both scores are determined by fp64
    stable = np.nonzero(ok)[0]
    assert int(stable[np.argmax(got[stable])]) == int(stable[np.argmax(want[stable])])
    a, c = np.arange(0, pop_n - 1, 2), np.arange(1, pop_n, 2)
    both = ok[a] & ok[c]
    assert ((got[a] >= got[c]) == (want[a] >= want[c]))[both].all()
    assert ((got == O.MIN_SCORE) == (want == O.MIN_SCORE)).all()


def test_ranking_a_million_scores_with_ties(ctx):
    """closyr_rank_resident at the size of configs[3] (2^20 + 1 expressions, so the network pads to 2^21): the tiled
    bitonic passes (shared memory) and the cross-tile steps give the host's order, ties in descending index."""
    kinds = ["x", "Plus(1,x)", "Times(2,x)", "Sin(x)", "Plus(x,Cos(x))", "Power(x,2)", "Log(x)", "Power(x,-1)",
             "Indeterminate", "Plus(Rational(1,2),x)", "Power(E,x)", "Times(x,x,x)"]
    n = (1 << 20) + 1
    texts = [kinds[(i * 7 + i // 13) % len(kinds)] for i in range(n)]
    xs = np.array([0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0])
    ys = np.array([1.0, 2.0, 30.0, 4.0, 5.0, 6.0, 10.0])
    b = cb.PopulationBuilder()
    b.add_many(texts)
    ctx.set_points(xs, ys)
    scores, _ = ctx.score_batch(b.view(), 40)
    order = ctx.rank()
    want = np.argsort(scores, kind="stable")[::-1]
    assert order.tolist() == want.tolist()
    top = ctx.rank(100)
    assert top.tolist() == want[:100].tolist()


def test_cta_order_and_cta_size_do_not_change_a_score():
    """score_kernel takes its expressions in the grouped order of order_kernel when several CTAs share an SM
    (kernels.cu); CLOSYR_ORDER=0 / 1 forces it off / on (read once per process).  Scores and flags are bit-identical."""
    import os, subprocess, sys, hashlib
    prog = (
        "import sys, hashlib, numpy as np; sys.path.insert(0, '.'); import closyr_b200 as cb\n"
        "b = cb.PopulationBuilder(); b.generate(4242, 5000, 40, 0)\n"
        "ctx = cb.Context([0]); out = []\n"
        "for n in (130, 1024, 5000):\n"
        "    xs = 0.7 * np.arange(n) / n; ys = np.cos(xs); ctx.set_points(xs, ys)\n"
        "    s, f = ctx.score_batch(b.view(), 40)\n"
        "    out.append(hashlib.sha256(np.asarray(s).tobytes() + (np.asarray(f) & 0xFB).astype(np.uint8).tobytes()).hexdigest())   # without the diagnostic bit\n"
        "print(' '.join(out))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    got = []
    # order off / on; CTAs capped at 2 and 8 warps (more rounds per tile, the lean instantiation of the fast pass where the
    # default takes the prefetching one)
    for knobs in ({"CLOSYR_ORDER": "0"}, {"CLOSYR_ORDER": "1"}, {"CLOSYR_SCORE_WARPS": "2"}, {"CLOSYR_SCORE_WARPS": "8", "CLOSYR_ORDER": "1"}):
        env = dict(os.environ, **knobs)
        r = subprocess.run([sys.executable, "-c", prog], cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        got.append(r.stdout.strip().splitlines()[-1])
    assert len(set(got)) == 1 and len(got[0].split()) == 3, got


def test_malformed_layouts_are_refused_and_the_context_stays_usable(ctx):
    """upload_shard starts the host-to-device copies first and walks the offsets while the DMA runs; a layout the TMA
    staging could not take (include/closyr_b200.h: 16-byte aligned programs, monotonic offsets, CL_END last) must come back
    as CLOSYR_E_ARG with nothing launched, and the next valid call must work (no sticky CUDA error)."""
    import ctypes as C
    xs, ys = synth_points(300)
    ctx.set_points(xs, ys)
    b = cb.PopulationBuilder()
    b.generate(77, 64, 40, 0)
    good_scores, good_flags = ctx.score_batch(b.view(), 40)
    a = b.arrays()

    def population(arr):
        keep = {k: np.ascontiguousarray(v) for k, v in arr.items()}
        pop = L.Population()
        pop.n_expr = len(keep["meta"])
        pop.code = keep["code"].ctypes.data_as(C.POINTER(C.c_uint32))
        pop.code_off = keep["code_off"].ctypes.data_as(C.POINTER(C.c_uint32))
        pop.consts = keep["consts"].ctypes.data_as(C.POINTER(C.c_double))
        pop.const_off = keep["const_off"].ctypes.data_as(C.POINTER(C.c_uint32))
        pop.meta = keep["meta"].ctypes.data_as(C.POINTER(C.c_uint32))
        pop.aux = keep["aux"].ctypes.data_as(C.POINTER(C.c_uint32))
        return pop, keep

    def broken(edit):
        arr = {k: v.copy() for k, v in a.items()}
        edit(arr)
        return population(arr)

    def misaligned(arr): arr["code_off"][5] += 1
    def decreasing(arr): arr["code_off"][7] = arr["code_off"][5]
    def empty_program(arr): arr["code_off"][9] = arr["code_off"][8]
    def no_end(arr): arr["code"][arr["code_off"][4] - 1] = 9          # CL_NEG where CL_END must be
    def odd_consts(arr): arr["const_off"][3] += 1
    for edit in (misaligned, decreasing, empty_program, no_end, odd_consts):
        pop, keep = broken(edit)
        with pytest.raises(L.ClosyrError) as err:
            ctx.score_batch(pop, 40)
        assert "population layout" in str(err.value), (edit.__name__, str(err.value))
        s, f = ctx.score_batch(b.view(), 40)
        assert s.tolist() == good_scores.tolist() and f.tolist() == good_flags.tolist(), edit.__name__
    pop, keep = population(a)                                             # the same arrays untouched: accepted
    s, f = ctx.score_batch(pop, 40)
    assert s.tolist() == good_scores.tolist()


def test_zz_parity_report():
    """Prints what the population checks of this session saw (pytest -s)."""
    for rec in PARITY_REPORT:
        print("PARITY", rec)

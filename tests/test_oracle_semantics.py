"""The value algebra the oracle (and the CUDA path) implement beyond what the reference's tests pin
(DESIGN.md "Semantics"): every rule here is an ASSUMPTION about Symja, written down so
that a run on a real JVM can falsify it quickly.  Tags: 0 real, 1 complex, 2 ComplexInfinity, 3 Indeterminate."""
import math

import numpy as np
import pytest

from oracle import oracle as O

REAL, CPLX, CINF, INDET = 0, 1, 2, 3


def val(expr, x):
    re, im, tag = O.eval_raw(expr, [x])
    return re[0], im[0], int(tag[0])


@pytest.mark.parametrize("expr,x,tag", [
    ("Power(x,-1)", 0.0, CINF),                          # 1/0.0 -> ComplexInfinity (A10)
    ("Plus(1,Power(x,-1))", 0.0, CINF),                  # finite + ComplexInfinity
    ("Times(x,Power(x,-1))", 0.0, INDET),                # 0 * ComplexInfinity
    ("Plus(Power(x,-1),Power(x,-2))", 0.0, INDET),       # ComplexInfinity + ComplexInfinity
    ("Power(E,Power(x,-1))", 0.0, INDET),                # E^ComplexInfinity
    ("Sin(Power(x,-1))", 0.0, INDET),
    ("Log(Power(x,-1))", 0.0, REAL),                     # Log(ComplexInfinity) = Infinity
    ("Abs(Power(x,-1))", 0.0, REAL),
    ("Power(x,Rational(1,2))", -4.0, CPLX),              # principal square root, exactly 2i
    ("Log(x)", -1.0, CPLX),
    ("ArcSin(x)", 2.0, CPLX),
    ("ArcCos(x)", -2.0, CPLX),
    ("Power(x,Rational(3,2))", -1.0, CPLX),
    ("Power(x,3)", -2.0, REAL),
    ("Plus(x,Indeterminate)", 1.0, INDET),
    ("unknownfn(x)", 1.0, INDET),                        # stays symbolic in Symja: never a real number
])
def test_tags(expr, x, tag):
    assert val(expr, x)[2] == tag


def test_values_that_come_back_to_the_real_line():
    assert val("Power(Plus(1,Power(x,-1)),-1)", 0.0) == (0.0, 0.0, REAL)            # 1/ComplexInfinity = 0
    assert val("Log(x)", 0.0) == (-math.inf, 0.0, REAL)                               # Log(0.0) = -Infinity
    assert val("Power(E,Log(x))", 0.0) == (0.0, 0.0, REAL)                            # E^-Infinity = 0
    re, im, tag = val("Cos(Power(x,Rational(1,2)))", -4.0)                            # cos(2i) = cosh(2)
    assert tag == REAL and re == pytest.approx(math.cosh(2.0), rel=1e-15)
    re, im, tag = val("Power(Power(x,Rational(1,2)),2)", -4.0)                        # (2i)^2 = -4
    assert (re, tag) == (-4.0, REAL)
    re, im, tag = val("Times(Power(Plus(-1,x),Rational(1,2)),Power(Plus(-2,x),Rational(1,2)))", 0.0)   # i * i*sqrt2
    assert tag == REAL and re == pytest.approx(-math.sqrt(2.0))
    re, im, tag = val("Abs(Log(x))", -1.0)                                            # |i pi|
    assert tag == REAL and re == pytest.approx(math.pi)
    re, im, tag = val("Power(E,Log(x))", -3.0)                                        # chop: e^(ln3 + i pi) = -3 (+ noise)
    assert tag == REAL and re == pytest.approx(-3.0, rel=1e-15)
    re, im, tag = val("Plus(Log(x),Times(-1,Log(Times(2,x))))", -3.0)                 # i pi - i pi
    assert tag == REAL and re == pytest.approx(-math.log(2.0))


def test_principal_values_off_the_real_domain():
    re, im, tag = val("ArcSin(x)", 2.0)
    assert re == pytest.approx(math.pi / 2) and im == pytest.approx(-math.log(2 + math.sqrt(3)))
    re, im, tag = val("ArcSin(x)", -2.0)
    assert re == pytest.approx(-math.pi / 2) and im == pytest.approx(math.log(2 + math.sqrt(3)))
    re, im, tag = val("ArcCos(x)", 2.0)
    assert re == 0.0 and im == pytest.approx(math.log(2 + math.sqrt(3)))
    re, im, tag = val("ArcCos(x)", -2.0)
    assert re == pytest.approx(math.pi) and im == pytest.approx(-math.log(2 + math.sqrt(3)))
    re, im, tag = val("Log(x)", -math.e)
    assert re == pytest.approx(1.0) and im == pytest.approx(math.pi)


def test_integer_powers_follow_the_square_and_multiply_chain():
    x = 1.0000001
    chain = x
    for bit in "0011":                 # 19 = 0b10011: MSB first, after the leading 1
        chain = chain * chain
        if bit == "1":
            chain = chain * x
    assert val("Power(x,19)", x)[0] == chain
    assert val("Power(x,-2)", 3.0)[0] == 1.0 / (3.0 * 3.0)
    assert val("Power(x,2)", 1e200)[0] == math.inf           # overflow is a directed infinity (IEEE)


def test_constant_expressions():  # eval.clj:83-96
    assert O.score_fn("Power(-2,Rational(1,2))", [1, 2, 3], [1, 2, 3], 40) == O.MIN_SCORE      # complex constant throws
    assert O.score_fn("DirectedInfinity()", [1, 2, 3], [1, 2, 3], 40) == O.MIN_SCORE
    assert O.score_fn("DirectedInfinity(1)", [1, 2, 3], [1, 2, 3], 40) == pytest.approx(-3e6 * (1 + 4e-7), rel=1e-12)
    assert O.eval_vec_pheno("Sin(Rational(1,2))", [1, 2, 3]).tolist() == [math.sin(0.5)] * 3


def test_jitter_probe_flags_chaotic_expressions_only():
    xs = np.linspace(0.1, 10, 200)
    ys = np.cos(xs)
    stable = ["Plus(x,Sin(x))", "Times(x,Power(E,Times(-1,x)))"]
    chaotic = ["Cos(Power(E,Power(x,2)))"]
    _, ok = O.score_batch_conditioned(stable + chaotic, xs, ys, 40)
    assert ok.tolist() == [True, True, False]

"""A tiny CPU emulator of the bytecode ISA (include/closyr_b200.h, v1.1 with the PUSH / SRCX modifiers), used to check
the COMPILER without a GPU: for generated trees, running the compiled program must give the value the oracle computes
from the tree itself.  The emulator works over Python complex numbers with the textbook principal branches, so the
comparison is made at points where nothing special happens (no poles, no overflow, finite values on both sides)."""
import cmath
import math

import numpy as np
import pytest

import closyr_b200 as cb
from oracle import oracle as O

END, LDX, LDC, LDZ, LDSPEC, ADDX, ADDC, MULX, MULC, NEG, RECIP, SQR, SQRT, RSQRT, POWI, POWC, EXP, LOG, SIN, COS, TAN, \
    ASIN, ACOS, ABS, ST, ADDS, MULS, POWS = range(28)
ST0, ADDS0, MULS0 = 28, 32, 36
M_SRCX, M_PUSH = 1, 2


class Skip(Exception):
    pass


def unary(op, z, b, consts):
    if op == NEG:
        return -z
    if op == RECIP:
        return 1.0 / z
    if op == SQR:
        return z * z
    if op == SQRT:
        return cmath.sqrt(z)
    if op == RSQRT:
        return 1.0 / cmath.sqrt(z)
    if op == POWI:
        n = b - 65536 if b >= 32768 else b
        return z ** n
    if op == POWC:
        return cmath.exp(consts[b] * cmath.log(z)) if z != 0 else 0j
    if op == EXP:
        return cmath.exp(z)
    if op == LOG:
        return cmath.log(z)
    if op == SIN:
        return cmath.sin(z)
    if op == COS:
        return cmath.cos(z)
    if op == TAN:
        return cmath.tan(z)
    if op in (ASIN, ACOS):
        w = cmath.asin(z) if op == ASIN else cmath.acos(z)
        # on the cuts (real |x| > 1) the algebra takes Mathematica's side, C99 the other one
        return w.conjugate() if z.imag == 0.0 and abs(z.real) > 1.0 else w
    if op == ABS:
        return complex(abs(z), 0.0)
    raise AssertionError("not an elementary function: %d" % op)


def chop(z):
    """rounding noise does not decide whether a value is real (relative 1e-10, both ways)."""
    if z.imag == 0.0 or abs(z.imag) <= 1e-10 * abs(z.real):
        return complex(z.real, 0.0)                 # also -0.0 -> +0.0: the algebra has no signed zero on the cuts
    if z.real != 0.0 and abs(z.real) <= 1e-10 * abs(z.imag):
        return complex(0.0, z.imag)
    return z


def run(code, consts, x):
    slot = {}
    acc = 0j
    for w in code:
        acc = chop(acc)
        w = int(w)
        op, a, b = w & 0xFF, (w >> 8) & 0xFF, w >> 16
        if op == END:
            return chop(acc)
        if op in (LDX, LDC) or NEG <= op <= ABS:            # the modifier byte
            if a & M_PUSH:
                slot[(a >> 2) & 3] = acc
            if (a & M_SRCX) and op >= NEG:
                acc = complex(x, 0.0)
        if op == LDX:
            acc = complex(x, 0.0)
        elif op == LDC:
            acc = complex(consts[b], 0.0)
        elif op == LDZ:
            acc = complex(consts[b], consts[b + 1])
        elif op == LDSPEC:
            raise Skip()
        elif op == ADDX:
            acc = acc + x
        elif op == ADDC:
            acc = acc + consts[b]
        elif op == MULX:
            acc = acc * x
        elif op == MULC:
            acc = acc * consts[b]
        elif NEG <= op <= ABS:
            acc = unary(op, acc, b, consts)
        elif op == ST or ST0 <= op < ADDS0:
            slot[a] = acc
        elif op == ADDS or ADDS0 <= op < MULS0:
            acc = slot[a] + acc
        elif op == MULS or op >= MULS0:
            acc = slot[a] * acc
        elif op == POWS:
            acc = cmath.exp(acc * cmath.log(slot[a])) if slot[a] != 0 else 0j
        else:
            raise AssertionError("unknown opcode %d" % op)
    raise AssertionError("program without END")


@pytest.mark.parametrize("seed,mix,max_leafs", [(11, 0, 40), (12, 1, 80), (13, 0, 12), (14, 1, 30)])
def test_compiled_programs_compute_what_the_trees_say(seed, mix, max_leafs):
    b = cb.PopulationBuilder()
    texts = b.generate(seed, 400, max_leafs, mix, want_texts=True)
    arrays = b.arrays()
    xs = np.array([0.31, 0.77, 1.18, 1.9, 2.45, 3.3, -0.6, -1.7])
    checked = 0
    for i, t in enumerate(texts):
        c0, c1 = int(arrays["code_off"][i]), int(arrays["code_off"][i + 1])
        k0, k1 = int(arrays["const_off"][i]), int(arrays["const_off"][i + 1])
        code, consts = arrays["code"][c0:c1], arrays["consts"][k0:k1]
        re, im, tag = O.eval_raw(t, xs)
        for k, x in enumerate(xs):
            if tag[k] > 1 or not (math.isfinite(re[k]) and math.isfinite(im[k])) or abs(complex(re[k], im[k])) > 1e8:
                continue
            try:
                z = run(code, consts, float(x))
            except (Skip, ZeroDivisionError, OverflowError, ValueError):
                continue
            if not (math.isfinite(z.real) and math.isfinite(z.imag)):
                continue
            want = complex(re[k], im[k] if tag[k] == 1 else 0.0)
            if abs(z - want) > 1e-7 * max(1.0, abs(want)):
                # ill-conditioned points (a pole or a branch cut next door) are not the compiler's business
                if not O.eval_discrepancy_explained(t, [x], [z.real if abs(z.imag) < 1e-9 * max(1.0, abs(z.real)) else math.inf],
                                                    O.eval_vec_pheno(t, [x]))[0]:
                    raise AssertionError("expr %d %s at x=%r: program %r, tree %r\n%s" % (
                        i, t, x, z, want, cb.disassemble(arrays, i)))
            checked += 1
    assert checked > 1000


def test_modifiers_are_what_the_listing_says():
    b = cb.PopulationBuilder()
    b.add("-1+x+Cos(x)/Sqrt(x)-x*Sin(1/2-x)")
    a = b.arrays()
    code, consts = a["code"], a["consts"]
    for x in (0.3, 1.7, 2.9):
        want = -1 + x + math.cos(x) / math.sqrt(x) - x * math.sin(0.5 - x)
        assert run(code, consts, x).real == pytest.approx(want, rel=1e-14)
        assert abs(run(code, consts, x).imag) < 1e-15

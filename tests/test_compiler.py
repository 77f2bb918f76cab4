"""Expression front-end and bytecode compiler (host side, no GPU): golden encodings, canonical forms the
reference pins as strings, leafCount, slot assignment, limits."""
import json
import os
import struct

import numpy as np
import pytest

import closyr_b200 as cb
from closyr_b200 import _lib as L
from closyr_b200 import ops
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "bytecode_golden.json")))
FORMS = json.load(open(os.path.join(ROOT, "tests", "golden", "symja_canonical_forms.json")))


def compile_one(text):
    b = cb.PopulationBuilder()
    b.add(text)
    return b.arrays()


def test_bytecode_matches_golden_bit_for_bit():
    assert len(GOLD["programs"]) >= 120
    for g in GOLD["programs"]:
        a = compile_one(g["text"])
        assert ["%08x" % int(w) for w in a["code"]] == g["code"], g["text"]
        assert [struct.pack(">d", float(c)).hex() for c in a["consts"]] == g["consts"], g["text"]
        assert "%08x" % int(a["meta"][0]) == g["meta"] and "%08x" % int(a["aux"][0]) == g["aux"], g["text"]
        assert cb.to_fullform(g["text"]) == g["fullform"]
        # the v1-equivalent instruction sequence (modifiers expanded) was frozen with ISA v1 and must not move
        assert cb.disassemble(a, 0, expand=True) == g["listing_v1"], g["text"]


def test_compiling_fullform_gives_the_same_program_as_the_printed_form():
    for g in GOLD["programs"]:
        a, b = compile_one(g["text"]), compile_one(g["fullform"])
        assert a["code"].tolist() == b["code"].tolist() and a["meta"].tolist() == b["meta"].tolist(), g["text"]
        assert a["consts"].tobytes() == b["consts"].tobytes()


def test_reference_canonical_strings_parse_and_agree_with_the_oracle_on_leaf_count():
    # test/closyr/ops_modify_test.clj:463-751: 96 canonical outputs; the oracle parses FullForm independently
    assert len(FORMS["mutations"]) == 96
    for m in FORMS["mutations"] + FORMS["extra"]:
        ff = cb.to_fullform(m["expr"])
        assert cb.leaf_count(m["expr"]) == O.leaf_count(ff) > 0
        assert cb.to_fullform(ff) == ff          # FullForm is a fixed point


@pytest.mark.parametrize("text,fullform,leafs", [
    ("x*x-1/2", "Plus(Rational(-1,2),Power(x,2))", 7),          # ops_test.clj:18-28 => -3.0000147 needs leafs 7
    ("x-1/2", "Plus(Rational(-1,2),x)", 5),
    ("x*1", "x", 1),                                             # ops_eval_test.clj:82-88
    ("Sqrt(x-1)", "Power(Plus(-1,x),Rational(1,2))", 7),         # ops_eval_test.clj:35 "Sqrt(-1+x)"
    ("E-1/2", "Plus(Rational(-1,2),E)", 5),                      # ops_eval_test.clj:29 "-1/2+E"
    ("x-Infinity", "DirectedInfinity(-1)", 2),                   # ops_eval_test.clj:168-173
    ("-1+1", "0", 1),
    ("1/2+1/10", "Rational(3,5)", 3),
    ("Cos(x)/(2*x^(3/2))", "Times(Rational(1,2),Cos(x),Power(x,Rational(-3,2)))", 11),
    ("Exp(x)", "Power(E,x)", 3),
    ("1/Sin(x)", "Power(Sin(x),-1)", 4),
])
def test_canonicalisation(text, fullform, leafs):
    assert cb.to_fullform(text) == fullform
    assert cb.leaf_count(text) == leafs


def test_listing_of_the_mutation_test_expression():
    a = compile_one("-1+x+Cos(x)/Sqrt(x)-x*Sin(1/2-x)")
    # the v1 instruction sequence the program stands for ...
    assert cb.disassemble(a, 0, expand=True) == [
        "LDC -1.0", "ADDX", "ST0", "LDX", "COS", "ST1", "LDX", "RSQRT", "MULS1", "ADDS0", "ST0",
        "LDC -1.0", "MULX", "ST1", "LDC -1.0", "MULX", "ADDC 0.5", "SIN", "MULS1", "ADDS0", "END"]
    # ... and the 14 instructions it is encoded in (ISA v1.1: stores fused into the next load, LDX into the function)
    assert cb.disassemble(a, 0) == [
        "LDC -1.0", "ADDX", "push 0: COS(x)", "push 1: RSQRT(x)", "MULS1", "ADDS0", "push 0: LDC -1.0", "MULX",
        "push 1: LDC -1.0", "MULX", "ADDC 0.5", "SIN", "MULS1", "ADDS0", "END"]
    assert [int(w) for w in a["code"][:4]] == [0x00000002, 0x00000005, 0x00000313, 0x0000070D]
    meta = int(a["meta"][0])
    assert meta & 0xFFFF == 22 and (meta >> 24) == 2 and ((meta >> 16) & 0xFF) == 0


def test_programs_are_padded_for_bulk_copies():
    b = cb.PopulationBuilder()
    texts = b.generate(3, 500, 40, 0, want_texts=True)
    a = b.arrays()
    assert (a["code_off"] % 4 == 0).all() and (a["const_off"] % 2 == 0).all()
    assert len(texts) == 500 and b.n_expr == 500
    ends = a["code"][a["code_off"][1:] - 1]
    assert (ends & 0xFF == 0).all()              # every program ends with END (padding is END too)


def test_flags_const_nil_and_quirk_table():
    a = compile_one("E-1/2")
    meta, aux = int(a["meta"][0]), int(a["aux"][0])
    assert (meta >> 16) & L.F_CONST
    argc, off = aux >> 16, aux & 0xFFFF
    assert argc == 2
    table = a["consts"][off:off + argc]
    assert table[0] == -0.5 and np.isinf(table[1])          # eval.clj:109: Rational is a real literal, E is not
    assert (int(compile_one("Indeterminate")["meta"][0]) >> 16) & L.F_NIL
    assert not (int(compile_one("x+1")["meta"][0]) >> 16) & L.F_CONST


def test_slot_order_prefers_the_deeper_operand_first():
    # of two leading non-leaf operands the one needing more slots is evaluated first (same bits: + and * commute)
    a = compile_one("Times(Sin(x),Plus(Cos(x),Times(Sin(x),Plus(Cos(x),Times(Sin(x),Cos(x))))))")
    assert (int(a["meta"][0]) >> 24) <= 2


def test_deep_expressions_use_generic_slots_and_limits_flag_unsupported():
    deep = "x"
    for _ in range(12):
        deep = "Plus(Sin(x),Cos(x),Times(Sin(x),Cos(x),%s))" % deep
    a = compile_one(deep)
    ops_used = set(int(w) & 0xFF for w in a["code"])
    assert 24 in ops_used or (int(a["meta"][0]) >> 24) <= 4       # CL_ST generic once past 4 slots
    huge = "Plus(" + ",".join("Sin(Times(%d,x))" % i for i in range(3000)) + ")"
    assert (int(compile_one(huge)["meta"][0]) >> 16) & L.F_UNSUPPORTED


def test_parse_errors_are_reported_not_swallowed():
    b = cb.PopulationBuilder()
    with pytest.raises(cb.ClosyrError) as ei:
        b.add("Plus(((")
    assert ei.value.code == L.E_PARSE and b.n_expr == 0
    with pytest.raises(cb.ClosyrError):
        b.add_many(["x", "Sin(", "x"])


def test_to_phenotype_defaults():  # common.clj:127-132; ops_common_test.clj:55-67
    for bad in (None, "", "Indeterminate", "Infinity", "Sin", "Times", "E"):
        assert ops.to_phenotype(bad)["expr"] == "x"
    assert ops.to_phenotype("x*x-1/2")["expr"] == "Plus(Rational(-1,2),Power(x,2))"


def test_generator_is_deterministic_and_respects_max_leafs():
    b1, b2 = cb.PopulationBuilder(), cb.PopulationBuilder()
    t1 = b1.generate(888, 2000, 40, 0, want_texts=True)
    t2 = b2.generate(888, 2000, 40, 0, want_texts=True)
    assert t1 == t2 and b1.arrays()["code"].tolist() == b2.arrays()["code"].tolist()
    leafs = b1.arrays()["meta"] & 0xFFFF
    assert leafs.max() <= 40 and leafs.min() >= 1
    assert [O.leaf_count(t) for t in t1[:200]] == leafs[:200].tolist()
    b3 = cb.PopulationBuilder()
    b3.add_many(t1)
    assert b3.arrays()["code"].tolist() == b1.arrays()["code"].tolist()      # text -> bytecode round trip
    heavy = cb.PopulationBuilder().generate(888, 300, 80, 1, want_texts=True)
    assert max(O.leaf_count(t) for t in heavy) <= 80


def test_extend_xs_matches_reference_formula():  # common.clj:262-291
    ext = ops.extend_xs([0.0, 1.0, 2.0, 3.0])
    assert len(ext["x-head"]) == len(ext["x-tail"]) == 2
    assert ext["x-tail"] == pytest.approx([3.75, 4.5]) and ext["x-head"] == pytest.approx([-1.5, -0.75])
    assert ext["xs"] == ext["x-head"] + [0.0, 1.0, 2.0, 3.0] + ext["x-tail"]

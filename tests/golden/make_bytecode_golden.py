"""Writes tests/golden/bytecode_golden.json: the exact instruction words, constant pools and metadata the
compiler emits for a fixed corpus.  The Clojure mirror (clj/closyr/gpu/compile.clj) must reproduce these
bit for bit; the C++ compiler is checked against the file by tests/test_compiler.py."""
import json
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import closyr_b200 as cb  # noqa: E402

forms = json.load(open(os.path.join(ROOT, "tests", "golden", "symja_canonical_forms.json")))
corpus = [m["expr"] for m in forms["mutations"]] + [m["expr"] for m in forms["extra"]] + [
    "x", "1", "-1/2", "0.123", "E", "Pi", "x*x-1/2", "E-1/2", "Sqrt(x-1)", "x*Sin(x)-1/2", "Infinity", "ComplexInfinity",
    "Indeterminate", "I", "2^x", "x^x", "x^2.5", "x^40", "x^-3", "Abs(x)", "unknownfn(x)", "Tan(x)+Cot(x)+Sec(x)+Csc(x)",
    "ArcSin(x)*ArcCos(x)", "Plus(Times(Sin(x),Cos(x)),Times(Log(x),Power(E,x)),Times(Sin(x),Log(x)))",
]
out = []
for text in corpus:
    b = cb.PopulationBuilder()
    b.add(text)
    a = b.arrays()
    out.append({
        "text": text,
        "fullform": cb.to_fullform(text),
        "leafs": cb.leaf_count(text),
        "code": ["%08x" % int(w) for w in a["code"]],
        "consts": [struct.pack(">d", float(c)).hex() for c in a["consts"]],
        "meta": "%08x" % int(a["meta"][0]),
        "aux": "%08x" % int(a["aux"][0]),
        "listing": cb.disassemble(a, 0),
        "listing_v1": cb.disassemble(a, 0, expand=True),
    })
json.dump({"isa": "1.1", "programs": out}, open(os.path.join(ROOT, "tests", "golden", "bytecode_golden.json"), "w"), indent=1)
print(len(out), "programs")

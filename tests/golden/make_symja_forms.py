"""Extracts the canonical Symja output strings the reference's own tests pin
(/root/reference/test/closyr/ops_modify_test.clj:463-751, `all-mods-applied-on-fn-expected`: every one of the
96 mutations applied to x+Cos(x)/Sqrt(x)+x*Sin(x-1/2)-1) into tests/golden/symja_canonical_forms.json.
Run in the build container only (the GPU box has no /root/reference); the JSON is committed."""
import json
import re

SRC = "/root/reference/test/closyr/ops_modify_test.clj"
text = open(SRC).read()
start = text.index("(def all-mods-applied-on-fn-expected")
end = text.index("(deftest mutations-test")
block = text[start:end]
rows = re.findall(r'\[:([a-z-]+)\s+"([^"]*)"\s+"([^"]*)"\]', block)
assert len(rows) == 96, len(rows)
extra = [  # other canonical strings asserted by the reference's tests
    ("ops_test.clj:150", "-1+x"), ("ops_test.clj:168", "-1+x^2"), ("ops_test.clj:186", "x^2+Sin(x)"),
    ("ops_eval_test.clj:29", "-1/2+E"), ("ops_eval_test.clj:35", "Sqrt(-1+x)"),
    ("README/screenshot", "ArcSin(Sin(E^x))+2*(2*(1/2+0.9900000000000001*(1/10-2*Cos(0.9*(1/20+x)))+Cos(x))+2*Sin(x))"),
]
out = {"source": "test/closyr/ops_modify_test.clj:463-751 (+ strings from ops_test.clj / ops_eval_test.clj)",
       "mutations": [{"op": a, "label": b, "expr": c} for a, b, c in rows],
       "extra": [{"where": w, "expr": e} for w, e in extra]}
json.dump(out, open("/root/repo/tests/golden/symja_canonical_forms.json", "w"), indent=1)
print(len(rows), "mutation strings written")

"""Accuracy check of the lean exp/log/sin/cos against numpy (correctly rounded enough) through eval_batch."""
import sys
import numpy as np
sys.path.insert(0, ".")
import closyr_b200 as cb

rng = np.random.default_rng(1)
ctx = cb.Context([0])
b = cb.PopulationBuilder()
exprs = ["Power(E,x)", "Log(x)", "Sin(x)", "Cos(x)"]
b.add_many(exprs)
sets = {
    "small": rng.uniform(-3, 3, 200000),
    "wide": rng.uniform(-700, 700, 200000),
    "pos": np.exp(rng.uniform(-700, 700, 200000)),
    "trig": rng.uniform(-1e6, 1e6, 200000),
    "edge": np.array([0.0, -0.0, 1.0, -1.0, 1e-310, 5e-324, 699.9, 700.0, 709.7, 710.0, -745.0, -746.0, 1e8, 1e9, 1e300, np.inf, -np.inf, np.nan, np.pi/2, np.pi, 1e-20]),
}
refs = [np.exp, np.log, np.sin, np.cos]
for name, xs in sets.items():
    f, st = ctx.eval_batch(b.view(), xs)
    for i, (e, ref) in enumerate(zip(exprs, refs)):
        with np.errstate(all="ignore"):
            want = ref(xs.astype(np.longdouble)).astype(np.float64) if name != "edge" else ref(xs)
            got = f[i]
            fin = np.isfinite(want) & np.isfinite(got)
            ulp = np.abs(got[fin] - want[fin]) / np.maximum(np.spacing(np.abs(want[fin])), 5e-324)
            cls_ok = (np.isfinite(want) == np.isfinite(got)) | (np.isnan(want) & np.isinf(got))
        print("%-6s %-12s max ulp %.2f  mean ulp %.3f  finite-class mismatches %d" % (name, e, ulp.max() if ulp.size else 0, ulp.mean() if ulp.size else 0, (~cls_ok).sum()))
        if name == "edge":
            print("   got ", got.tolist())
            print("   want", want.tolist())

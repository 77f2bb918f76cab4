"""Debug helper: list the points where eval_batch on the GPU and the oracle disagree (conditioned)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import closyr_b200 as cb
from oracle import oracle as O

def synth(n):
    xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
    return xs, 2.0 + 1.5 * xs / np.pi + np.cos(xs)

ctx = cb.Context([0])
def run(seed, count, ml, mix, xs, label):
    b = cb.PopulationBuilder()
    texts = b.generate(seed, count, ml, mix, want_texts=True)
    f, st = ctx.eval_batch(b.view(), xs)
    nbad = 0
    for i, t in enumerate(texts):
        try:
            w, ok = O.eval_conditioned(t, xs)
        except ArithmeticError:
            if st[i] != 2: print(label, "status mismatch", i, t)
            continue
        if w is None:
            continue
        with np.errstate(invalid="ignore"):
            same = (f[i] == w) | (np.abs(f[i] - w) <= 1e-9 * np.abs(w))
        bad = np.nonzero(ok & ~same)[0]
        if len(bad):
            nbad += 1
            print(label, "expr", i, t)
            for k in bad[:4]:
                re, im, tg = O.eval_raw(t, [xs[k]])
                print("    pt %d x=%r gpu=%r oracle=%r raw=(%r,%r,%d)" % (k, xs[k], f[i][k], w[k], re[0], im[0], tg[0]))
    print(label, "expressions with mismatches:", nbad, "of", count)

xs, _ = synth(300)
run(4242, 200, 40, 0, np.concatenate([[-2.0, -0.5, 0.0], xs]), "evaltest")
xs, _ = synth(4099)
run(888 + 4099, 400, 80, 1, xs, "score4099")

import sys
import numpy as np
sys.path.insert(0, ".")
import closyr_b200 as cb
from oracle import oracle as O
t = "Times(Plus(Times(Rational(1,10),Plus(Rational(2,5),ArcSin(x))),Times(-1,Sin(x))),x)"
for n in (64, 128, 256, 1024):
    xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
    ys = 2.0 + 1.5 * xs / np.pi + np.cos(xs)
    ctx = cb.Context([0])
    ctx.set_points(xs, ys)
    for texts in ([t], ["ArcSin(x)", "Plus(Rational(2,5),ArcSin(x))", "Plus(ArcSin(x),Sin(x))", "Plus(ArcSin(x),Times(-1,Sin(x)))", t, "Times(ArcSin(x),x)"]):
        b = cb.PopulationBuilder(); b.add_many(texts)
        s, f = ctx.score_batch(b.view(), 40)
        w, _ = O.score_batch(texts, xs, ys, 40)
        print(n, "score", s.tolist(), "oracle", w.tolist(), "flags", f.tolist())
    ctx.close()

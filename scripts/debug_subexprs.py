"""Debug helper (GPU box): for one expression, evaluate every sub-tree on the GPU and with the oracle and print the
innermost sub-trees whose values differ, with the points."""
import sys
import numpy as np
sys.path.insert(0, ".")
import closyr_b200 as cb
from oracle import oracle as O


def subtrees(t):
    """every Head(...) substring of a FullForm text, innermost first"""
    out, stack = [], []
    i = 0
    while i < len(t):
        ch = t[i]
        if ch == "(":
            j = i - 1
            while j >= 0 and (t[j].isalnum() or t[j] == "_"):
                j -= 1
            stack.append(j + 1)
        elif ch == ")":
            s = stack.pop()
            out.append(t[s:i + 1])
        i += 1
    seen, uniq = set(), []
    for s in out:
        if s not in seen:
            seen.add(s)
            uniq.append(s)
    return uniq


def main():
    expr = sys.argv[1]
    if len(sys.argv) > 2:
        xs = np.array([float(v) for v in sys.argv[2].split(",")])
    else:
        n = 300
        xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
        xs = np.concatenate([[-2.0, -0.5, 0.0], xs])
    ctx = cb.Context([0])
    subs = [s for s in subtrees(cb.to_fullform(expr)) if "x" in s]
    b = cb.PopulationBuilder()
    b.add_many(subs)
    f, st = ctx.eval_batch(b.view(), xs)
    shown = 0
    for i, s in enumerate(subs):
        try:
            w, ok = O.eval_conditioned(s, xs)
        except ArithmeticError:
            continue
        if w is None:
            continue
        with np.errstate(invalid="ignore"):
            same = (f[i] == w) | (np.abs(f[i] - w) <= 1e-9 * np.abs(w))
        bad = np.nonzero(ok & ~same)[0]
        explained, nb = O.eval_discrepancy_explained(s, xs, f[i], w)
        if not explained:
            print("UNEXPLAINED in", s)
            for k in np.nonzero(~same)[0][:40]:
                e1, _ = O.eval_discrepancy_explained(s, xs[k:k + 1], f[i][k:k + 1], w[k:k + 1])
                if not e1:
                    re, im, tg = O.eval_raw(s, [xs[k]])
                    print("    pt %d x=%r gpu=%r oracle=%r ok=%s raw=(%r,%r,%d)" % (k, xs[k], f[i][k], w[k], ok[k], re[0], im[0], tg[0]))
            shown += 1
        if len(bad):
            print("MISMATCH in", s)
            for k in bad[:6]:
                re, im, tg = O.eval_raw(s, [xs[k]])
                print("    pt %d x=%r gpu=%r oracle=%r raw=(%r,%r,%d)" % (k, xs[k], f[i][k], w[k], re[0], im[0], tg[0]))
            shown += 1
            if shown >= 4:
                break
    print("sub-trees checked:", len(subs), "with mismatches shown:", shown)


if __name__ == "__main__":
    main()

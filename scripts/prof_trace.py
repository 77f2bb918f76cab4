"""Per-CTA phase times of score_kernel from the instrumented build (scripts/build_trace.sh).
CLOSYR_B200_LIB=closyr_b200/libclosyr_b200_trace.so python scripts/prof_trace.py [pop] [points]"""
import ctypes as C
import os
import sys
import numpy as np
sys.path.insert(0, ".")
os.environ.setdefault("CLOSYR_B200_LIB", os.path.abspath("closyr_b200/libclosyr_b200_trace.so"))
import closyr_b200 as cb
from closyr_b200 import _lib as L

pop = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ml = 40
xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
ys = 2.0 + 1.5 * xs / np.pi + np.cos(xs)
b = cb.PopulationBuilder()
which = sys.argv[3] if len(sys.argv) > 3 else "mix"
texts = b.generate(888, pop, ml, 0, want_texts=True)
if which != "mix":
    # copies of ONE expression: "fast" / "full" = the median-length one of that kind (kinds from a first scoring)
    c0 = cb.Context([0]); c0.set_points(xs, ys); c0.upload_population(b.view()); c0.score_resident(ml)
    _, f0 = c0.download_scores()
    a = b.arrays(); length = np.diff(a["code_off"].astype(np.int64))
    kind = (np.asarray(f0) & 4) != 0
    idx = np.flatnonzero(kind == (which == "full"))
    one = texts[int(idx[np.argsort(length[idx])[len(idx) // 2]])]
    print("copies of", one)
    b = cb.PopulationBuilder(); b.add_many([one] * pop)
ctx = cb.Context([0])
ctx.set_points(xs, ys)
ctx.upload_population(b.view())
for _ in range(3):
    ms = ctx.score_resident(ml)
s, f = ctx.download_scores()
S = 16
items = min(pop, 32768)
buf = np.zeros(items * S, dtype=np.uint64)
lib = L.lib()
lib.closyr_debug_trace.argtypes = [C.c_void_p, C.c_int]
rc = lib.closyr_debug_trace(buf.ctypes.data, items)
assert rc == 0, rc
t = buf.reshape(items, S).astype(np.int64)
GHZ = 1.965
full = (np.asarray(f[:items]) & 4) != 0
print("kernel %.3f ms for %d x %d; %d traced CTAs, %.1f %% use the complete algebra" % (ms, pop, n, items, 100 * full.mean()))
names = ["entry->desc", "desc->tma done", "tma->fast pass done (warp 0)", "fast->barrier", "barrier->slow half done", "->loop end", "->final barrier", "->emit"]
def us(c): return c / GHZ / 1e3
def row(label, d):
    d = d[np.isfinite(d)]
    if len(d) == 0: return
    print("  %-34s mean %7.2f us  p50 %7.2f  p90 %7.2f  max %8.2f" % (label, us(d.mean()), us(np.median(d)), us(np.percentile(d, 90)), us(d.max())))
for sel, lab in ((~full, "fast-only expressions"), (full, "complete-algebra expressions")):
    tt = t[sel]
    if len(tt) == 0: continue
    print(lab, len(tt))
    row(names[0], (tt[:, 2] - tt[:, 1]).astype(float))
    row(names[1], (tt[:, 3] - tt[:, 2]).astype(float))
    row(names[2], (tt[:, 4] - tt[:, 3]).astype(float))
    row(names[3], (tt[:, 5] - tt[:, 4]).astype(float))
    slow = np.where(tt[:, 6] > tt[:, 5], tt[:, 6] - tt[:, 5], 0).astype(float)
    row(names[4], slow)
    row("loop end -> final barrier", (tt[:, 8] - tt[:, 7]).astype(float))
    row("final barrier -> emit", (tt[:, 9] - tt[:, 8]).astype(float))
    row("whole CTA (entry -> emit)", (tt[:, 9] - tt[:, 1]).astype(float))
    if n >= 4 * 4096:
        row("round 0 (tma done -> round end)", (tt[:, 12] - tt[:, 3]).astype(float))
        for r in (1, 2, 3):
            row("round %d" % r, (tt[:, 12 + r] - tt[:, 11 + r]).astype(float))
# gaps between consecutive CTAs of the same SM (same clock)
gaps = []
life = []
for sm in np.unique(t[:, 0]):
    q = t[t[:, 0] == sm]
    q = q[np.argsort(q[:, 1])]
    if len(q) < 3: continue
    g = q[1:, 1] - q[:-1, 9]
    gaps.append(g[(g > -10_000_000) & (g < 10_000_000)])
    life.append(q[:, 9] - q[:, 1])
gaps = np.concatenate(gaps).astype(float)
life = np.concatenate(life).astype(float)
row("gap: emit of a CTA -> entry of the next on the SM", gaps)
print("  share of the SM's time: in CTAs %.1f %%, between CTAs %.1f %%" % (100 * life.sum() / (life.sum() + gaps.sum()), 100 * gaps.sum() / (life.sum() + gaps.sum())))

if n == 4096:
    tt = t[full]
    tot = tt[:, 11]
    print("complete-algebra round of a 4096-point CTA by the number of reported points (64 per warp task, 32 warps):")
    edges = [0, 1, 2, 65, 513, 1025, 2049, 3073, 4095, 4097]
    for lo, hi in zip(edges[:-1], edges[1:]):
        m = (tot >= lo) & (tot < hi)
        if not m.any(): continue
        q = tt[m]
        print("  reported %4d..%4d: %5d CTAs (%4.1f %%)  compaction %5.2f us  tasks of warp 0 %6.2f us  wait for the others %6.2f us  rest of the slow half %5.2f us  fast pass %5.2f us  share of all SM time %4.1f %%"
              % (lo, hi - 1, m.sum(), 100 * m.sum() / len(t), us((q[:, 13] - q[:, 5]).mean()), us((q[:, 14] - q[:, 13]).mean()), us((q[:, 15] - q[:, 14]).mean()),
                 us((q[:, 6] - q[:, 15]).mean()), us((q[:, 4] - q[:, 3]).mean()), 100 * (q[:, 9] - q[:, 1]).sum() / (t[:, 9] - t[:, 1]).sum()))

"""Alternative grouping keys for the CTA order of short tiles, applied on the host (CLOSYR_ORDER=0 so that the library keeps
the order it is given)."""
import os, sys
os.environ["CLOSYR_ORDER"] = "0"
import numpy as np
sys.path.insert(0, ".")
import closyr_b200 as cb
pop, ml = 100000, 40
b = cb.PopulationBuilder()
texts = b.generate(888, pop, ml, 0, want_texts=True)
a = b.arrays()
code, off = a["code"], a["code_off"].astype(np.int64)
ops = (code & 0xFF).astype(np.int64)
srcx = ((code >> 8) & 1).astype(bool) & (ops >= 10) & (ops <= 22)
ctx = cb.Context([0])
def timeit(order, label, ns=(256, 1024)):
    tx = [texts[i] for i in order]
    bb = cb.PopulationBuilder(); bb.add_many(tx)
    out = []
    for n in ns:
        xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
        ys = 2.0 + 1.5 * xs / np.pi + np.cos(xs)
        ctx.set_points(xs, ys); ctx.upload_population(bb.view())
        for _ in range(3): ms = ctx.score_resident(ml)
        out.append(ms)
    print("%-52s %s" % (label, "  ".join("%7.3f ms" % m for m in out)), flush=True)
timeit(np.arange(pop), "generator order")
# the library's key, rebuilt here: functions that run a routine (not the tabulated f(x)), heaviest first, + length class
fn_bits = {15: 10, 27: 10, 17: 9, 12: 8, 13: 8, 21: 7, 22: 6, 20: 5, 16: 4, 18: 3, 19: 3, 10: 2, 14: 2}
total_tab = {16, 18, 19, 20}          # exp sin cos tan of the raw variable come from the tables
key = np.zeros(pop, np.int64); fset = np.zeros(pop, np.int64)
for e in range(pop):
    o = ops[off[e]:off[e + 1]]; sx = srcx[off[e]:off[e + 1]]
    k = 0; f = 0
    for op, s in zip(o, sx):
        if op in fn_bits and not (s and op in total_tab): k |= 1 << fn_bits[op]
        if 9 <= op <= 23 and not (s and op in total_tab): f |= 1 << (op - 9)
    key[e] = k | min(3, len(o) // 10); fset[e] = f
timeit(np.argsort(key, kind="stable"), "library key (12 bits)")
timeit(np.lexsort((np.diff(off), fset)), "exact function set, then length")
timeit(np.lexsort((fset, np.diff(off))), "length, then exact function set")
seq = [tuple(int(x) for x in ops[off[e]:off[e + 1]] if 9 <= x <= 23 or x in (15, 27)) for e in range(pop)]
timeit(sorted(range(pop), key=lambda e: seq[e]), "sequence of function opcodes, lexicographic")
timeit(sorted(range(pop), key=lambda e: (key[e] >> 2, seq[e])), "library key, then function sequence")
timeit(sorted(range(pop), key=lambda e: texts[e]), "text, lexicographic")
nfun = np.array([bin(int(f)).count("1") for f in fset])
timeit(np.lexsort((fset, nfun)), "number of distinct functions, then set")

"""Does the ORDER of the expressions in a shard matter?  Times the same population in generator order, sorted by the
set of opcodes it uses, sorted by (result flag, opcode set, length), shuffled, and as copies of one expression."""
import sys
import numpy as np
sys.path.insert(0, ".")
import closyr_b200 as cb

pop = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ml = int(sys.argv[3]) if len(sys.argv) > 3 else 40
mix = int(sys.argv[4]) if len(sys.argv) > 4 else 0
xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
ys = 2.0 + 1.5 * xs / np.pi + np.cos(xs)
b = cb.PopulationBuilder()
texts = b.generate(888, pop, ml, mix, want_texts=True)
a = b.arrays()
code, off = a["code"], a["code_off"]
ops = (code & 0xFF).astype(np.uint64)
bits = np.left_shift(np.uint64(1), ops)
mask = np.bitwise_or.reduceat(bits, off[:-1].astype(np.int64))
length = np.diff(off.astype(np.int64))
ctx = cb.Context([0])
ctx.set_points(xs, ys)

def timeit(texts, label):
    bb = cb.PopulationBuilder()
    bb.add_many(texts)
    ctx.upload_population(bb.view())
    for _ in range(3):
        ms = ctx.score_resident(ml)
    s, f = ctx.download_scores()
    print("%-40s %7d expr: %8.3f ms  %.3e expr-pts/s" % (label, len(texts), ms, len(texts) * n / ms * 1e3), flush=True)
    return s, f

s, f = timeit(texts, "generator order")
full = (np.asarray(f) & 4) != 0
heavy = mask & ~np.uint64((1 << 12) - 1 | (0xFFFF << 24))      # RECIP.. ABS bits: the function handlers
o = np.lexsort((length, mask))
timeit([texts[i] for i in o], "sorted by opcode set")
o = np.lexsort((length, heavy))
timeit([texts[i] for i in o], "sorted by function-opcode set")
o = np.lexsort((length, heavy, full))
timeit([texts[i] for i in o], "sorted by (full, function set, len)")
o = np.lexsort((length, full))
timeit([texts[i] for i in o], "sorted by (full, len)")
o = np.argsort(length, kind="stable")
timeit([texts[i] for i in o], "sorted by length")
rng = np.random.default_rng(1)
o = rng.permutation(pop)
timeit([texts[i] for i in o], "shuffled")
med = int(np.argsort(length)[pop // 2])
timeit([texts[med]] * pop, "copies of the median-length expr")
fi = np.flatnonzero(full)
if len(fi):
    m2 = int(fi[np.argsort(length[fi])[len(fi) // 2]])
    timeit([texts[m2]] * pop, "copies of a median full-algebra expr")

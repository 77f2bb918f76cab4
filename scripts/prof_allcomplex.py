"""Which expressions leave the real line at (almost) every point?  (instrumented build: reported points of the first round)"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ.setdefault("CLOSYR_B200_LIB", os.path.abspath("closyr_b200/libclosyr_b200_trace.so"))
import closyr_b200 as cb
from closyr_b200 import _lib as L
pop, n, ml = 20000, 4096, 40
xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
ys = 2.0 + 1.5 * xs / np.pi + np.cos(xs)
b = cb.PopulationBuilder(); texts = b.generate(888, pop, ml, 0, want_texts=True)
ctx = cb.Context([0]); ctx.set_points(xs, ys); ctx.upload_population(b.view()); ctx.score_resident(ml)
S = 16
buf = np.zeros(pop * S, dtype=np.uint64)
lib = L.lib(); lib.closyr_debug_trace.argtypes = [C.c_void_p, C.c_int]
assert lib.closyr_debug_trace(buf.ctypes.data, pop) == 0
t = buf.reshape(pop, S).astype(np.int64)
s, f = ctx.download_scores()
full = (np.asarray(f) & 4) != 0
tot = np.where(full, t[:, 11], 0)
idx = np.flatnonzero(tot >= 3073)
rng = np.random.default_rng(3)
with open("gpurun_out/allcomplex.txt", "w") as fo:
    fo.write("%d of %d expressions report >= 3073 of 4096 points\n" % (len(idx), pop))
    for i in idx:
        fo.write("%d\t%d\t%s\n" % (i, tot[i], texts[i]))
np.save("gpurun_out/reported_first_round.npy", np.stack([tot, full.astype(np.int64), t[:, 9] - t[:, 1]], 1))
print(len(idx))

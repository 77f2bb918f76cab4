"""Tags of the two points a lane hands to a complex function call in run_cplx (instrumented build)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, ".")
os.environ.setdefault("CLOSYR_B200_LIB", os.path.abspath("closyr_b200/libclosyr_b200_trace.so"))
import closyr_b200 as cb
from closyr_b200 import _lib as L
pop, n, ml, mix = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
ys = 2.0 + 1.5 * xs / np.pi + np.cos(xs)
b = cb.PopulationBuilder(); b.generate(888, pop, ml, mix)
ctx = cb.Context([0]); ctx.set_points(xs, ys); ctx.upload_population(b.view()); ctx.score_resident(ml)
buf = np.zeros(8, dtype=np.uint64)
lib = L.lib(); lib.closyr_debug_ztags.argtypes = [C.c_void_p]
assert lib.closyr_debug_ztags(buf.ctypes.data) == 0
c = buf.astype(float)
print("%d x %d mix %d: %d warp-level complex function calls: every lane has two complex points %.1f %%, no lane has %.1f %%, mixed %.1f %%;"
      " lanes: both complex %.1f %%, neither %.1f %% (both real %.1f %%), one of the two %.1f %%"
      % (pop, n, mix, c[0], 100 * c[1] / c[0], 100 * c[2] / c[0], 100 * c[3] / c[0], 100 * c[4] / (32 * c[0]), 100 * c[5] / (32 * c[0]),
         100 * c[6] / (32 * c[0]), 100 * (32 * c[0] - c[4] - c[5]) / (32 * c[0])))

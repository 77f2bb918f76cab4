"""Where the end-to-end time of closyr_score_batch goes (configs[2]): wall times of the three-step resident form."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import closyr_b200 as cb
pop, n, ml = 100000, 4096, 40
xs = (np.pi / 15.0) * 49.0 * np.arange(n) / max(1, n - 1)
ys = 2.0 + 1.5 * xs / np.pi + np.cos(xs)
b = cb.PopulationBuilder(pinned=True); b.generate(888, pop, ml, 0)
v = b.view()
ctx = cb.Context([0]); ctx.set_points(xs, ys)
for _ in range(3): ctx.score_batch(v, ml)
T = {"upload": 0.0, "score": 0.0, "download": 0.0, "batch": 0.0, "dev": 0.0}
R = 10
s = np.empty(pop); f = np.empty(pop, np.uint8)
for _ in range(R):
    t0 = time.perf_counter(); ctx.upload_population(v); t1 = time.perf_counter()
    T["dev"] += ctx.score_resident(ml); t2 = time.perf_counter()
    ctx.download_scores(s, f); t3 = time.perf_counter()
    ctx.score_batch(v, ml); t4 = time.perf_counter()
    T["upload"] += t1 - t0; T["score"] += t2 - t1; T["download"] += t3 - t2; T["batch"] += t4 - t3
print({k: round(1e3 * x / R, 3) if k != "dev" else round(x / R, 3) for k, x in T.items()})

"""Per-function view of an ncu source-page CSV of score_kernel: executed warp instructions, FP64 share and stall samples
per device function (symbol ranges from the cubin).
   python scripts/ncu_byfunc.py src.csv kernels.sm_100a.cubin [kernel-substring]"""
import csv, subprocess, sys, re
rows = list(csv.reader(open(sys.argv[1])))
cubin = sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "score_kernel"
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
syms = []
out = subprocess.run(["readelf", "-sW", cubin], capture_output=True, text=True).stdout
for line in out.splitlines():
    f = line.split()
    if len(f) >= 8 and f[3] == "FUNC" and kern in f[7]:
        size = int(f[2], 0)
        name = f[7]
        m = re.search(r"\$_ZN6closyr3dev\d+([a-zA-Z_0-9]+?)(E|I)", name) or re.search(r"closyr43_GLOBAL__N__[0-9a-f_]+kernels_cu_[0-9a-f]+\d+([a-z_]+)E[A-Z]", name.split("$")[-1]) 
        short = name.split("$")[-1]
        short = re.sub(r"_ZN6closyr3dev\d+", "", short)
        short = re.sub(r"_ZN41_INTERNAL_[0-9a-f_]+kernels_cu_[0-9a-f]+6closyr43_GLOBAL__N__[0-9a-f_]+kernels_cu_[0-9a-f]+\d+", "", short)
        short = re.sub(r"_ZN6closyr43_GLOBAL__N__[0-9a-f_]+kernels_cu_[0-9a-f]+\d+", "", short)
        syms.append((int(f[1], 16), size, short[:40]))
syms.sort()
data = []
for r in rows[2:]:
    if len(r) <= iex: continue
    data.append((int(r[ia], 16), r[isrc].strip(), int(r[iex] or 0), int(r[ismp] or 0), r))
base = data[0][0]
def func_of(off):
    best = None
    for a, s, n in syms:
        if a <= off < a + s and (best is None or s < best[1]): best = (a, s, n)
    return best[2] if best else "?"
agg = {}
for a, s, e, smp, r in data:
    fn = func_of(a - base)
    d = agg.setdefault(fn, {"ex": 0, "smp": 0, "fp64": 0, "n": 0, "st": {}})
    d["ex"] += e; d["smp"] += smp; d["n"] += 1
    op = s.split()[0] if not s.startswith("@") else s.split()[1]
    if op.startswith(("DADD", "DMUL", "DFMA", "DSETP", "DMNMX")): d["fp64"] += e
    for i, h in stall_cols:
        v = int(r[i] or 0)
        if v: d["st"][h] = d["st"].get(h, 0) + v
tot_ex = sum(d["ex"] for d in agg.values()); tot_smp = sum(d["smp"] for d in agg.values()); tot64 = sum(d["fp64"] for d in agg.values())
print("total %.3e warp-instr, FP64 share %.1f %%, samples %d" % (tot_ex, 100.0 * tot64 / tot_ex, tot_smp))
for fn, d in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:40]:
    st = sorted(d["st"].items(), key=lambda kv: -kv[1])[:4]
    print("%-42s n=%5d instr%%=%5.2f fp64%%=%5.1f smp%%=%5.2f  %s" % (fn, d["n"], 100.0 * d["ex"] / tot_ex, 100.0 * d["fp64"] / max(1, d["ex"]), 100.0 * d["smp"] / tot_smp,
          " ".join("%s=%.1f" % (k.replace("stall_", ""), 100.0 * v / tot_smp) for k, v in st)))

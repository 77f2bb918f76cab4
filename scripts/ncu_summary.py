"""Summarise an ncu raw+source CSV pair: python scripts/ncu_summary.py raw.csv src.csv kernels.cubin"""
import csv, sys, subprocess
from collections import Counter
raw, src, cubin = sys.argv[1:4]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'sm__inst_executed_pipe_fp64.avg.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread ',
        'smsp__average_warps_issue_stalled_no_inst', 'smsp__average_warps_issue_stalled_wait', 'smsp__average_warps_issue_stalled_branch',
        'smsp__average_warps_issue_stalled_short', 'smsp__average_warps_issue_stalled_long', 'smsp__average_warps_issue_stalled_math',
        'smsp__average_warps_issue_stalled_not_sel', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'smsp__inst_executed.sum ']
for h, u, v in zip(hdr, units, vals):
    if any(w in h + ' ' for w in want):
        print("%-85s %-10s %s" % (h, u, v))
rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, isrc, iex, ith = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
data = [(int(r[ia], 16), r[isrc].strip(), int(r[iex] or 0), int(r[ith] or 0)) for r in rows[2:] if len(r) > iex]
base = data[0][0]
total = sum(d[2] for d in data)
out = subprocess.run("readelf -sW " + cubin, shell=True, capture_output=True, text=True).stdout
syms = []
for l in out.splitlines():
    f = l.split()
    if len(f) >= 8 and f[3] == "FUNC" and "score_kernel" in f[7]:
        syms.append((int(f[1], 16), int(f[2]), f[7].split("$")[-1] if "$" in f[7] else "MAIN"))
syms.sort()
def fn(off):
    for a, sz, nm in syms:
        if a <= off < a + sz:
            return nm
    return "MAIN"
per, thr, dp = Counter(), Counter(), Counter()
for a, s, e, t in data:
    f = fn(a - base)
    per[f] += e
    thr[f] += t
    op = (s.split()[1] if s.startswith('@') else s.split()[0]).split('.')[0]
    if op in ('DFMA', 'DADD', 'DMUL', 'DSETP', 'DMNMX'):
        dp[f] += e
print("total warp instructions %.3e, FP64-pipe share %.1f%%" % (total, 100 * sum(dp.values()) / total))
for f, v in per.most_common(16):
    print("%-48s %5.1f%% of instr  FP64 share %4.1f%%  avg active threads %.1f" % (f[:48], 100 * v / total, 100 * dp[f] / max(1, v), thr[f] / max(1, v)))

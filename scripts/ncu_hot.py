"""Per-instruction view of an ncu source-page CSV: regions of SASS ranked by executed instructions and stall samples.
   python scripts/ncu_hot.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) <= iex:
        continue
    data.append((int(r[ia], 16), r[isrc].strip(), int(r[iex] or 0), int(r[ismp] or 0), r))
base = data[0][0]
tot_ex = sum(d[2] for d in data)
tot_smp = sum(d[3] for d in data)
print("total executed %.3e warp-instr, samples %d, instructions in image %d" % (tot_ex, tot_smp, len(data)))
# contiguous blocks split at branches/labels: group by runs with same executed count
blocks = []
cur = None
for a, s, e, smp, r in data:
    if cur is None or e != cur["ex"]:
        cur = {"start": a - base, "ex": e, "n": 0, "smp": 0, "first": s, "stalls": {}}
        blocks.append(cur)
    cur["n"] += 1
    cur["smp"] += smp
    cur["last"] = s
    for i, h in stall_cols:
        v = int(r[i] or 0)
        if v:
            cur["stalls"][h] = cur["stalls"].get(h, 0) + v
blocks.sort(key=lambda b: -b["smp"])
for b in blocks[:top]:
    st = sorted(b["stalls"].items(), key=lambda kv: -kv[1])[:3]
    print("@%06x n=%3d exec/instr=%9d  instr%%=%5.2f smp%%=%5.2f  %-34s .. %-30s %s" % (
        b["start"], b["n"], b["ex"], 100.0 * b["ex"] * b["n"] / tot_ex, 100.0 * b["smp"] / tot_smp, b["first"][:34], b["last"][:30],
        " ".join("%s=%d" % (k.replace("stall_", ""), v) for k, v in st)))

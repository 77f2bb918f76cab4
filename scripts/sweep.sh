#!/bin/bash
# configs[4] of BASELINE.json: scoring-only sweep, population 10K-1M x points 256-1M, fixed random expressions (seed 888,
# GA mix, max-leafs 40), one B200.  Points whose generation would take more than a few seconds are left to the 8-GPU run.
mkdir -p gpurun_out
out=gpurun_out/sweep.jsonl
: > $out
for cfg in "10000 256" "10000 4096" "10000 65536" "10000 1048576" "100000 256" "100000 4096" "100000 65536" "1000000 256" "1000000 4096"; do
  set -- $cfg
  timeout 300 python bench.py --workload c3 --population $1 --points $2 --steps 5 --warmup 3 --no-cpu-baseline >> $out 2>> gpurun_out/sweep.err || echo "{\"failed\": \"$cfg\"}" >> $out
done
python - <<'PY'
import json
for l in open("gpurun_out/sweep.jsonl"):
    d = json.loads(l)
    if "failed" in d: print(d); continue
    c = d["config"]
    print("%8d x %8d  %10.3f ms  %.3e expr-pts/s  e2e %.3e  frac %.3f" % (c["population"], c["points"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"]))
PY

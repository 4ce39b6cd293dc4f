#!/bin/bash
# Round 2, session L: class-API wall-time breakdown; ring-depth check on the small-chunk workloads.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/e2e_breakdown.py c2 c1 c3 > $OUT/e2e_breakdown_r02l.txt 2>&1; echo "breakdown rc=$?"
cat $OUT/e2e_breakdown_r02l.txt | grep -v Warning | tail -12
for wl in c3 c1; do
  for ring in 0 8; do
    timeout 300 python bench.py --workload $wl --ring $ring --steps 20 --warmup 5 --no-cpu-baseline --no-variants --no-e2e > $OUT/bench_${wl}_ring${ring}_r02l.json 2> $OUT/bench_${wl}_ring${ring}_r02l.err
    python - <<PY
import json
j = json.loads(open("$OUT/bench_${wl}_ring${ring}_r02l.json").read().strip().split("\n")[-1])
print("$wl ring $ring", round(j["value"], 1), j["details"]["engine"], round(j["ms_per_step"] / 16 * 1e3, 1), "us/chunk")
PY
  done
done

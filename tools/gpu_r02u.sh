#!/bin/bash
# Round 2, session U: is the streaming engine host-bound on small chunks?  Host time per submit against the device timeline,
# one and two lanes; FB kernel without the staging prologue.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "factorised" > $OUT/pytest_fb_r02u.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_fb_r02u.log
for w in c1 c3; do
  LAG=6 timeout 120 python tools/trace_engine.py $w 64 > $OUT/trace_${w}_r02u.txt 2>&1; grep "host:" $OUT/trace_${w}_r02u.txt; tail -1 $OUT/trace_${w}_r02u.txt
  PCS_SHARD_LANES=1 LAG=6 timeout 120 python tools/trace_engine.py $w 64 > $OUT/trace_${w}_lanes1_r02u.txt 2>&1; grep "host:" $OUT/trace_${w}_lanes1_r02u.txt; tail -1 $OUT/trace_${w}_lanes1_r02u.txt
done
timeout 300 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu-baseline --no-variants > $OUT/bench_c1_r02u.json 2> $OUT/bench_c1_r02u.err; echo "bench c1 rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/bench_c1_r02u.json").read().strip().split("\n")[-1])
print("value", round(j["value"], 1), "ms/chunk", round(j["details"]["ms_per_chunk"], 4), "e2e", j["e2e"]["value"], "kernel_ms", j["roofline"]["kernel_ms"], "frac", j["roofline"]["frac"])
PY

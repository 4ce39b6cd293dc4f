#!/bin/bash
# Round 2, session F: generic search kernel v2 (sum/max epilogue + LOCATE pass, table twiddles) and 4 table slots.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_r02f.log 2>&1; echo "pytest rc=$?"
tail -12 $OUT/pytest_gpu_r02f.log
for w in c1 c3; do
timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_${w}_r02f.json 2> $OUT/bench_${w}_r02f.err; echo "bench $w rc=$?"
python - $OUT/bench_${w}_r02f.json <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print("value", j["value"], "ms/chunk", j["details"]["ms_per_chunk"], "e2e", (j.get("e2e") or {}).get("value"), "h2d", j["value_with_h2d"]["value"],
          "parity", j.get("parity_vs_single_gpu"), "frac", j["roofline"]["frac"], "stage", j["stage_ms"], "variant", (j.get("variants") or {}).get("cufft_callback", {}).get("ms_per_search"))
except Exception as e:
    print("unreadable", e, open(sys.argv[1]).read()[-1500:], open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-variants > $OUT/bench_c2_r02f.json 2> $OUT/bench_c2_r02f.err; echo "bench c2 rc=$?"
python - $OUT/bench_c2_r02f.json <<'PY'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
print("C2 value", j["value"], "ms/chunk", j["details"]["ms_per_chunk"], "e2e", j["e2e"]["value"], "h2d", j["value_with_h2d"]["value"], "engine", j["details"]["engine"], "stage", j["stage_ms"])
PY

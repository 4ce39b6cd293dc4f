#!/bin/bash
# Round 2, session Q: factorised filter bank (search_fb_kernel, C1) + four tail streams / eight chunks in flight: tests, then
# C1 with and without the factorisation, C3 and C2 (regression), engine timelines.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_r02q.log 2>&1; echo "pytest rc=$?"
tail -6 $OUT/pytest_gpu_r02q.log
show() {
python - "$1" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print("value", round(j["value"], 1), "ms/chunk", round(j["details"]["ms_per_chunk"], 4), "e2e", (j.get("e2e") or {}).get("value"), "h2d", j["value_with_h2d"]["value"],
          "parity", j.get("parity_vs_single_gpu"), "kernel", j["roofline"]["kernel"], round(j["roofline"]["kernel_ms"], 4), "frac", round(j["roofline"]["frac"], 3),
          "bank", j["details"].get("bank_factor"), "stage", j["stage_ms"])
except Exception as e:
    print("unreadable", e, open(sys.argv[1]).read()[-1500:], open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
timeout 300 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu-baseline --no-variants > $OUT/bench_c1_r02q.json 2> $OUT/bench_c1_r02q.err; echo "bench c1 rc=$?"; show $OUT/bench_c1_r02q.json
timeout 300 python bench.py --workload c1 --search-form 3 --steps 20 --warmup 5 --no-cpu-baseline --no-variants > $OUT/bench_c1_form3_r02q.json 2> $OUT/bench_c1_form3_r02q.err; echo "bench c1 form3 rc=$?"; show $OUT/bench_c1_form3_r02q.json
timeout 300 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline --no-variants > $OUT/bench_c3_r02q.json 2> $OUT/bench_c3_r02q.err; echo "bench c3 rc=$?"; show $OUT/bench_c3_r02q.json
timeout 300 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline --no-variants > $OUT/bench_c2_r02q.json 2> $OUT/bench_c2_r02q.err; echo "bench c2 rc=$?"; show $OUT/bench_c2_r02q.json
for w in c1 c3; do LAG=6 timeout 120 python tools/trace_engine.py $w 40 > $OUT/trace_${w}_r02q.txt 2>&1; tail -3 $OUT/trace_${w}_r02q.txt; done

#!/bin/bash
# Round 2, session G: priorities + two tail streams + generic-search lanes: tests, traces, benches.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_r02g.log 2>&1; echo "pytest rc=$?"
tail -6 $OUT/pytest_gpu_r02g.log
for w in c2 c3 c1; do python tools/trace_engine.py $w 40 > $OUT/trace_$w.txt 2>&1; tail -4 $OUT/trace_$w.txt; done
for w in c2 c1 c3; do
timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-variants > $OUT/bench_${w}_r02g.json 2> $OUT/bench_${w}_r02g.err; echo "bench $w rc=$?"
python - $OUT/bench_${w}_r02g.json <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print("value", j["value"], "ms/chunk", j["details"]["ms_per_chunk"], "e2e", (j.get("e2e") or {}).get("value"), "h2d", j["value_with_h2d"]["value"],
          "parity", j.get("parity_vs_single_gpu"), "frac", j["roofline"]["frac"], "stage", j["stage_ms"])
except Exception as e:
    print("unreadable", e, open(sys.argv[1]).read()[-1500:], open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done

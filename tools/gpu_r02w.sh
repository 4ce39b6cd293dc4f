#!/bin/bash
# Round 2, session W: tail graph with the re-seeded flag value: sharded tests, C1 / C3 / C2 benches, 2 ranks on one GPU.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > $OUT/pytest_sharded_r02w.log 2>&1; echo "pytest rc=$?"
tail -4 $OUT/pytest_sharded_r02w.log
show() {
python - "$1" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print("value", round(j["value"], 1), "ms/chunk", round(j["details"]["ms_per_chunk"], 4), "e2e", (j.get("e2e") or {}).get("value"), "stream", (j.get("e2e_stream") or {}).get("value"), "h2d", j["value_with_h2d"]["value"],
          "parity", j.get("parity_vs_single_gpu"), "kernel", j["roofline"]["kernel"], round(j["roofline"]["kernel_ms"], 4), "frac", round(j["roofline"]["frac"], 3), "launches", j.get("gpu_launches"))
except Exception as e:
    print("unreadable", e, open(sys.argv[1]).read()[-1500:], open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
for w in c1 c3 c2; do
timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-variants > $OUT/bench_${w}_r02w.json 2> $OUT/bench_${w}_r02w.err; echo "bench $w rc=$?"; show $OUT/bench_${w}_r02w.json
done

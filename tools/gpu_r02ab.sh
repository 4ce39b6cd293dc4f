#!/bin/bash
# Round 2, session AB: H2D of the ingest rank on its own stream: sharded tests + one-GPU bench (h2d / e2e_stream figures).
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -x -q -k "sharded or rank or ring or ingest or engine or graph" > $OUT/pytest_sharded_r02ab.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_sharded_r02ab.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-variants > $OUT/bench_c2_r02ab.json 2> $OUT/bench_c2_r02ab.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/bench_c2_r02ab.json").read().strip().split("\n")[-1])
print("value", round(j["value"], 1), "h2d", round(j["value_with_h2d"]["value"], 1), "e2e", round(j["e2e"]["value"], 1), "stream", round(j["e2e_stream"]["value"], 1), "parity", j["parity_vs_single_gpu"])
PY

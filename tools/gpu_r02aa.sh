#!/bin/bash
# Round 2, session AA: reproducibility / race test on the factorised path, C5 (64 channels) on one GPU with the final library.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bit_reproducible" > $OUT/pytest_repro_r02aa.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_repro_r02aa.log
timeout 400 python bench.py --workload c5 --steps 20 --warmup 5 > $OUT/bench_c5_r02aa.json 2> $OUT/bench_c5_r02aa.err; echo "bench c5 rc=$?"; tail -2 $OUT/bench_c5_r02aa.err
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/bench_c5_r02aa.json").read().strip().split("\n")[-1])
    print("c5 value", j["value"], "ms/step", j["ms_per_step"], "e2e", (j.get("e2e") or {}).get("value"), j["config"])
except Exception as e:
    print("unreadable", e)
PY

#!/bin/bash
# Round 2, session M: chunks from a registered (page-locked) sample ring -- parity test + the e2e figures.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "registered or streaming_ingest" > $OUT/pytest_r02m.log 2>&1; echo "pytest rc=$?"
tail -15 $OUT/pytest_r02m.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-variants > $OUT/bench_r02m.json 2> $OUT/bench_r02m.err; echo "bench rc=$?"
tail -5 $OUT/bench_r02m.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/bench_r02m.json").read().strip().split("\n")[-1])
print("value", round(j["value"], 1), "e2e", json.dumps(j["e2e"], indent=0)[:900], "stream", j.get("e2e_stream"), "parity", j["parity_vs_single_gpu"])
PY

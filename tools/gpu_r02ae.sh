#!/bin/bash
# Round 2, session AE: the faster stitcher on the GPU paths that use it (stream / post-processing / sharded bit-stream tests) and the class-API e2e.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference_kernels.py -m gpu -q -x -k "stream or postprocessing or ring or ingest" > $OUT/pytest_stitch_r02ae.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_stitch_r02ae.log
timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-variants > $OUT/bench_c2_r02ae.json 2> $OUT/bench_c2_r02ae.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/bench_c2_r02ae.json").read().strip().split("\n")[-1])
print("value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "fill loop", round(j["e2e"]["caller_fill_loop"]["value"], 1), "stream", round(j["e2e_stream"]["value"], 1), "verify", j["verify"]["sha"], "parity", j["parity_vs_single_gpu"])
PY

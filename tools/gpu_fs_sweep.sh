#!/bin/bash
# One gpurun call: GPU tests, then a sweep of the shifted-filter search kernel's tiling knobs against the rotate form.
# Usage:  gpurun --timeout 900 -- bash tools/gpu_fs_sweep.sh [tag]
set -u
TAG=${1:-fs}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -15 $OUT/pytest_gpu_$TAG.log
Q="python bench.py --steps 300 --warmup 20 --e2e-steps 100 --no-cpu-baseline --no-variants"
run() { name=$1; shift; $Q "$@" > $OUT/sweep_${TAG}_$name.json 2> $OUT/sweep_${TAG}_$name.err; echo "$name rc=$?";
  python - "$OUT/sweep_${TAG}_$name.json" "$name" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value %.1f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "search %.4f" % d["stage_ms"]["search"],
          "e2e %.1f" % d["e2e"]["value"], "checksum", d["checksum"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
P
}
run form2 --search-form 2
run i16 --items-per-cta 16
run i32 --items-per-cta 32
run i64 --items-per-cta 64
run i128 --items-per-cta 128
run i256 --items-per-cta 256
run g4i64 --groups-per-cta 4 --items-per-cta 64
run g16i128 --groups-per-cta 16 --items-per-cta 128
run i64f1 --items-per-cta 64 --inflight 1
run form2f1 --search-form 2 --inflight 1

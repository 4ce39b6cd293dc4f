#!/bin/bash
# One gpurun call: full ncu capture of the dominant search kernel (after a plain run of the same command exits 0).
# Usage:  gpurun --timeout 600 -- bash tools/gpu_ncu_search.sh [tag] [extra bench args]
set -u
TAG=${1:-ncu}; shift || true
OUT=gpurun_out
mkdir -p $OUT
SMALL="python bench.py --steps 4 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-variants --inflight 1 $*"
$SMALL > $OUT/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:search_ -s 4 -c 2 -f -o $OUT/prof_search_$TAG $SMALL > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
tail -2 $OUT/ncu_full_$TAG.log

#!/bin/bash
# One gpurun call: GPU tests, the headline bench, the ncu launch list and one full capture of the top kernel.
# Usage (from the repo root):  gpurun --timeout 1500 -- bash tools/gpu_profile.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu_$TAG.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/pytest_gpu_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"
SMALL="python bench.py --steps 4 --warmup 3 --e2e-steps 2 --no-cpu-baseline"
$SMALL > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $SMALL > $OUT/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
$SMALL > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:search_ -s 4 -c 2 -f -o $OUT/prof_search_$TAG $SMALL > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
cat $OUT/bench_$TAG.json | head -c 3000

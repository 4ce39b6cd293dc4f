#!/bin/bash
# gpurun -- bash tools/gpu_test_quick.sh tag "bench args 1" ... : GPU tests, then quick bench lines
set -u
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -12 $OUT/pytest_gpu_$TAG.log
bash tools/gpu_quick.sh $TAG "$@"

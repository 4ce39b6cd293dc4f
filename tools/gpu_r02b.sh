#!/bin/bash
# Round 2, session B: fused search kernel + native sharded engine on one GPU (world 1, and 2-3 ranks sharing cuda:0).
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -s > $OUT/pytest_sharded_r02b.log 2>&1; echo "sharded tests rc=$?"
tail -25 $OUT/pytest_sharded_r02b.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_sharded.py > $OUT/pytest_gpu_r02b.log 2>&1; echo "pytest rc=$?"
tail -8 $OUT/pytest_gpu_r02b.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_r02b.json 2> $OUT/bench_r02b.err; echo "bench rc=$?"
tail -5 $OUT/bench_r02b.err
head -c 6000 $OUT/bench_r02b.json

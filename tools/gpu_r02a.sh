#!/bin/bash
# Round 2, session A: the new parity tests on the benchmarked configs with the round-1 kernels, the whole GPU suite, and
# ncu --set full of the two shipped search kernels (C2: search_fs256_kernel, C1: search_os_kernel<11,2,true>).
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu_r02a.txt 2>&1
python -m pytest tests/test_gpu_bench_configs.py tests/test_doppler_grid.py -m gpu -x -q -s > $OUT/pytest_new_r02a.log 2>&1; echo "new tests rc=$?"
tail -15 $OUT/pytest_new_r02a.log
python -m pytest tests -m gpu -q > $OUT/pytest_gpu_r02a.log 2>&1; echo "pytest rc=$?"
tail -5 $OUT/pytest_gpu_r02a.log
SMALL="python bench.py --steps 4 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-variants --inflight 1"
$SMALL > $OUT/plain_c2_r02a.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:search_fs256 -s 4 -c 1 -f -o $OUT/prof_fs256_r02a $SMALL > $OUT/ncu_full_c2_r02a.log 2>&1
echo "ncu c2 rc=$?"
$SMALL --workload c1 > $OUT/plain_c1_r02a.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:search_os_kernel -s 4 -c 1 -f -o $OUT/prof_os11_r02a $SMALL --workload c1 > $OUT/ncu_full_c1_r02a.log 2>&1
echo "ncu c1 rc=$?"
tail -2 $OUT/plain_c1_r02a.log | head -c 1500

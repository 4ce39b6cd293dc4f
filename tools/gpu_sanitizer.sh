#!/bin/bash
# compute-sanitizer on the factorised-bank path (C1, a few chunks through the engine): memcheck, then racecheck (shared memory).
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/ncu_chunks.py c1 3 > $OUT/sanitizer_memcheck_c1.log 2>&1; echo "memcheck rc=$?"; tail -5 $OUT/sanitizer_memcheck_c1.log
timeout 280 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/ncu_chunks.py c1 2 > $OUT/sanitizer_racecheck_c1.log 2>&1; echo "racecheck rc=$?"; tail -5 $OUT/sanitizer_racecheck_c1.log

#!/bin/bash
# Round 2, session R: ncu --set full of search_fb_kernel (C1) after a plain run of the same command.
set -u
OUT=gpurun_out
mkdir -p $OUT
python tools/ncu_chunks.py c1 10 > $OUT/plain_c1_r02r.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:search_fb_kernel -s 6 -c 1 -f -o $OUT/prof_fb11_r02r python tools/ncu_chunks.py c1 10 > $OUT/ncu_full_c1_r02r.log 2>&1
echo "ncu full c1 rc=$?"; tail -3 $OUT/ncu_full_c1_r02r.log

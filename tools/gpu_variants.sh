#!/bin/bash
# A/B of search-kernel build variants (pycusdr_b200/variants/lib_*.so, selected with PYCUSDR_B200_LIB): search-kernel time
# (profiling pass) and device-resident throughput of C2 on one GPU.
set -u
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/variants.txt
for lib in pycusdr_b200/variants/lib_*.so; do
  for extra in "" "--items-per-cta 64"; do
    PYCUSDR_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e $extra > $OUT/var.json 2> $OUT/var.err
    python - "$lib" "$extra" <<'PY' >> $OUT/variants.txt
import json,sys
try:
    j=json.loads(open('gpurun_out/var.json').read().strip().split('\n')[-1])
    print(sys.argv[1], sys.argv[2], 'value %.1f'%j['value'], 'search_ms %.4f'%j['stage_ms']['search'], 'parity', j['parity_vs_single_gpu'], 'h2d %.1f'%j['value_with_h2d']['value'])
except Exception as e:
    print(sys.argv[1], sys.argv[2], 'FAILED', e, open('gpurun_out/var.err').read()[-300:])
PY
  done
done
cat $OUT/variants.txt

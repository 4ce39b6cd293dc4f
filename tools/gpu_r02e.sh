#!/bin/bash
# Round 2, session E: Doppler-rate test, cuFFT+LTO-callback comparison variant, items-per-CTA sweep on a rank's 32-bin slice.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_doppler_rate.py -m gpu -x -q > $OUT/pytest_rate_r02e.log 2>&1; echo "rate tests rc=$?"
tail -15 $OUT/pytest_rate_r02e.log
timeout 300 python tools/cufft_variant.py --workload c3 --reps 5 > $OUT/cufft_variant_c3.log 2>&1; echo "cufft variant c3 rc=$?"
tail -5 $OUT/cufft_variant_c3.log | cut -c1-1500
timeout 300 python tools/cufft_variant.py --workload c2 --reps 3 > $OUT/cufft_variant_c2.log 2>&1; echo "cufft variant c2 rc=$?"
tail -3 $OUT/cufft_variant_c2.log | cut -c1-1500
: > $OUT/slice_sweep.txt
for ipc in 0 8 16 24 40 72 96; do
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --doppler-bins 32 --items-per-cta $ipc > $OUT/sl.json 2> $OUT/sl.err
  python - $ipc <<'PY' >> $OUT/slice_sweep.txt
import json,sys
try:
    j=json.loads(open('gpurun_out/sl.json').read().strip().split('\n')[-1])
    print('items_per_cta', sys.argv[1], 'value %.1f'%j['value'], 'ms/chunk %.4f'%j['details']['ms_per_chunk'], 'search_ms %.4f'%j['stage_ms']['search'], 'parity', j['parity_vs_single_gpu'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e, open('gpurun_out/sl.err').read()[-300:])
PY
done
cat $OUT/slice_sweep.txt

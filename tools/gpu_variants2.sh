#!/bin/bash
# A/B: generic-kernel build variants on C1, and engine lane / priority knobs on C2.
set -u
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/variants2.txt
report() {
python - "$1" <<'PY' >> gpurun_out/variants2.txt
import json,sys
try:
    j=json.loads(open('gpurun_out/var.json').read().strip().split('\n')[-1])
    print(sys.argv[1], 'value %.1f'%j['value'], 'ms/chunk %.4f'%j['details']['ms_per_chunk'], 'search_ms %.4f'%j['stage_ms']['search'], 'reduce %.4f'%j['stage_ms']['reduce'], 'parity', j['parity_vs_single_gpu'], 'h2d %.1f'%j['value_with_h2d']['value'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e, open('gpurun_out/var.err').read()[-300:])
PY
}
for lib in pycusdr_b200/variants/lib_*.so; do
  PYCUSDR_B200_LIB=$PWD/$lib timeout 300 python bench.py --workload c1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > $OUT/var.json 2> $OUT/var.err
  report "c1 $lib"
done
for knobs in "PCS_SHARD_LANES=2" "PCS_SHARD_LANES=1" "PCS_SHARD_PRIO=1"; do
  env $knobs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > $OUT/var.json 2> $OUT/var.err
  report "c2 $knobs"
  env $knobs timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > $OUT/var.json 2> $OUT/var.err
  report "c3 $knobs"
done
cat $OUT/variants2.txt

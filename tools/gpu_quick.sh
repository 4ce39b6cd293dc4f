#!/bin/bash
# One gpurun call: quick bench lines for a list of knob settings.  Usage: gpurun -- bash tools/gpu_quick.sh tag "args1" "args2" ...
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
i=0
for a in "$@"; do
  i=$((i+1))
  python bench.py --steps 300 --warmup 20 --e2e-steps 100 --no-cpu-baseline --no-variants $a > $OUT/quick_${TAG}_$i.json 2> $OUT/quick_${TAG}_$i.err; rc=$?
  python - "$OUT/quick_${TAG}_$i.json" "$a" $rc <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("[%s] rc=%s value %.1f ms/step %.4f search %.4f e2e %.1f checksum %s" % (sys.argv[2], sys.argv[3], d["value"], d["ms_per_step"],
          d["stage_ms"]["search"], d["e2e"]["value"], d["checksum"]))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
P
done

#!/bin/bash
# gpurun --gpus N -- bash tools/gpu_n_sweep.sh N tag "args1" "args2" ... : bench.py under torchrun on N GPUs for several knob settings
set -u
N=$1; TAG=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
i=0
for a in "$@"; do
  i=$((i+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + i)) bench.py --gpus $N --steps 200 --warmup 20 $a > $OUT/nsweep_${TAG}_$i.json 2> $OUT/nsweep_${TAG}_$i.err; rc=$?
  python - $OUT/nsweep_${TAG}_$i.json "$a" $rc <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("[%s] rc=%s n=%d value %.1f ms/step %.4f e2e %.1f" % (sys.argv[2], sys.argv[3], l["n_gpus"], l["value"], l["ms_per_step"], (l.get("e2e") or {}).get("value", 0)))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
done

#!/bin/bash
# gpurun --gpus N -- bash tools/gpu_scale.sh N tag : what the driver runs at N GPUs (bench.py --steps 20 --warmup 5 under torchrun)
set -u
N=$1; TAG=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) "$@"; }
run bench.py --gpus $N --steps 20 --warmup 5 "$@" > $OUT/scale_${TAG}_n$N.json 2> $OUT/scale_${TAG}_n$N.err; echo "bench rc=$?"
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $OUT/scale_${TAG}_n$N.err | tail -5
python - $OUT/scale_${TAG}_n$N.json <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e=l.get("e2e") or {}
print(l["n_gpus"], "value", round(l["value"],1), "ms/step", round(l["ms_per_step"],4), "h2d", round(l["value_with_h2d"]["value"],1), "e2e", round(e.get("value",0),1), "parity", l.get("parity_vs_single_gpu"), "share", round(l["roofline"]["kernel_share_of_step"],3))
PY

#!/bin/bash
# gpurun --gpus N -- bash tools/gpu_n.sh N tag [bench args] : bench.py under torchrun on N GPUs (+ the sharded parity check)
set -u
N=$1; TAG=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) "$@"; }
run tools/check_sharded_gpu.py > $OUT/sharded_check_${TAG}_n$N.log 2>&1; echo "sharded check rc=$?"; tail -1 $OUT/sharded_check_${TAG}_n$N.log
run bench.py --gpus $N --steps 200 --warmup 20 "$@" > $OUT/scale_${TAG}_n$N.json 2> $OUT/scale_${TAG}_n$N.err; echo "bench rc=$?"
tail -3 $OUT/scale_${TAG}_n$N.err
python - $OUT/scale_${TAG}_n$N.json <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(l["n_gpus"], "value", round(l["value"],1), "ms/step", round(l["ms_per_step"],4), "e2e", l.get("e2e"))
PY

#!/bin/bash
# gpurun --gpus N -- bash tools/gpu_scale_c5.sh N tag : bench.py as the driver runs it + the C5 (64 channels) workload on N GPUs
set -u
N=$1; TAG=$2
bash tools/gpu_scale.sh $N $TAG
OUT=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --workload c5 --steps 40 --warmup 8 > $OUT/scale_${TAG}_c5_n$N.json 2> $OUT/scale_${TAG}_c5_n$N.err; echo "c5 rc=$?"
python - $OUT/scale_${TAG}_c5_n$N.json <<'PY'
import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("c5", l["n_gpus"], "value", round(l["value"],1), "ms/step", round(l["ms_per_step"],4), l["launch_mode"])
PY

#!/bin/bash
# gpurun --gpus 8 -- bash tools/gpu_multi8.sh [tag] : lean 8-GPU validation (parity check + bench at 8 and 4 ranks)
set -u
TAG=${1:-r01}; OUT=gpurun_out; mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) "${@:2}"; }
run 8 tools/check_sharded_gpu.py > $OUT/sharded_check_${TAG}_n8.log 2>&1; echo "sharded check rc=$?"; grep "sharded check" $OUT/sharded_check_${TAG}_n8.log
for n in 8 4; do
  run $n bench.py --gpus $n --steps 400 --warmup 24 > $OUT/scale_${TAG}_n$n.json 2> $OUT/scale_${TAG}_n$n.err; echo "n$n rc=$?"
done
run 8 bench.py --gpus 8 --steps 100 --warmup 8 --workload c4 > $OUT/scale_${TAG}_c4_n8.json 2> $OUT/scale_${TAG}_c4_n8.err; echo "c4 n8 rc=$?"
for f in $OUT/scale_${TAG}_n8.json $OUT/scale_${TAG}_n4.json $OUT/scale_${TAG}_c4_n8.json $OUT/scale_${TAG}_c5_n8.json; do python - "$f" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], l["n_gpus"], round(l["value"],1), l["unit"], round(l["ms_per_step"],4), "ms/step", l.get("stage_ms"))
except Exception as e: print(sys.argv[1], "unreadable", e)
PY
done
tail -n 3 $OUT/scale_${TAG}_n8.err

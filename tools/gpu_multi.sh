#!/bin/bash
# gpurun --gpus N -- bash tools/gpu_multi.sh N [tag] : sharded parity check + scaling bench on N GPUs of one box
set -u
N=${1:-2}; TAG=${2:-r01}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) "${@:2}"; }
run $N tools/check_sharded_gpu.py > $OUT/sharded_check_${TAG}_n$N.log 2>&1; echo "sharded check rc=$?"; tail -5 $OUT/sharded_check_${TAG}_n$N.log

python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 50 > $OUT/scale_${TAG}_n1.json 2> $OUT/scale_${TAG}_n1.err; echo "n1 rc=$?"
for n in 2 4 8; do
  if [ $n -le $N ]; then
    run $n bench.py --gpus $n --steps 200 --warmup 20 > $OUT/scale_${TAG}_n$n.json 2> $OUT/scale_${TAG}_n$n.err; echo "n$n rc=$?"
    run $n bench.py --gpus $n --steps 200 --warmup 20 --exchange nccl > $OUT/scale_${TAG}_n${n}_nccl.json 2> $OUT/scale_${TAG}_n${n}_nccl.err; echo "n$n nccl rc=$?"
  fi
done
for f in $OUT/scale_${TAG}_n*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(l["n_gpus"], round(l["value"],1), l["unit"], round(l["ms_per_step"],4), "ms/step", l["config"]["parallelism"][:60])
except Exception as e: print("unreadable", e)
PY
done
tail -3 $OUT/scale_${TAG}_n$N.err

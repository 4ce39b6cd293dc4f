#!/bin/bash
# N-GPU session: sharded-engine parity with one GPU per rank, then bench.py under torchrun exactly as the driver launches it.
# Usage: gpurun --gpus N --timeout 900 -- bash tools/gpu_multi.sh N [tag] [extra bench args]
set -u
N=${1:-2}; TAG=${2:-r02}; shift 2 || true
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L | head -8
PCS_TEST_ONE_GPU_PER_RANK=1 timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "ranks_sharing" > $OUT/pytest_sharded_n${N}_$TAG.log 2>&1; echo "sharded tests (one GPU per rank) rc=$?"
tail -4 $OUT/pytest_sharded_n${N}_$TAG.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 "$@" > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err; echo "bench N=$N rc=$?"
grep -v "^W\|^\[W\|NCCL\|^$" $OUT/bench_n${N}_$TAG.err | tail -8
python - $OUT/bench_n${N}_$TAG.json <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print("value", j["value"], "ms/chunk", j["details"]["ms_per_chunk"], "e2e", (j.get("e2e") or {}).get("value"), "h2d", j["value_with_h2d"]["value"],
          "parity", j.get("parity_vs_single_gpu"), "kernel_ms", j["roofline"]["kernel_ms"], "share", j["roofline"]["kernel_share_of_step"],
          "launches/chunk", j["launches_per_chunk"], "stage", j["stage_ms"])
except Exception as e:
    print("unreadable", e, open(sys.argv[1]).read()[-2000:])
PY

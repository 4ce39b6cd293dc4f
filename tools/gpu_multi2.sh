#!/bin/bash
# N-GPU bench session (after tools/gpu_multi.sh validated the protocol): C2 as the driver runs it, optionally C4.
set -u
N=${1:-8}; TAG=${2:-r02}; shift 2 || true
OUT=gpurun_out
mkdir -p $OUT
run() {
  w=$1; steps=$2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $w --steps $steps --warmup 5 > $OUT/bench_${w}_n${N}_$TAG.json 2> $OUT/bench_${w}_n${N}_$TAG.err; echo "bench $w N=$N rc=$?"
  grep -v "^W\|^\[W\|NCCL\|^$\|\*\*\*\*\|OMP_NUM" $OUT/bench_${w}_n${N}_$TAG.err | tail -5
  python - $OUT/bench_${w}_n${N}_$TAG.json <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print("value", j["value"], "ms/chunk", j["details"]["ms_per_chunk"], "e2e", (j.get("e2e") or {}).get("value"), "h2d", j["value_with_h2d"]["value"],
          "parity", j.get("parity_vs_single_gpu"), "kernel_ms", j["roofline"]["kernel_ms"], "share", j["roofline"]["kernel_share_of_step"],
          "launches/chunk", j["launches_per_chunk"])
except Exception as e:
    print("unreadable", e, open(sys.argv[1]).read()[-2000:])
PY
}
run c2 20
if [ "${1:-}" = "c4" ]; then run c4 3; fi

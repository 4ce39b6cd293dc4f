#!/bin/bash
# gpurun -- bash tools/gpu_workloads.sh [tag] : the other BASELINE configs on one GPU
set -u
TAG=${1:-r01}; OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
for w in c1 c3; do
  python bench.py --workload $w --steps 200 --warmup 20 > $OUT/bench_${w}_$TAG.json 2> $OUT/bench_${w}_$TAG.err; echo "$w rc=$?"
done
python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_c4_$TAG.json 2> $OUT/bench_c4_$TAG.err; echo "c4 rc=$?"
python bench.py --workload c5 --steps 30 --warmup 5 > $OUT/bench_c5_$TAG.json 2> $OUT/bench_c5_$TAG.err; echo "c5 rc=$?"
python bench.py --impl reference --workload c3 --steps 50 --warmup 5 > $OUT/bench_ref_c3_$TAG.json 2> $OUT/bench_ref_c3_$TAG.err; echo "ref c3 rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref_c2_$TAG.json 2> $OUT/bench_ref_c2_$TAG.err; echo "ref c2 rc=$?"
for f in $OUT/bench_c?_$TAG.json $OUT/bench_ref_c?_$TAG.json; do python - "$f" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=l.get("e2e") or {}
    print(sys.argv[1].split('/')[-1], round(l["value"],2), l["unit"], round(l["ms_per_step"],4), "ms/step  e2e", round(e.get("value",0),2), (l.get("roofline") or {}).get("frac"))
except Exception as ex: print(sys.argv[1], "unreadable", ex)
PY
done
tail -2 $OUT/bench_c4_$TAG.err $OUT/bench_c5_$TAG.err

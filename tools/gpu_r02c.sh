#!/bin/bash
# Round 2, session C: the kernel that ships (64-lane finish, 112 registers): GPU suite, bench both arms, launch list, ncu full.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_r02c.log 2>&1; echo "pytest rc=$?"
tail -6 $OUT/pytest_gpu_r02c.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_r02c.json 2> $OUT/bench_r02c.err; echo "bench rc=$?"
tail -3 $OUT/bench_r02c.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_r02c.json 2> $OUT/bench_ref_r02c.err; echo "ref rc=$?"
tail -3 $OUT/bench_ref_r02c.err
SMALL="python bench.py --steps 2 --warmup 3 --chunks-per-step 4 --no-cpu-baseline --no-e2e"
$SMALL > $OUT/plain_r02c.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file $OUT/launches_r02c.csv $SMALL > $OUT/ncu_launch_r02c.log 2>&1
echo "ncu launches rc=$?"
$SMALL > $OUT/plain2_r02c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:search_fs256 -s 6 -c 1 -f -o $OUT/prof_fs256_r02c $SMALL > $OUT/ncu_full_r02c.log 2>&1
echo "ncu full rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_r02c.json", "gpurun_out/bench_ref_r02c.json"):
    try:
        j = json.loads(open(f).read().strip().split("\n")[-1])
        print(f, "value", j["value"], "e2e", (j.get("e2e") or {}).get("value"), "h2d", (j.get("value_with_h2d") or {}).get("value"),
              "stage", j.get("stage_ms"), "verify", j.get("verify"), "parity", j.get("parity_vs_single_gpu"))
    except Exception as e:
        print(f, "unreadable", e)
PY

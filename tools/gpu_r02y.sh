#!/bin/bash
# Round 2, session Y (final code): GPU suite, bench both arms, ncu launch list + full captures of the C2 and C1 search kernels.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu_r02y.log 2>&1; echo "pytest rc=$?"
tail -4 $OUT/pytest_gpu_r02y.log
timeout 700 python bench.py --steps 20 --warmup 5 > $OUT/bench_r02y.json 2> $OUT/bench_r02y.err; echo "bench rc=$?"
tail -3 $OUT/bench_r02y.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_r02y.json 2> $OUT/bench_ref_r02y.err; echo "ref rc=$?"
python tools/ncu_chunks.py c2 10 > $OUT/plain_c2_r02y.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 90 -c 80 --csv --log-file $OUT/launches_r02y.csv python tools/ncu_chunks.py c2 10 > $OUT/ncu_launch_r02y.log 2>&1
echo "ncu launches rc=$?"
python tools/ncu_chunks.py c2 10 > $OUT/plain2_c2_r02y.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:search_fs256 -s 3 -c 1 -f -o $OUT/prof_fs256_r02y python tools/ncu_chunks.py c2 10 > $OUT/ncu_full_c2_r02y.log 2>&1
echo "ncu full c2 rc=$?"
python tools/ncu_chunks.py c1 10 > $OUT/plain_c1_r02y.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:search_fb_kernel -s 6 -c 1 -f -o $OUT/prof_fb11_r02y python tools/ncu_chunks.py c1 10 > $OUT/ncu_full_c1_r02y.log 2>&1
echo "ncu full c1 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_r02y.json", "gpurun_out/bench_ref_r02y.json"):
    try:
        j = json.loads(open(f).read().strip().split("\n")[-1])
        print(f, "value", j["value"], "e2e", (j.get("e2e") or {}).get("value"), "h2d", (j.get("value_with_h2d") or {}).get("value"),
              "verify", (j.get("verify") or {}).get("sha"), "parity", j.get("parity_vs_single_gpu"), "variant", (j.get("variants") or {}).get("cufft_callback", {}).get("ms_per_search"), "cfg", j["config"])
    except Exception as e:
        print(f, "unreadable", e)
PY

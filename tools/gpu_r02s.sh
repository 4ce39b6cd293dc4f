#!/bin/bash
# Round 2, session S: search_fb_kernel with the complete-binary-bank epilogue (static selection, shared partial sums, three
# buffers in rotation: four CTAs per SM): targeted tests, C1 bench, ncu capture.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference_kernels.py tests/test_gpu_sharded.py -m gpu -x -q -k "factorised or search_energy or cc11xx or demod_surface" > $OUT/pytest_fb_r02s.log 2>&1; echo "pytest rc=$?"
tail -6 $OUT/pytest_fb_r02s.log
show() {
python - "$1" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print("value", round(j["value"], 1), "ms/chunk", round(j["details"]["ms_per_chunk"], 4), "e2e", (j.get("e2e") or {}).get("value"), "h2d", j["value_with_h2d"]["value"],
          "parity", j.get("parity_vs_single_gpu"), "kernel", j["roofline"]["kernel"], round(j["roofline"]["kernel_ms"], 4), "frac", round(j["roofline"]["frac"], 3),
          "bank", j["details"].get("bank_factor"), "stage", j["stage_ms"])
except Exception as e:
    print("unreadable", e, open(sys.argv[1]).read()[-1500:], open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
timeout 300 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu-baseline --no-variants > $OUT/bench_c1_r02s.json 2> $OUT/bench_c1_r02s.err; echo "bench c1 rc=$?"; show $OUT/bench_c1_r02s.json
python tools/ncu_chunks.py c1 10 > $OUT/plain_c1_r02s.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:search_fb_kernel -s 6 -c 1 -f -o $OUT/prof_fb11_r02s python tools/ncu_chunks.py c1 10 > $OUT/ncu_full_c1_r02s.log 2>&1
echo "ncu full c1 rc=$?"; tail -2 $OUT/ncu_full_c1_r02s.log
LAG=6 timeout 120 python tools/trace_engine.py c1 40 > $OUT/trace_c1_r02s.txt 2>&1; tail -3 $OUT/trace_c1_r02s.txt

#!/bin/bash
# Round 2, session Z: smoke(), the edge tests with the tie report (-s: the [parity] lines), nothing else.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_r02z.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_r02z.log
timeout 600 python -m pytest tests/test_gpu_edges.py -m gpu -q -s > $OUT/pytest_edges_r02z.log 2>&1; echo "pytest rc=$?"
grep "parity\]" $OUT/pytest_edges_r02z.log | sort | uniq -c | sort -rn | head -20; tail -3 $OUT/pytest_edges_r02z.log

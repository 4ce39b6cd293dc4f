#!/bin/bash
# gpurun -- bash tools/gpu_refgolden.sh : generate the reference-kernel golden fixtures on the B200 and check the oracle against them
set -u
mkdir -p gpurun_out
python -m oracle.ref_gpu.make_golden_gpu > gpurun_out/refgolden.log 2>&1; echo "golden rc=$?"; tail -15 gpurun_out/refgolden.log
cp gpurun_out/refgpu_*.npz tests/golden/ 2>/dev/null
python -m pytest tests/test_oracle_refgpu.py -q -x 2>&1 | tail -30 > gpurun_out/refgolden_pytest.log; cat gpurun_out/refgolden_pytest.log

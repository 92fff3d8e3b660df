#!/bin/bash
# One GPU call: TC kernel tests, model parity tests, kernel timings.  Usage: tools/gpu_check.sh [tag]
tag=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu > gpurun_out/pytest_tc_$tag.log 2>&1; echo "tc rc=$?" 
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_parity_$tag.log 2>&1; echo "parity rc=$?"
for m in EDGE NODE; do timeout 120 python tools/tc_profile.py 2000000 $m bf16; done > gpurun_out/tc_times_$tag.log 2>&1
cat gpurun_out/tc_times_$tag.log
tail -5 gpurun_out/pytest_tc_$tag.log; tail -5 gpurun_out/pytest_parity_$tag.log

#!/bin/bash
# round-2 GPU session H: event-driven MMA issue in the forward kernel: tc tests, kernel timings, bench without extras
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for m in EDGE NODE; do timeout 120 python tools/tc_profile.py 8000000 $m f16; done 2>&1 | tee gpurun_out/r2h_tc_profile.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --kernel-summary gpurun_out/r2h_kernels_f16_4m.txt 2>gpurun_out/r2h_bench.err | tee gpurun_out/r2h_bench.json | cut -c1-200
head -8 gpurun_out/r2h_kernels_f16_4m.txt | cut -c1-120

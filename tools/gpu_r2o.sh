#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "node_level" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --kernel-summary gpurun_out/r2o_kernels_f16_4m.txt 2>gpurun_out/r2o_bench.err | tee gpurun_out/r2o_bench.json | cut -c1-200
head -8 gpurun_out/r2o_kernels_f16_4m.txt | cut -c1-130

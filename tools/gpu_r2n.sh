#!/bin/bash
# round-2 GPU session N: node-level layer-1 backward: full suite, A/B bench (edge-level vs node-level), kernel table
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
FVGN_NODE_LEVEL_LAYER1=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --kernel-summary gpurun_out/r2n_kernels_f16_4m.txt 2>gpurun_out/r2n_bench.err | tee gpurun_out/r2n_bench.json | cut -c1-200
head -16 gpurun_out/r2n_kernels_f16_4m.txt | cut -c1-130

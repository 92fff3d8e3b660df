#!/bin/bash
# round-2 GPU session X: last-block dead edge stream + two-kernel node-level layer 1: suite under both settings, A/B bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest0.log 2>&1; echo "pytest(node-level 0) rc=$?"; tail -2 gpurun_out/r2x_pytest0.log
FVGN_NODE_LEVEL_LAYER1=2 timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest2.log 2>&1; echo "pytest(node-level 2) rc=$?"; tail -2 gpurun_out/r2x_pytest2.log
FVGN_NODE_LEVEL_LAYER1=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | cut -c1-200
FVGN_NODE_LEVEL_LAYER1=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --kernel-summary gpurun_out/r2x_kernels_f16_4m.txt 2>gpurun_out/r2x_bench.err | tee gpurun_out/r2x_bench.json | cut -c1-200
FVGN_NODE_LEVEL_LAYER1=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --net TransFVGN_v2 2>/dev/null | cut -c1-200

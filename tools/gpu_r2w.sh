#!/bin/bash
# round-2 GPU session W: two-kernel node-level layer-1 backward: tests, A/B bench (edge-level vs node-level 2), kernel table; full suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k node_level > gpurun_out/r2w_node.log 2>&1; echo "node rc=$?"; tail -3 gpurun_out/r2w_node.log
FVGN_NODE_LEVEL_LAYER1=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | cut -c1-200
FVGN_NODE_LEVEL_LAYER1=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --kernel-summary gpurun_out/r2w_kernels_f16_4m_node2.txt 2>gpurun_out/r2w_bench.err | tee gpurun_out/r2w_bench.json | cut -c1-200
head -18 gpurun_out/r2w_kernels_f16_4m_node2.txt | cut -c1-130
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2w_pytest.log

#!/bin/bash
# round-2 GPU session AM: 16-bit gradient streams (f16 mode): tests, parity suite, A/B bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "latent or last_block or node_level" 2>&1 | tail -12 | cut -c1-300
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2am_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2am_pytest.log | tail -8 | cut -c1-300
FVGN_GRAD16=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --kernel-summary gpurun_out/r2am_kernels_f16_4m.txt 2>gpurun_out/r2am_bench.err | tee gpurun_out/r2am_bench.json | cut -c1-200
head -16 gpurun_out/r2am_kernels_f16_4m.txt | cut -c1-130; tail -3 gpurun_out/r2am_bench.err

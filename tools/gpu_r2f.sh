#!/bin/bash
# round-2 GPU session F: staged reductions + degree prescale: full suite, bench without extras, kernel table
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --kernel-summary gpurun_out/r2f_kernels_f16_4m.txt 2>gpurun_out/r2f_bench.err | tee gpurun_out/r2f_bench.json | cut -c1-330
head -24 gpurun_out/r2f_kernels_f16_4m.txt | cut -c1-150

#!/bin/bash
# round-2 GPU session B: regression tests after the graph-step fix; 4 M-cell bench at the driver's step counts with a per-kernel table
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --precision bf16 --kernel-summary gpurun_out/r2b_kernels_bf16_4m.txt 2>gpurun_out/r2b_bench_bf16.err | tee gpurun_out/r2b_bench_bf16.json | cut -c1-300
head -40 gpurun_out/r2b_kernels_bf16_4m.txt
nvidia-smi --query-gpu=name,power.limit,power.max_limit,clocks.max.sm,clocks.max.mem,memory.total --format=csv

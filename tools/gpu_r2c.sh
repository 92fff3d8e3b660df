#!/bin/bash
# round-2 GPU session C: full GPU suite after the stale-pack / Z1 gating fixes; 4 M bench bf16 + f16
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c_pytest.log
for p in bf16 f16; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision $p --kernel-summary gpurun_out/r2c_kernels_${p}_4m.txt 2>gpurun_out/r2c_bench_$p.err | tee gpurun_out/r2c_bench_$p.json | cut -c1-300; done
head -30 gpurun_out/r2c_kernels_bf16_4m.txt

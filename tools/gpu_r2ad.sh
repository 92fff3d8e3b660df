#!/bin/bash
# round-2 GPU session AD: 16-bit latent streams: tests, parity suite, A/B bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "latent or last_block" 2>&1 | tail -15
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2ad_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2ad_pytest.log | cut -c1-300
FVGN_LATENTS16=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --kernel-summary gpurun_out/r2ad_kernels_f16_4m.txt 2>gpurun_out/r2ad_bench.err | tee gpurun_out/r2ad_bench.json | cut -c1-200
head -12 gpurun_out/r2ad_kernels_f16_4m.txt | cut -c1-130

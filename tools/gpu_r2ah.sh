#!/bin/bash
# round-2 GPU session AH: token attention kernels on the GPU: Transolver tests + full suite + v2 bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2ah_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2ah_pytest.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --net TransFVGN_v2 --mp 3 --kernel-summary gpurun_out/r2ah_kernels_v2_4m.txt 2>/dev/null | cut -c1-200
grep -n "token_attention\|bmm\|softmax\|cutlass\|gemm\|sgemm" gpurun_out/r2ah_kernels_v2_4m.txt | cut -c1-160

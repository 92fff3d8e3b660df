#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2u_pytest.log
for tcg in 0 1; do FVGN_TC_GEMM=$tcg timeout 600 python bench.py --net TransFVGN_v2 --mp 3 --steps 5 --warmup 3 --no-cpu-baseline --no-extras --kernel-summary gpurun_out/r2u_kernels_v2_4m_tcgemm$tcg.txt 2>/dev/null | cut -c1-200; done
grep -i "gemm\|cutlass" gpurun_out/r2u_kernels_v2_4m_tcgemm1.txt | cut -c1-150

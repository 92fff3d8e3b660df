#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -15

#!/bin/bash
# round-2 GPU session AP (2 GPUs): partition parity check with 16-bit latent streams (exchange-free halo), full suite, 2-GPU bench with cells record
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/check_partition_gpu.py 2>&1 | grep -v "^W\|\*\*\*\|OMP" | tee gpurun_out/r2ap_partition_check_2gpu.txt | tail -12
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2ae_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2ae_pytest.log | cut -c1-200
bash tools/gpu_r2ab.sh 2

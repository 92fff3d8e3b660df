#!/bin/bash
# round-2 GPU session AC (8 GPUs): DP-8 bench line + `cells` record at 4 M cells; one 16 M-cell mesh cell-partitioned over 8 GPUs
mkdir -p gpurun_out
bash tools/gpu_r2ab.sh 8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --cells 16000000 --parallel cells --no-extras --no-cpu-baseline > gpurun_out/r2ac_bench_cells8_16m.json 2>gpurun_out/r2ac_bench16m.err; echo "bench16m rc=$?"
tail -c 1500 gpurun_out/r2ac_bench_cells8_16m.json; tail -3 gpurun_out/r2ac_bench16m.err

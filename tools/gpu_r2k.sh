#!/bin/bash
# 2 GPUs: ONE 1 M-cell mesh partitioned over 2 ranks (500 k cells per rank, the per-rank size of the 4 M / 8 GPU case), eager vs whole-step CUDA graph
for g in "" "--graph"; do for hl in 20 3; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --parallel cells --cells 1000000 --halo-layers $hl --steps 20 --warmup 3 --no-cpu-baseline --no-extras $g 2>gpurun_out/r2k.err | cut -c1-200,740-1100; tail -2 gpurun_out/r2k.err | cut -c1-300
done; done

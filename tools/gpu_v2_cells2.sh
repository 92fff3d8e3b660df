#!/bin/bash
# 2-GPU call: TransFVGN_v2 on ONE 1 M-cell mesh, cell-partitioned (halo mode of the Transolver kernels over NCCL), plus
# the whole-step CUDA graph on a small TransFVGN_v2 case (single GPU).
tag=${1:-x}
mkdir -p gpurun_out
for c in 10000 30000; do
  timeout 300 python bench.py --net TransFVGN_v2 --mp 3 --cells $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_v2_${c}_eager_$tag.log 2>&1; echo "eager $c rc=$?"
  grep '^{' gpurun_out/bench_v2_${c}_eager_$tag.log | cut -c1-250
  timeout 300 python bench.py --net TransFVGN_v2 --mp 3 --cells $c --steps 20 --warmup 5 --no-cpu-baseline --graph > gpurun_out/bench_v2_${c}_graph_$tag.log 2>&1; echo "graph $c rc=$?"
  grep '^{' gpurun_out/bench_v2_${c}_graph_$tag.log | cut -c1-250 || tail -5 gpurun_out/bench_v2_${c}_graph_$tag.log
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 \
  --parallel cells --net TransFVGN_v2 --mp 3 --cells 1000000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2_cells2_$tag.log 2>&1; echo "cells2 rc=$?"
grep '^{' gpurun_out/bench_v2_cells2_$tag.log | cut -c1-1200 || tail -20 gpurun_out/bench_v2_cells2_$tag.log

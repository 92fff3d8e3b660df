#!/bin/bash
# round-2 GPU session AQ (8 GPUs): DP-8 line and the 16 M-cell cell-partition line again with 5 warm-up / 20 timed steps
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2aq_bench_dp8_4m.json 2>gpurun_out/r2aq_dp8.err; echo "dp8 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --cells 16000000 --parallel cells --no-extras --no-cpu-baseline > gpurun_out/r2aq_bench_cells8_16m.json 2>gpurun_out/r2aq_16m.err; echo "16m rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2aq_bench_dp8_4m.json','gpurun_out/r2aq_bench_cells8_16m.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, {k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling','steps','warmup','clocks')}, 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'])
PY

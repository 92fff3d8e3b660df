#!/bin/bash
# round-2 GPU session Q: 8 M cells on ONE GPU (memory / throughput at twice the headline size), 16 M attempt guarded by a memory estimate
mkdir -p gpurun_out
timeout 900 python bench.py --cells 8000000 --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/r2q_8m.err | tee gpurun_out/r2q_bench_8m.json | cut -c1-250
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_bench_8m.json').read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["step_roofline"])
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
timeout 600 python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
import bench
# peak memory of one 8 M-cell step
import argparse
torch.cuda.reset_peak_memory_stats()
from gen_fvgn_steady_b200.mesh.batching import graphs_from_meshes
dev=torch.device("cuda")
mesh,uvp=bench.make_mesh(8_000_000,0,dev)
graphs=graphs_from_meshes([mesh],[uvp],dev); del mesh
job=bench.Job(dev,0,1,graphs,"EPD",6,"f16")
for _ in range(2): job.step(False)
torch.cuda.synchronize()
print("peak allocated GB at 8 M cells:", torch.cuda.max_memory_allocated()/1e9, "reserved", torch.cuda.max_memory_reserved()/1e9)
PY

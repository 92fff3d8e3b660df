"""Sweep of dz_incidence_kernel variants (rows in flight x occupancy) on the bench's 4 M-cell mesh; needs a library built with
FVGN_EXTRA_NVCC_FLAGS=-DFVGN_DZI_SWEEP.  One process per variant (the variant is latched at first use):
    for v in 0 1 2 3 4 5; do FVGN_DZI_VARIANT=$v python tools/dzi_sweep.py; done"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from gen_fvgn_steady_b200.plan import GraphPlan
from gen_fvgn_steady_b200.mesh.batching import graphs_from_meshes

dev = torch.device("cuda")
mesh, uvp = bench.make_mesh(4_000_000, 0, dev)
plan = GraphPlan.of(graphs_from_meshes([mesh], [uvp], dev)[0])
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
params = [rn(128, 384) / 384 ** 0.5, 0.1 * rn(128), rn(128, 128) / 128 ** 0.5, 0.1 * rn(128), rn(128, 128) / 128 ** 0.5,
          0.1 * rn(128), 1 + 0.1 * rn(128), 0.1 * rn(128)]
run, _ = bench.edge_backward_runner(plan, dev, "f16", params)
for _ in range(3):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        run()
    torch.cuda.synchronize()
names = {0: "R1 MINB5", 1: "R1 MINB4", 2: "R2 MINB4", 3: "R2 MINB3", 4: "R1 MINB6", 5: "R2 MINB5"}
v = int(os.environ.get("FVGN_DZI_VARIANT", "0"))
for e in prof.key_averages():
    if "dz_incidence" in e.key or "bwd_node_kernel" in e.key:
        print(f"variant {v} ({names.get(v)}): {e.key[:60]:60s} {e.device_time_total / e.count / 1e3:.3f} ms x {e.count}")

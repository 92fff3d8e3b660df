"""Profiling driver (used under ncu; see profiles/): the fused edge-MLP backward call of one GnBlock -- the call bench.py's
`roofline` times -- on the bench's own 4 M-cell quad mesh (8 M edges, 4 M nodes), a few repetitions."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gen_fvgn_steady_b200.plan import GraphPlan
from gen_fvgn_steady_b200.mesh.batching import graphs_from_meshes

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
prec = sys.argv[2] if len(sys.argv) > 2 else "f16"
dev = torch.device("cuda")
mesh, uvp = bench.make_mesh(cells, 0, dev)
plan = GraphPlan.of(graphs_from_meshes([mesh], [uvp], dev)[0])
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
params = [rn(128, 384) / 384 ** 0.5, 0.1 * rn(128), rn(128, 128) / 128 ** 0.5, 0.1 * rn(128), rn(128, 128) / 128 ** 0.5,
          0.1 * rn(128), 1 + 0.1 * rn(128), 0.1 * rn(128)]
run, label = bench.edge_backward_runner(plan, dev, prec, params)
for _ in range(2):
    run()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(3):
    run()
ev1.record()
torch.cuda.synchronize()
print(f"{label}: N={plan.N} E={plan.E} {prec}: {ev0.elapsed_time(ev1) / 3:.3f} ms per call")

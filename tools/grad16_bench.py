"""A/B of the backward kernels with fp32 vs 16-bit gradient streams (fvgn_mlp_desc.d_outh / d_in1h) on the bench's 4 M-cell mesh:
per-kernel device time (CUPTI) of the fused EDGE and NODE backward calls."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from gen_fvgn_steady_b200 import _lib, ops
from gen_fvgn_steady_b200.plan import GraphPlan
from gen_fvgn_steady_b200.mesh.batching import graphs_from_meshes

dev = torch.device("cuda")
prec, hdt = "f16", torch.float16
mesh, uvp = bench.make_mesh(int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000, 0, dev)
plan = GraphPlan.of(graphs_from_meshes([mesh], [uvp], dev)[0])
N, E = plan.N, plan.E
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
mlp = lambda k1: [rn(128, k1) / k1 ** 0.5, 0.1 * rn(128), rn(128, 128) / 128 ** 0.5, 0.1 * rn(128), rn(128, 128) / 128 ** 0.5,
                  0.1 * rn(128), 1 + 0.1 * rn(128), 0.1 * rn(128)]
eb, nb = mlp(384), mlp(192)
aggh, eh, xh, a2h = rn(N, 128).to(hdt), rn(E, 128).to(hdt), rn(N, 128).to(hdt), rn(N, 64).to(hdt)
d_e_out, d_x_out, d_a1h = rn(E, 128), rn(N, 128), rn(N, 64).to(hdt)
ze = ops.new_z1(_lib.FVGN_MLP_EDGE, prec, E, d_e_out)
zn = ops.new_z1(_lib.FVGN_MLP_NODE, prec, N, d_e_out)
ops.mlp_forward(_lib.FVGN_MLP_EDGE, prec, E, eb, None, None, plan.edge_s, plan.edge_r, want_out=False, z1=ze, in0h=aggh, in1h=eh, want_outh=True)
ops.mlp_forward(_lib.FVGN_MLP_NODE, prec, N, nb, None, None, want_out=False, z1=zn, in0h=a2h, in1h=xh, want_outh=True)
d_agg, d_a2h = torch.empty((N, 128), dtype=hdt, device=dev), torch.empty((N, 64), dtype=hdt, device=dev)
d_e, d_x = torch.empty((E, 128), device=dev), torch.empty((N, 128), device=dev)
d_eh, d_xh = torch.empty((E, 128), dtype=hdt, device=dev), torch.empty((N, 128), dtype=hdt, device=dev)
d_e_outh, d_x_outh = d_e_out.to(hdt), d_x_out.to(hdt)


def run(g16):
    ops.mlp_backward(_lib.FVGN_MLP_NODE, prec, N, nb, None, None, None, None, None if g16 else d_x_out, None, None, None if g16 else d_x,
                     z1=zn, in0h=a2h, in1h=xh, d_in0h=d_a2h, d_in0_row_ptr=plan.inc_ptr, d_outh=d_x_outh if g16 else None,
                     d_in1h=d_xh if g16 else None)
    ops.mlp_backward(_lib.FVGN_MLP_EDGE, prec, E, eb, None, None, plan.edge_s, plan.edge_r, None if g16 else d_e_out, None, None,
                     None if g16 else d_e, z1=ze, in0h=aggh, in1h=eh, d_gatherh=d_a1h, node_path=(plan, d_agg),
                     d_outh=d_e_outh if g16 else None, d_in1h=d_eh if g16 else None)


for g16 in (False, True):
    for _ in range(3):
        run(g16)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            run(g16)
        torch.cuda.synchronize()
    print("16-bit gradient streams" if g16 else "fp32 gradient streams")
    for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
        if e.device_time_total > 0 and "mlp_tc_bwd" in e.key:
            print(f"   {e.device_time_total / e.count / 1e3:7.3f} ms  {e.key[28:80]}")
rel = float((d_eh.float() - d_e).norm() / d_e.norm())
print("d_e 16-bit vs fp32 rel", rel, " d_x", float((d_xh.float() - d_x).norm() / d_x.norm()))

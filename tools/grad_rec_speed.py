"""The reference's gradient-reconstruction speed test (src/grad_rec_speed_test.py:118-168: node_based_WLSQ in a loop on
one mesh, scalar field, precomputed moments) through the mirror API on a B200 -- BASELINE.json config 2.

    python tools/grad_rec_speed.py [n_side ...]        # default: 80 (the 81x81 Poisson example size), 1000, 2000

Prints, per mesh: microseconds per node_based_WLSQ call (CUDA events over many calls), achieved GB/s against the
kernel's algorithmic bytes (nnz * 12 B stencil entries + N * (4 + 8) B field in / gradient out for one channel, all five
moments requested as in the reference call), and the relative L2 error on the analytic field of the reference's test."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import math
import torch
from gen_fvgn_steady_b200.mesh import synthetic_torch as ST
from gen_fvgn_steady_b200.FVMmodel.FVdiscretization.FVgrad import node_based_WLSQ

def measure(n, dev, peak, cuda_graph=False):
    """One mesh of n x n quad cells: us per node_based_WLSQ call (eager, or replayed as a one-call CUDA graph)."""
    mesh, _ = ST.make_case(n, kind="quad", bc="cavity", seed=0, device=dev)
    pos = mesh["node|pos"].float().contiguous()
    fx, se = mesh["face_node_x"].long(), mesh["support_edge"].long()
    A, B1, Bx = mesh["A_node_to_node"].float(), mesh["single_B_node_to_node"].float(), mesh["extra_B_node_to_node"].float()
    x, y = pos[:, 0].double(), pos[:, 1].double()
    # utilities.py:180-259 style analytic scalar: smooth exponential-trigonometric field
    phi = (1.0 + 0.01 * torch.sin(5 * x) + 0.01 * torch.cos(5 * y) + 0.01 * torch.sin(5 * x * y))
    gx = 0.05 * torch.cos(5 * x) + 0.05 * y * torch.cos(5 * x * y)
    gy = -0.05 * torch.sin(5 * y) + 0.05 * x * torch.cos(5 * x * y)
    phi32 = phi.float().reshape(-1, 1).contiguous()
    call = lambda: node_based_WLSQ(phi_node=phi32, edge_index=fx, extra_edge_index=se, mesh_pos=pos, order="2nd",
                                   precompute_Moments=[A, B1, Bx], rt_cond=False)
    with torch.no_grad():
        g = call()
        torch.cuda.synchronize()
        run = call
        if cuda_graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    call()
            torch.cuda.current_stream().wait_stream(side)
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                g = call()
            run = cg.replay
        reps = 2000 if n <= 200 else 100
        for _ in range(10):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    N, X = pos.shape[0], fx.shape[1]
    nnz = 2 * X + 2
    alg = nnz * (4 + 5 * 4) + N * (4 + 5 * 4)  # column index + five folded weights per entry; field in, 5 moments out
    ref = torch.stack([gx, gy], 1)
    err = float((g[:, 0, 0:2].double() - ref).norm() / ref.norm())
    return {"n_side": n, "nodes": N, "stencil_entries": nnz, "cuda_graph": cuda_graph, "us_per_call": round(us, 2),
            "nodes_per_s": round(N / us * 1e6), "alg_GBps": round(alg / us / 1e3, 1),
            "frac_of_measured_hbm": round(alg / us / 1e3 / peak, 3), "grad_rel_l2_err": err}


if __name__ == "__main__":
    sides = [int(a) for a in sys.argv[1:]] or [80, 1000, 2000]
    dev = torch.device("cuda")
    peak = 6545.3
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for n in sides:
        for cg in (False, True):
            print(json.dumps(measure(n, dev, peak, cg)))

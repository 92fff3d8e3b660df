"""Times fvgn_gemm_tf32 against torch (cuBLAS, TF32 allowed) for the Transolver block's shapes: python tools/gemm_bench.py [rows]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gen_fvgn_steady_b200 import ops

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
dev = torch.device("cuda")
torch.set_float32_matmul_precision("high")


def timeit(f, reps=5):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for O, I in ((256, 128), (128, 128), (128, 256)):
    x, w, dy = torch.randn(rows, I, device=dev), torch.randn(O, I, device=dev), torch.randn(rows, O, device=dev)
    b = torch.randn(O, device=dev)
    res = {}
    res["fwd  ours"] = timeit(lambda: ops.linear_fwd(x, w, b, tc=True))
    res["fwd  lib "] = timeit(lambda: ops.linear_fwd(x, w, b, tc=False))
    res["dgrad ours"] = timeit(lambda: ops.linear_dgrad(dy, w, tc=True))
    res["dgrad lib "] = timeit(lambda: ops.linear_dgrad(dy, w, tc=False))
    res["wgrad ours"] = timeit(lambda: ops.linear_wgrad(dy, x, tc=True))
    res["wgrad lib "] = timeit(lambda: ops.linear_wgrad(dy, x, tc=False))
    gb = rows * (O + I) * 4 / 1e6
    print(f"O={O} I={I} rows={rows}: " + "  ".join(f"{k} {v:.3f} ms ({gb / v:.0f} GB/s)" for k, v in res.items()), flush=True)

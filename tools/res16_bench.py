"""A/B of the forward kernels' residual source: fp32 stream (read 512 B, write 512 + 256 B per row) vs 16-bit shadow
(FVGN_MLP_RESIDUAL_FROM_SHADOW: read 256 B from the operand shadow, write 256 B)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gen_fvgn_steady_b200 import _lib, ops

dev = torch.device("cuda")
prec = "f16"
hdt = ops.HDTYPE[prec]
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
for mode, rows in (("NODE", 4_000_000), ("EDGE", 8_000_000)):
    nodes = rows // 2
    k1 = {"EDGE": 384, "NODE": 192}[mode]
    params = [rn(128, k1) / k1 ** 0.5, 0.1 * rn(128), rn(128, 128) / 128 ** 0.5, 0.1 * rn(128), rn(128, 128) / 128 ** 0.5,
              0.1 * rn(128), 1 + 0.1 * rn(128), 0.1 * rn(128)]
    if mode == "EDGE":
        s = torch.arange(rows, device=dev) // 2
        r = torch.clamp(s + torch.randint(1, 2000, (rows,), device=dev, generator=g), max=nodes - 1)
        s, r = s.to(torch.int32), r.to(torch.int32)
        in0h, in1 = rn(nodes, 128).to(hdt), rn(rows, 128)
        code = _lib.FVGN_MLP_EDGE
    else:
        s = r = None
        in0h, in1 = rn(rows, 64).to(hdt), rn(rows, 128)
        code = _lib.FVGN_MLP_NODE
    in1h = in1.to(hdt)
    z1 = ops.new_z1(code, prec, rows, in1)
    variants = {
        "fp32 residual stream": lambda: ops.mlp_forward(code, prec, rows, params, None, in1, s, r, want_out=False, want_res=True, z1=z1,
                                                        in0h=in0h, in1h=in1h, want_outh=mode == "EDGE", want_resh=True),
        "16-bit residual": lambda: ops.mlp_forward(code, prec, rows, params, None, None, s, r, want_out=False, want_res=False, z1=z1,
                                                   in0h=in0h, in1h=in1h, want_outh=mode == "EDGE", want_resh=True,
                                                   flags=_lib.FVGN_MLP_RESIDUAL_FROM_SHADOW),
        "no residual at all": lambda: ops.mlp_forward(code, prec, rows, params, None, None, s, r, want_out=False, want_res=False, z1=z1,
                                                      in0h=in0h, in1h=in1h, want_outh=True, want_resh=False),
    }
    for name, fn in variants.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(10):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        print(f"{mode} rows={rows}: {name:24s} {ev0.elapsed_time(ev1) / 10:.3f} ms")

"""Profiling driver: one forward + one backward of the fused edge / node MLP blocks at a given row count
(used under ncu; see profiles/)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gen_fvgn_steady_b200 import _lib, ops

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
mode = sys.argv[2] if len(sys.argv) > 2 else "EDGE"
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
dev = torch.device("cuda")
nodes = rows // 2
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)
k1 = {"EDGE": 384, "NODE": 192}[mode]
params = [rn(128, k1) / k1 ** 0.5, 0.1 * rn(128), rn(128, 128) / 128 ** 0.5, 0.1 * rn(128), rn(128, 128) / 128 ** 0.5, 0.1 * rn(128),
          1 + 0.1 * rn(128), 0.1 * rn(128)]
# mesh-like locality: senders sorted, receivers nearby
if mode == "EDGE":
    s = torch.arange(rows, device=dev) // 2
    r = torch.clamp(s + torch.randint(1, 2000, (rows,), device=dev, generator=g), max=nodes - 1)
    s, r = s.to(torch.int32), r.to(torch.int32)
    in0, in1 = rn(nodes, 128), rn(rows, 128)
    d_in0, d_in1 = torch.empty((rows, 256), device=dev), torch.empty((rows, 128), device=dev)
    d_gather = rn(nodes, 64)
    code = _lib.FVGN_MLP_EDGE
else:
    s = r = None
    in0, in1 = rn(rows, 64), rn(rows, 128)
    d_in0, d_in1 = torch.empty((rows, 64), device=dev), torch.empty((rows, 128), device=dev)
    d_gather = None
    code = _lib.FVGN_MLP_NODE
d_out = rn(rows, 128)
z1 = ops.new_z1(code, prec, rows, in0)
bf = ops.is_tc(prec)
hdt = ops.HDTYPE.get(prec)
in0h, in1h = (ops.shadow(in0, dtype=hdt), ops.shadow(in1, dtype=hdt)) if bf else (None, None)
d_in0h = torch.empty((rows, 256), dtype=hdt, device=dev) if (bf and mode == "EDGE") else None
res16 = bf and os.environ.get("FVGN_TC_PROFILE_RES16", "1") != "0"   # 16-bit latent streams (the models' default)
fwd = lambda: ops.mlp_forward(code, prec, rows, params, in0, None if res16 else in1, s, r, want_out=not bf, want_res=not res16,
                              z1=z1, in0h=in0h, in1h=in1h, want_outh=bf and mode == "EDGE", want_resh=bf,
                              flags=_lib.FVGN_MLP_RESIDUAL_FROM_SHADOW if res16 else 0)
bwd = lambda: ops.mlp_backward(code, prec, rows, params, in0, in1, s, r, d_out, d_gather, None if d_in0h is not None else d_in0,
                               d_in1, z1=z1, in0h=in0h, in1h=in1h, d_in0h=d_in0h)
for _ in range(2):
    fwd()
    bwd()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
reps = 5
tf = tb = 0.0
for _ in range(reps):
    ev[0].record()
    fwd()
    ev[1].record()
    bwd()
    ev[2].record()
    torch.cuda.synchronize()
    tf += ev[0].elapsed_time(ev[1]) / reps
    tb += ev[1].elapsed_time(ev[2]) / reps
print(f"{mode} rows={rows} {prec}: fwd {tf:.3f} ms, bwd {tb:.3f} ms (mean of {reps})")

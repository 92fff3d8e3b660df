"""Profiling driver: forward + backward of one Transolver block (csrc/transolver.cu + library GEMMs) on N node rows in
B graphs; prints CUDA-event timings, used under ncu for profiles/.   python tools/ts_profile.py [rows] [graphs] [reps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gen_fvgn_steady_b200.FVMmodel.Models.GraphTransolver.GraphTransolver import Transolver_block

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
graphs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda")
torch.set_float32_matmul_precision("high")   # the reference's TF32 setting (src/pre_train_Adam.py:29) for the library GEMMs
torch.manual_seed(0)
blk = Transolver_block(num_heads=8, hidden_dim=128, dropout=0, act="gelu", mlp_ratio=2, slice_num=32).to(dev)
blk.precision = "bf16"
x = torch.randn(rows, 128, device=dev, requires_grad=True)
emb = torch.randn(rows, 128, device=dev, requires_grad=True)
batch = (torch.arange(rows, device=dev) * graphs // rows).to(torch.int64)
cot = torch.randn(rows, 128, device=dev)


def step():
    x.grad = emb.grad = None
    for p in blk.parameters():
        p.grad = None
    t = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t[0].record()
    out = blk(x, batch, embedding=emb)
    t[1].record()
    out.backward(cot)
    t[2].record()
    torch.cuda.synchronize()
    return t[0].elapsed_time(t[1]), t[1].elapsed_time(t[2])


for _ in range(2):
    step()
ts = [step() for _ in range(reps)]
f = sum(a for a, _ in ts) / reps
b = sum(b for _, b in ts) / reps
print(f"Transolver_block rows={rows} graphs={graphs}: fwd {f:.3f} ms, bwd {b:.3f} ms (mean of {reps}); "
      f"{(f + b) * 1e6 / rows:.2f} ns per node")

"""Times the CSR reductions on a synthetic quad-mesh graph: python tools/reduce_bench.py [n_side]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gen_fvgn_steady_b200 import _lib, ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = torch.device("cuda")
# node grid (n+1)^2, edges horizontal + vertical, CSR of the symmetric adjacency in (senders, receivers) entry order
m = n + 1
idx = torch.arange(m * m, device=dev).view(m, m)
s = torch.cat([idx[:, :-1].reshape(-1), idx[:-1, :].reshape(-1)])
r = torch.cat([idx[:, 1:].reshape(-1), idx[1:, :].reshape(-1)])
E, N = s.numel(), m * m
dst = torch.cat([s, r]); nb = torch.cat([r, s]); code = torch.cat([torch.arange(E, device=dev) * 2, torch.arange(E, device=dev) * 2 + 1])
order = torch.sort(dst, stable=True).indices
ptr = torch.zeros(N + 1, dtype=torch.int32, device=dev); ptr[1:] = torch.cumsum(torch.bincount(dst, minlength=N), 0).int()
nbr = nb[order].int().contiguous(); cod = code[order].int().contiguous()
st = _lib.stream_ptr(dev)
def timeit(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for W in (128, 64):
    x = torch.randn((N, W), device=dev); xh = x.bfloat16(); out = torch.empty_like(x); outh = torch.empty_like(xh)
    e = torch.randn((E, 2 * W), device=dev); eh = e.bfloat16()
    for name, extra in (("chunk", 0), ("simple", _lib.FVGN_ADJ_SIMPLE_KERNEL)):
        for fl_name, fl in (("none", 0), ("acc", _lib.FVGN_ADJ_ACCUMULATE), ("src", _lib.FVGN_ADJ_DIV_SRC_BY_DEG)):
            t = timeit(lambda: _lib.call("fvgn_adj_reduce", _lib.fptr(x), _lib.iptr(ptr), _lib.iptr(nbr), _lib.fptr(out), N, W, fl | extra, st))
            print(f"adj W={W} f32->f32 {name:6s} {fl_name:4s}: {t:.3f} ms  ({N * W * 4 * (3 if fl_name == 'acc' else 2) / t / 1e6:.0f} GB/s alg)")
        t = timeit(lambda: _lib.call("fvgn_adj_reduce_t", _lib.ptr(xh), 1, _lib.iptr(ptr), _lib.iptr(nbr), _lib.ptr(outh), 1, N, W, extra, st))
        print(f"adj W={W} bf16->bf16 {name:6s}: {t:.3f} ms")
    t = timeit(lambda: _lib.call("fvgn_inc_reduce", _lib.fptr(e), _lib.iptr(ptr), _lib.iptr(cod), _lib.fptr(out), N, W, st))
    print(f"inc W={W} f32->f32: {t:.3f} ms")
    t = timeit(lambda: _lib.call("fvgn_inc_reduce_t", _lib.ptr(eh), 1, _lib.iptr(ptr), _lib.iptr(cod), _lib.ptr(outh), 1, N, W, st))
    print(f"inc W={W} bf16->bf16: {t:.3f} ms")

"""Builds variants of csrc/segreduce.cu (macro sweeps) into tools/_variants/ and, on a GPU, times the six typed CSR
reductions of one GnBlock step on a quad-mesh graph with each of them.

    python tools/reduce_variants.py build            (container: nvcc only)
    python tools/reduce_variants.py run [n_side]     (GPU box)
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "tools", "_variants")
VARIANTS = {
    "R1_mb3": ["-DFVGN_PIPE_ROWS=1", "-DFVGN_PIPE_MINB=3"],
    "R1_mb5": ["-DFVGN_PIPE_ROWS=1", "-DFVGN_PIPE_MINB=5"],
    "R1_mb8": ["-DFVGN_PIPE_ROWS=1", "-DFVGN_PIPE_MINB=8"],
    "R2_mb4": ["-DFVGN_PIPE_ROWS=2", "-DFVGN_PIPE_MINB=4"],
    "R2_mb5": ["-DFVGN_PIPE_ROWS=2", "-DFVGN_PIPE_MINB=5"],
    "R2_mb6": ["-DFVGN_PIPE_ROWS=2", "-DFVGN_PIPE_MINB=6"],
}


def build():
    os.makedirs(VDIR, exist_ok=True)
    src = os.path.join(ROOT, "gen_fvgn_steady_b200", "csrc", "segreduce.cu")
    procs = []
    for name, flags in VARIANTS.items():
        out = os.path.join(VDIR, f"libseg_{name}.so")
        cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler",
               "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-Xptxas", "-v"] + flags + [src, "-o", out]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        log = p.communicate()[0]
        regs = [l for l in log.splitlines() if "registers" in l]
        print(name, "rc", p.returncode, "max regs", max(int(l.split("Used ")[1].split()[0]) for l in regs) if regs else None)
        if p.returncode:
            print(log[-3000:])


def run(n):
    import torch
    dev = torch.device("cuda")
    m = n + 1
    idx = torch.arange(m * m, device=dev).view(m, m)
    s = torch.cat([idx[:, :-1].reshape(-1), idx[:-1, :].reshape(-1)])
    r = torch.cat([idx[:, 1:].reshape(-1), idx[1:, :].reshape(-1)])
    E, N = s.numel(), m * m
    dst = torch.cat([s, r]); nb = torch.cat([r, s])
    code = torch.cat([torch.arange(E, device=dev) * 2, torch.arange(E, device=dev) * 2 + 1])
    order = torch.sort(dst, stable=True).indices
    ptr = torch.zeros(N + 1, dtype=torch.int32, device=dev); ptr[1:] = torch.cumsum(torch.bincount(dst, minlength=N), 0).int()
    nbr = nb[order].int().contiguous(); cod = code[order].int().contiguous()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    BF, F32 = 1, 0

    def timeit(f, reps=10):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            f()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    x128 = torch.randn((N, 128), device=dev).bfloat16(); x64 = torch.randn((N, 64), device=dev).bfloat16()
    e128 = torch.randn((E, 128), device=dev).bfloat16(); e256 = torch.randn((E, 256), device=dev).bfloat16()
    o128 = torch.empty_like(x128); o64 = torch.empty_like(x64); of = torch.zeros((N, 128), device=dev)
    idxb = (N + 1) * 4 + 2 * E * 4
    cases = [  # name, call, algorithmic bytes
        ("adj128 h->h      ", lambda L: L.fvgn_adj_reduce_t(P(x128), BF, P(ptr), P(nbr), P(o128), BF, ctypes.c_int64(N), 128, 0, st), N * 512 + idxb),
        ("inc64  h->h      ", lambda L: L.fvgn_inc_reduce_t(P(e128), BF, P(ptr), P(cod), P(o64), BF, ctypes.c_int64(N), 64, st), E * 256 + N * 128 + idxb),
        ("adj64  h->h dst  ", lambda L: L.fvgn_adj_reduce_t(P(x64), BF, P(ptr), P(nbr), P(o64), BF, ctypes.c_int64(N), 64, 2, st), N * 256 + idxb),
        ("adj64  h->h src  ", lambda L: L.fvgn_adj_reduce_t(P(x64), BF, P(ptr), P(nbr), P(o64), BF, ctypes.c_int64(N), 64, 4, st), N * 256 + idxb),
        ("inc128 h->h      ", lambda L: L.fvgn_inc_reduce_t(P(e256), BF, P(ptr), P(cod), P(o128), BF, ctypes.c_int64(N), 128, st), E * 512 + N * 256 + idxb),
        ("adj128 h->f32 acc", lambda L: L.fvgn_adj_reduce_t(P(x128), BF, P(ptr), P(nbr), P(of), F32, ctypes.c_int64(N), 128, 1, st), N * (256 + 1024) + idxb),
    ]
    libs = sorted(f for f in os.listdir(VDIR) if f.endswith(".so"))
    print(f"N={N} E={E}; ms (GB/s algorithmic)")
    print("variant   " + " | ".join(c[0] for c in cases) + " | total ms")
    for lf in libs:
        L = ctypes.CDLL(os.path.join(VDIR, lf))
        row, tot = [], 0.0
        for name, f, nbytes in cases:
            rc = f(L)
            assert rc == 0, (lf, name, rc)
            t = timeit(lambda: f(L))
            tot += t
            row.append(f"{t:6.3f} ({nbytes / t / 1e6:5.0f})   ")
        print(f"{lf[7:-3]:9s} " + " | ".join(row) + f" | {tot:.3f}")


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 2000)

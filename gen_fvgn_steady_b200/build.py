"""Builds libfvgn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m gen_fvgn_steady_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfvgn_b200.so")
SOURCES = ["segreduce.cu", "mlp_simt.cu", "mlp_tc.cu", "mlp_tc_bwd.cu", "mlp_api.cu", "misc.cu", "fv.cu", "transolver.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def _newer(lib, deps):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "fvgn_b200.h")]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not force and not _newer(LIB, deps):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + os.environ.get("FVGN_EXTRA_NVCC_FLAGS", "").split()
    cmd = [nvcc] + flags + ["-o", LIB] + srcs + ["-lcuda"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libfvgn_b200.so (see %s)" % log)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

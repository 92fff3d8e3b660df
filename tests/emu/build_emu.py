"""TEST INFRASTRUCTURE ONLY: builds tests/emu/libfvgn_emu.so = the SIMT kernels of
gen_fvgn_steady_b200/csrc compiled for the HOST with g++ over the CPU SIMT emulator (cuda_emu.h).
Lets the GPU-less build container check kernel index math / numerics against the oracle.  The
product never loads this library (gen_fvgn_steady_b200/_lib.py only ever opens libfvgn_b200.so)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "gen_fvgn_steady_b200", "csrc")
LIB = os.path.join(HERE, "libfvgn_emu.so")
SOURCES = ["segreduce.cu", "mlp_simt.cu", "mlp_api.cu", "misc.cu", "fv.cu", "transolver.cu", "plan_build.cu"]


def build(force=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "cuda_emu.h"),
                   os.path.join(ROOT, "include", "fvgn_b200.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    cmd = ["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-DFVGN_EMU", "-ffp-contract=off", "-I", HERE,
           "-Wno-unused-parameter", "-o", LIB]
    for s in srcs:
        cmd += ["-x", "c++", s]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emu build failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))

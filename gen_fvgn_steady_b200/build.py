"""Builds libfvgn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m gen_fvgn_steady_b200.build [--force]

Every .cu is compiled to its own object (in parallel, only when it or a header changed) and the objects are linked into
the shared library; objects live under csrc/_obj (git-ignored).
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libfvgn_b200.so")
SOURCES = ["segreduce.cu", "mlp_simt.cu", "mlp_tc.cu", "mlp_tc_bwd.cu", "mlp_tc_bwd_node.cu", "mlp_api.cu", "misc.cu", "fv.cu", "transolver.cu",
           "plan_build.cu", "gemm_tf32.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    return hs + [os.path.join(HERE, "..", "include", "fvgn_b200.h")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("FVGN_EXTRA_NVCC_FLAGS", "").split()
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "flags.txt")
    flag_str = " ".join(NVCC_FLAGS + extra)
    if not os.path.exists(stamp) or open(stamp).read() != flag_str:
        force = True
    hdrs = _headers()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    jobs = []
    for s in srcs:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if force or _stale(obj, [src] + hdrs):
            jobs.append((s, [nvcc] + NVCC_FLAGS + extra + ["-c", src, "-o", obj]))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    if not jobs and not _stale(LIB, objs):
        return LIB
    logs = {}

    def run(job):
        name, cmd = job
        res = subprocess.run(cmd, capture_output=True, text=True)
        logs[name] = " ".join(cmd) + "\n" + res.stdout + res.stderr
        return name, res.returncode

    failed = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4) or 1) as ex:
        for name, rc in ex.map(run, jobs):
            if rc != 0:
                failed.append(name)
    link = [nvcc, "-shared", "-o", LIB] + objs + ["-lcuda"]
    if not failed:
        res = subprocess.run(link, capture_output=True, text=True)
        logs["link"] = " ".join(link) + "\n" + res.stdout + res.stderr
        if res.returncode != 0:
            failed.append("link")
    log = os.path.join(HERE, "build.log")
    mode = "w" if force or not os.path.exists(log) else "a"
    with open(log, mode) as f:
        for k in sorted(logs):
            f.write(logs[k] + "\n")
    if failed:
        for k in failed:
            sys.stderr.write(logs.get(k, ""))
        raise RuntimeError("nvcc failed building libfvgn_b200.so (%s; see %s)" % (", ".join(failed), log))
    with open(stamp, "w") as f:
        f.write(flag_str)
    if verbose:
        for k in sorted(logs):
            print(logs[k])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

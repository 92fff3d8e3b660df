"""Cycle breakdown of kernel A's per-tile chain (debug build: FVGN_EXTRA_NVCC_FLAGS=-DFVGN_TIMING python -m
gen_fvgn_steady_b200.build --force).  Usage on the GPU box: python tools/phase_profile_a.py [rows] [mode]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gen_fvgn_steady_b200 import _lib
import subprocess
rows = sys.argv[1] if len(sys.argv) > 1 else "2000000"
mode = sys.argv[2] if len(sys.argv) > 2 else "EDGE"
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_ulonglong * 16)()
import runpy
sys.argv = ["tc_profile.py", rows, mode, sys.argv[3] if len(sys.argv) > 3 else "f16"]
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tc_profile.py"), run_name="__main__")
torch.cuda.synchronize()
lib2 = _lib.load()
fn = ctypes.CDLL(_lib.LIB_PATH).fvgn_debug_profile_a
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
rc = fn(buf, 0)
names = ["wait Z1", "E1", "wait R2", "E2", "wait R3", "wait dO", "E3+db3", "wait M3", "E4+db2", "wait M4", "E5+db1"]
bf = (ctypes.c_ulonglong * 16)()
ff = ctypes.CDLL(_lib.LIB_PATH).fvgn_debug_profile_f
ff.argtypes = [ctypes.c_void_p, ctypes.c_int]
ff(bf, 0)
nf = ["wait L1", "epi L1", "wait L2", "epi L2", "wait L3", "LN stats", "output"]
tf = max(int(bf[15]), 1)
totf = sum(int(bf[i]) for i in range(7))
print(f"forward (pipeline 0 of CTA 0): tiles={tf} cycles/tile={totf / tf:.0f}")
for i, n in enumerate(nf):
    print(f"  {n:10s} {int(bf[i]) / tf:8.0f} cyc  {int(bf[i]) / max(totf, 1) * 100:5.1f}%")
tiles = max(int(buf[15]), 1)
tot = sum(int(buf[i]) for i in range(11))
print(f"rc={rc} tiles(CTA0, all launches)={tiles} cycles/tile={tot / tiles:.0f}")
for i, n in enumerate(names):
    print(f"  {n:10s} {int(buf[i]) / tiles:8.0f} cyc  {int(buf[i]) / max(tot, 1) * 100:5.1f}%")

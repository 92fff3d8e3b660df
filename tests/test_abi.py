"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/fvgn_b200.h declares."""
import ctypes
import os
import subprocess

from gen_fvgn_steady_b200 import _lib, build


def test_library_exports_every_declared_symbol():
    lib = build.build()
    h = ctypes.CDLL(lib)
    for name in _lib.PROTOS:
        assert hasattr(h, name), name
    assert h.fvgn_version() & 3 == 3
    assert h.fvgn_mlp_param_count(_lib.FVGN_MLP_EDGE) == 128 * 384 + 128 + 2 * (128 * 128 + 128) + 256


def test_header_structs_match_c_layout(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include "fvgn_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu\\n", sizeof(fvgn_mlp_desc), sizeof(fvgn_fv_desc));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.dirname(_lib.HEADER), str(src), "-o", str(exe)])
    a, b = subprocess.check_output([str(exe)]).split()
    assert int(a) == ctypes.sizeof(_lib.MlpDesc) and int(b) == ctypes.sizeof(_lib.FvDesc)


def test_sass_is_sm100a():
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out

"""ctypes binding of libfvgn_b200.so (the C-ABI declared in include/fvgn_b200.h).

The prototypes and the two descriptor structs are parsed from the header itself, so the Python
side cannot drift from the C side.  There is NO CPU fallback: if the CUDA library is missing or a
tensor is not on a CUDA device, the call raises.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "fvgn_b200.h")
LIB_PATH = os.path.join(_HERE, "libfvgn_b200.so")

_SCALARS = {"int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "int": ctypes.c_int, "float": ctypes.c_float}


def _strip_comments(src):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", " ", src)


def _ctype_of(tp, structs):
    tp = tp.replace("const", " ").strip()
    ptr = tp.endswith("*")
    base = tp.rstrip("*").strip()
    if ptr:
        if base in structs:
            return ctypes.POINTER(structs[base])
        return ctypes.c_void_p
    return _SCALARS[base]


def parse_header(path=HEADER):
    """-> (defines: {name: int}, structs: {name: ctypes.Structure}, protos: {name: (restype, [argtypes], [argnames])})"""
    src = _strip_comments(open(path).read())
    defines = {}
    for m in re.finditer(r"#define\s+(FVGN_\w+)\s+\(?(-?\d+)\)?\s*$", src, flags=re.M):
        defines[m.group(1)] = int(m.group(2))
    structs = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        fields = []
        for stmt in m.group(2).split(";"):
            stmt = " ".join(stmt.split())
            if not stmt:
                continue
            mm = re.match(r"^((?:const\s+)?\w+\s*\*?)\s*(.+)$", stmt)
            tp, names = mm.group(1), mm.group(2)
            for nm in names.split(","):
                nm = nm.strip()
                t = tp
                if nm.startswith("*"):
                    t, nm = tp + "*", nm[1:].strip()
                fields.append((nm, _ctype_of(t, structs)))
        cls = type(m.group(3), (ctypes.Structure,), {"_fields_": fields})
        structs[m.group(3)] = cls
    protos = {}
    body = re.sub(r"typedef\s+struct\s+\w+\s*\{.*?\}\s*\w+\s*;", " ", src, flags=re.S)
    for m in re.finditer(r"\b(int|int32_t|int64_t)\s+(fvgn_\w+)\s*\(([^)]*)\)\s*;", body, flags=re.S):
        args, names = [], []
        arglist = " ".join(m.group(3).split())
        if arglist and arglist != "void":
            for a in arglist.split(","):
                a = a.strip()
                mm = re.match(r"^(.*?[\s\*])(\w+)$", a)
                args.append(_ctype_of(mm.group(1), structs))
                names.append(mm.group(2))
        protos[m.group(2)] = (_SCALARS[m.group(1)], args, names)
    return defines, structs, protos


DEFINES, STRUCTS, PROTOS = parse_header()
globals().update(DEFINES)
MlpDesc = STRUCTS["fvgn_mlp_desc"]
FvDesc = STRUCTS["fvgn_fv_desc"]

_ERRORS = {v: k for k, v in DEFINES.items() if k.startswith("FVGN_ERR_")}

_handle = None
_allow_host_tensors = False  # flipped only by tests that drive the kernels through the CPU SIMT emulator


def _bind(handle):
    for name, (res, args, _) in PROTOS.items():
        fn = getattr(handle, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return handle


def load():
    """Open libfvgn_b200.so (building it with nvcc first if the in-tree copy is missing or stale)."""
    global _handle
    if _handle is not None:
        return _handle
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    _handle = _bind(ctypes.CDLL(LIB_PATH))
    return _handle


def _set_library_for_tests(handle, allow_host_tensors):
    """TEST HOOK (tests/emu only): run the host logic against the CPU-emulated kernels.  Never called by the product."""
    global _handle, _allow_host_tensors
    _handle = _bind(handle) if handle is not None else None
    _allow_host_tensors = bool(allow_host_tensors)


def stream_ptr(device=None):
    if _allow_host_tensors:
        return None
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t, dtype=None, allow_none=False):
    """Device pointer of a contiguous tensor (None -> NULL)."""
    if t is None:
        if allow_none:
            return None
        raise RuntimeError("fvgn_b200: required tensor is None")
    if not t.is_cuda and not _allow_host_tensors:
        raise RuntimeError("fvgn_b200: tensor is not on a CUDA device -- this package has no CPU path")
    if not t.is_contiguous():
        raise RuntimeError("fvgn_b200: tensor must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"fvgn_b200: expected dtype {dtype}, got {t.dtype}")
    return t.data_ptr()


def fptr(t, allow_none=False):
    return ptr(t, torch.float32, allow_none)


def iptr(t, allow_none=False):
    return ptr(t, torch.int32, allow_none)


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"fvgn_b200: {what} failed with {_ERRORS.get(rc, rc)}")


# kernels launched per C-ABI call (for bench.py's gpu_launches claim)
# (fvgn_mlp_backward: kernel A + kernel B + the deterministic partial reduction in bf16 mode, kernel + reduction in fp32 mode:
#  counted as 2, a lower bound)
LAUNCHES_PER_CALL = {"fvgn_mlp_backward": 2, "fvgn_fv_backward": 2, "fvgn_fv_outputs": 2}
launch_count = 0


def call(name, *args):
    global launch_count
    rc = getattr(load(), name)(*args)
    check(rc, name)
    launch_count += LAUNCHES_PER_CALL.get(name, 1)

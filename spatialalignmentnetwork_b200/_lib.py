"""ctypes binding of ``libsan_b200.so`` (the C ABI declared in ``include/san_b200.h``).

The prototypes are parsed from the header itself so the binding cannot drift
from the declared ABI.  There is no fallback: if the shared library is missing
(``python __graft_entry__.py build`` / ``make -C spatialalignmentnetwork_b200/csrc``)
importing any op raises.
"""
import ctypes
import os
import re

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_PKG), "include", "san_b200.h")
LIB_PATH = os.path.join(_PKG, "libsan_b200.so")

_CTYPES = {
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "size_t": ctypes.c_size_t,
    "const char*": ctypes.c_char_p,
}


def _ctype(t):
    t = " ".join(t.replace("*", " * ").split()).replace(" *", "*")
    if t.endswith("*") and t != "const char*":
        return ctypes.c_void_p
    return _CTYPES[t]


def parse_header(path=HEADER):
    """-> {name: (restype_str, [(type_str, arg_name), ...])} for every prototype."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith("#"))
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(san_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        arglist = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)$", a)
                arglist.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = (ret, arglist)
    return protos


PROTOS = parse_header()
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not built: run `python __graft_entry__.py build` "
                "(there is no CPU / PyTorch fallback for the san_b200 kernels)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (ret, args) in PROTOS.items():
            fn = getattr(L, name)
            fn.restype = _ctype(ret)
            fn.argtypes = [_ctype(t) for t, _ in args]
        _lib = L
    return _lib


def _ptr(x, name, arg):
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError(f"{name}: argument '{arg}' must be a CUDA tensor (no CPU fallback)")
        if not x.is_contiguous():
            raise RuntimeError(f"{name}: argument '{arg}' must be contiguous")
        if x.is_conj() or x.is_neg():
            raise RuntimeError(f"{name}: argument '{arg}' is a lazy conj / neg view; resolve it before the call")
        return x.data_ptr()
    return int(x)


def call(name, *args):
    """Call ``san_<name>`` on the current CUDA stream (last ABI argument) and raise on error."""
    full = "san_" + name
    L = lib()
    _, protoargs = PROTOS[full]
    conv = []
    if len(args) != len(protoargs) - 1:
        raise TypeError(f"{full}: expected {len(protoargs) - 1} arguments (+stream), got {len(args)}")
    for a, (t, an) in zip(args, protoargs):
        conv.append(_ptr(a, full, an) if "*" in t else a)
    conv.append(torch.cuda.current_stream().cuda_stream)
    if _profile is not None:
        # bench.py's instrumented step: bracket the launch with CUDA events on the launching stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(L, full)(*conv)
        e1.record()
        _profile.append((name, tuple(None if a is None else (a if isinstance(a, (int, float)) else 0) for a in args),
                         e0, e1))
    else:
        rc = getattr(L, full)(*conv)
    if rc != 0:
        raise RuntimeError(f"{full} failed ({rc}): {L.san_last_error().decode()}")


_profile = None


def profile_begin():
    """Start recording (op name, scalar args, start/stop CUDA events) for every C-ABI call."""
    global _profile
    _profile = []


def profile_end():
    """Stop recording; synchronise and return [(name, args, milliseconds)]."""
    global _profile
    recs, _profile = _profile, None
    torch.cuda.synchronize()
    return [(n, a, e0.elapsed_time(e1)) for n, a, e0, e1 in recs]


def launch_count():
    return int(lib().san_launch_count())

"""ctypes binding of libstego_b200.so — the C-ABI drop-in boundary (include/stego_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  Nothing here imports the oracle.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstego_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "stego_b200.h")

_lib = None
replayed_launches = 0  # kernels launched through CUDA-graph replays (not seen by stego_launch_count)

_CTYPES = {
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "long long": ctypes.c_longlong,
    "void*": ctypes.c_void_p,
    "const void*": ctypes.c_void_p,
    "const float*": ctypes.c_void_p,
    "float*": ctypes.c_void_p,
    "const long long*": ctypes.c_void_p,
    "long long*": ctypes.c_void_p,
    "const int*": ctypes.c_void_p,
    "int*": ctypes.c_void_p,
    "const char*": ctypes.c_char_p,
    "unsigned char*": ctypes.c_void_p,
    "const unsigned char*": ctypes.c_void_p,
    "const double*": ctypes.c_void_p,
}


def header_prototypes() -> Dict[str, Tuple[str, List[str]]]:
    """Parse `STEGO_API <ret> name(args);` declarations from the public header."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"^\s*#[^\n]*", "", text, flags=re.M)  # preprocessor lines
    protos: Dict[str, Tuple[str, List[str]]] = {}
    for m in re.finditer(r"STEGO_API\s+([\w\s\*]+?)\s*\b(stego_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        arg_types: List[str] = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                t = re.sub(r"\s*\b\w+$", "", a).strip()  # drop the parameter name
                t = t.replace(" *", "*")
                arg_types.append(t)
        protos[name] = (ret.replace(" *", "*"), arg_types)
    return protos


def load():
    """Load the shared library (once) and attach argtypes from the header."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"stego_b200: {LIB_PATH} is missing. Build it with `python -m stego_b200.build` "
            "(or __graft_entry__.build()). There is no CPU / eager fallback for the hot path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (ret, args) in header_prototypes().items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = _CTYPES[ret]
        fn.argtypes = [_CTYPES[a] for a in args]
    _lib = lib
    return lib


def launch_count() -> int:
    """Kernels this library launched in this process: direct launches + launches inside replayed CUDA graphs."""
    return int(load().stego_launch_count()) + replayed_launches


def last_error() -> str:
    return load().stego_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"stego_b200.{what} failed with status {rc}: {last_error()}")


def ptr(t) -> int:
    """Raw device pointer of a tensor (0 for None)."""
    if t is None:
        return 0
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("stego_b200: hot-path tensors must live on a CUDA device (no CPU fallback)")

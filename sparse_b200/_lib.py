"""ctypes binding of libsparse_b200.so (the C ABI declared in include/sparse_b200.h).

There is no CPU fallback: if the library is missing, or no CUDA device is
visible when an operation needs one, a RuntimeError is raised -- loudly.
"""
from __future__ import annotations

import ctypes
import os
import re

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libsparse_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "sparse_b200.h")

# dtype codes (b2s_dtype)
F32, F64, I32, I64, BOOL = 0, 1, 2, 3, 4

_lib = None


class B2SError(RuntimeError):
    """A call into libsparse_b200 failed (message from b2s_last_error)."""


def header_symbols(path: str = HEADER_PATH):
    """Names of every function declared in include/sparse_b200.h."""
    with open(path) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", text)))


def load():
    """Load the shared library (does not need a GPU; symbol resolution only)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"sparse_b200: {LIB_PATH} is missing. Build it with `python -m sparse_b200._build` "
            "(or __graft_entry__.build()). There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.b2s_last_error.restype = ctypes.c_char_p
    lib.b2s_launch_count.restype = ctypes.c_int64
    for name in header_symbols():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        if name not in ("b2s_last_error", "b2s_launch_count"):
            fn.restype = ctypes.c_int
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().b2s_last_error().decode("utf-8", "replace")
        raise B2SError(f"{what or 'libsparse_b200'} failed (status {rc}): {msg}")


def require_device():
    """Raise unless a CUDA device is usable. No silent fallback."""
    lib = load()
    n = ctypes.c_int(0)
    rc = lib.b2s_device_count(ctypes.byref(n))
    if rc != 0 or n.value < 1:
        raise RuntimeError(
            "sparse_b200: no CUDA device visible. This package runs its hot path on a B200 only; "
            "there is no CPU fallback."
        )
    return n.value


def launch_count() -> int:
    return int(load().b2s_launch_count())


# small helpers for argument marshalling
def vp(x):
    return ctypes.c_void_p(int(x))


def i64(x):
    return ctypes.c_int64(int(x))


def i32(x):
    return ctypes.c_int(int(x))

"""Device-memory plumbing: torch owns HBM allocations and streams, nothing else.

Containers keep their arrays as ``torch`` CUDA tensors (moved to HBM once) and
hand raw device pointers to the C ABI.  NumPy mirrors are materialised lazily
when a caller reads ``.coords`` / ``.data`` / ``.indices`` / ``.indptr``.
"""
from __future__ import annotations

import numpy as np

from . import _lib

_torch = None
# Test hook ONLY (tests/_mock_kernels.py): lets the host-side logic be exercised on a box without a GPU by
# pairing CPU torch tensors with a NumPy mock of `_kernels`.  Never set by the product; with it False (always,
# outside the test-suite) every operation requires a CUDA device.
_TEST_CPU = False


def torch():
    global _torch
    if _torch is None:
        import torch as _t

        _torch = _t
    return _torch


_NP2T = None


def _np2t():
    global _NP2T
    if _NP2T is None:
        t = torch()
        _NP2T = {
            np.dtype("float32"): t.float32,
            np.dtype("float64"): t.float64,
            np.dtype("int32"): t.int32,
            np.dtype("int64"): t.int64,
            np.dtype("int16"): t.int16,
            np.dtype("int8"): t.int8,
            np.dtype("uint8"): t.uint8,
            np.dtype("bool"): t.bool,
        }
        # storage-only dtypes: containers hold them and every structural (element-size generic) kernel moves them;
        # arithmetic on them raises TypeError in the op layer (outside the CUDA dtype matrix)
        for name in ("uint16", "uint32", "uint64", "float16", "complex64", "complex128"):
            if hasattr(t, name):
                _NP2T[np.dtype(name)] = getattr(t, name)
    return _NP2T


_CODES = {
    np.dtype("float32"): _lib.F32,
    np.dtype("float64"): _lib.F64,
    np.dtype("int32"): _lib.I32,
    np.dtype("int64"): _lib.I64,
    np.dtype("bool"): _lib.BOOL,
}


def dtype_code(dt) -> int:
    dt = np.dtype(dt)
    if dt not in _CODES:
        raise TypeError(f"sparse_b200: dtype {dt} is outside the supported CUDA dtype matrix "
                        f"({', '.join(str(k) for k in _CODES)})")
    return _CODES[dt]


# b2s_cast alone also takes the storage-only integer widths (include/sparse_b200.h: B2S_I8 .. B2S_U64)
_CAST_CODES = {**_CODES, np.dtype("int8"): 5, np.dtype("int16"): 6, np.dtype("uint8"): 7, np.dtype("uint16"): 8,
               np.dtype("uint32"): 9, np.dtype("uint64"): 10}


def cast_code(dt) -> int:
    dt = np.dtype(dt)
    if dt not in _CAST_CODES:
        raise TypeError(f"sparse_b200: no device cast to / from dtype {dt} "
                        f"({', '.join(str(k) for k in _CAST_CODES)})")
    return _CAST_CODES[dt]


def torch_dtype(dt):
    dt = np.dtype(dt)
    m = _np2t()
    if dt not in m:
        raise TypeError(f"sparse_b200: dtype {dt} has no device representation")
    return m[dt]


def np_dtype(t) -> np.dtype:
    for k, v in _np2t().items():
        if v == t.dtype:
            return k
    raise TypeError(f"unsupported torch dtype {t.dtype}")


def have_device() -> bool:
    if _TEST_CPU:
        return True
    try:
        return bool(torch().cuda.is_available())
    except Exception:  # pragma: no cover
        return False


def require_device():
    if not have_device():
        raise RuntimeError(
            "sparse_b200: no CUDA device available. The hot path runs only on a B200 (sm_100a); "
            "there is no CPU fallback."
        )
    if not _TEST_CPU:
        _lib.load()


def device():
    t = torch()
    if _TEST_CPU:
        return t.device("cpu")
    return t.device("cuda", t.cuda.current_device())


def stream_ptr() -> int:
    if _TEST_CPU:
        return 0
    return int(torch().cuda.current_stream().cuda_stream)


def is_device_tensor(x) -> bool:
    import sys

    t = _torch or sys.modules.get("torch")  # a tensor can only exist if torch is already imported
    return t is not None and isinstance(x, t.Tensor) and (x.is_cuda or _TEST_CPU)


def upload(arr, dtype=None):
    """Host ndarray -> contiguous CUDA tensor (one H2D copy; async if `arr` is pinned)."""
    require_device()
    t = torch()
    a = np.asarray(arr)
    if dtype is not None and a.dtype != np.dtype(dtype):
        a = a.astype(dtype)
    torch_dtype(a.dtype)
    a = np.ascontiguousarray(a)
    if not a.flags.writeable:
        a = a.copy()
    return t.from_numpy(a).to(device(), non_blocking=True)


def device_index_dtype(dt) -> np.dtype:
    """Index arrays live on the device as int32 or int64 only (what the kernels take); narrower or unsigned index
    dtypes a caller asks for are a host-side view of the same numbers (`_idx_vis` on the containers)."""
    dt = np.dtype(dt)
    if dt in (np.dtype(np.int32), np.dtype(np.int64)):
        return dt
    if dt.kind in "iu" and (dt.itemsize < 4):
        return np.dtype(np.int32)
    return np.dtype(np.int64)


def upload_index(arr):
    """Host index array (coords / indices / indptr) -> device int32 / int64 tensor."""
    a = np.asarray(arr)
    if a.size == 0 and a.dtype.kind not in "iub":
        a = a.astype(np.int64)  # np.asarray([]) is float64
    if a.dtype.kind not in "iub":
        raise ValueError(f"index arrays must be integers, got {a.dtype}")
    return upload(a, device_index_dtype(a.dtype))


_PIN_MIN_BYTES = 1 << 20


def pinned_empty(shape, dtype) -> np.ndarray:
    """NumPy array backed by page-locked host memory (torch's caching host allocator keeps the pages mapped,
    so repeated results of the same size do not pay cudaHostAlloc again).  The array keeps the allocation alive."""
    t = torch()
    if _TEST_CPU or not t.cuda.is_available():
        return np.empty(shape, dtype=dtype)
    return t.empty(shape, dtype=torch_dtype(dtype), pin_memory=True).numpy()


def download(x) -> np.ndarray:
    """Device tensor -> NumPy (one D2H; large results land in pinned memory so the copy runs at link rate)."""
    x = x.detach()
    if _TEST_CPU or not x.is_cuda:
        return x.cpu().numpy()
    if x.numel() * x.element_size() >= _PIN_MIN_BYTES:
        t = torch()
        host = t.empty(x.shape, dtype=x.dtype, pin_memory=True)
        host.copy_(x, non_blocking=False)
        return host.numpy()
    return x.cpu().numpy()


def empty(shape, dtype):
    require_device()
    return torch().empty(shape, dtype=torch_dtype(dtype), device=device())


def zeros(shape, dtype):
    require_device()
    return torch().zeros(shape, dtype=torch_dtype(dtype), device=device())


def ptr(x) -> int:
    return int(x.data_ptr())

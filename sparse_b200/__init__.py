"""sparse_b200 -- a B200-native (sm_100a) implementation of pydata/sparse's data-parallel hot path.

Drop-in names for that path: ``COO``, ``GCXS`` (``CSR``/``CSC``), ``tensordot``, ``matmul``, ``dot``,
``elemwise``, reductions (``sum``/``max``/``min``/``prod``/``mean``/``any``/``all`` and the ``nan*`` forms) and the NumPy protocols
(``__array_ufunc__``, ``__array_function__``, ``@``), plus the fused ``sddmm`` and ``mttkrp`` example paths.
Host code is Python; every data-path step is a hand-written CUDA kernel in ``libsparse_b200.so`` reached through a
thin C ABI (``include/sparse_b200.h``) via ctypes.  There is no CPU fallback: without the library or a CUDA device
operations raise.
"""
from ._argreduce import argmax, argmin
from ._coo import COO, as_coo
from ._creation import (abs, argwhere, asarray, asnumpy, astype, can_cast, diff, empty, empty_like, equal, eye,
                        full, full_like, imag, interp, isinf, isnan, isneginf, isposinf, nonzero, ones, ones_like, real,
                        reshape, result_type, round, std, var, vecdot, zeros, zeros_like)
from ._dot import dot, matmul, tensordot
from ._einsum import einsum
from ._elemwise import broadcast_to, elemwise, where
from ._fused import mttkrp, sddmm
from ._gcxs import CSC, CSR, GCXS
from ._io import load_npz, save_npz
from ._manip import (concatenate, diagonal, diagonalize, expand_dims, flip, kron, matrix_transpose, moveaxis, outer,
                     pad, permute_dims, repeat, roll, squeeze, stack, swapaxes, take, tile, tril, triu, unstack)
from ._nanreduce import nanmax, nanmean, nanmin, nanprod, nanreduce, nansum
from ._random import random
from ._sorting import sort, unique_counts, unique_values
from ._sparse_array import SparseArray

__version__ = "0.1.0"


def clip(a, min=None, max=None, out=None):
    """_coo/common.py:1028-1071."""
    return (a if isinstance(a, SparseArray) else as_coo(a)).clip(min, max, out=out)


concat = concatenate


def sum(x, /, *, axis=None, dtype=None, keepdims=False):
    return x.sum(axis=axis, keepdims=keepdims, dtype=dtype)


def max(x, /, *, axis=None, keepdims=False):
    return x.max(axis=axis, keepdims=keepdims)


def min(x, /, *, axis=None, keepdims=False):
    return x.min(axis=axis, keepdims=keepdims)


def prod(x, /, *, axis=None, dtype=None, keepdims=False):
    return x.prod(axis=axis, keepdims=keepdims, dtype=dtype)


def mean(x, /, *, axis=None, keepdims=False, dtype=None):
    return x.mean(axis=axis, keepdims=keepdims, dtype=dtype)


def any(x, /, *, axis=None, keepdims=False):
    return x.any(axis=axis, keepdims=keepdims)


def all(x, /, *, axis=None, keepdims=False):
    return x.all(axis=axis, keepdims=keepdims)


__all__ = ["COO", "GCXS", "CSR", "CSC", "SparseArray", "as_coo", "asarray", "tensordot", "matmul", "dot", "stack",
           "elemwise", "broadcast_to", "where", "sddmm", "mttkrp", "random", "sum", "max", "min", "prod", "mean", "any", "all",
           "einsum", "save_npz", "load_npz", "nansum", "nanprod", "nanmean", "nanmax", "nanmin", "nanreduce",
           # array manipulation and creation next to the hot path (widened per SURVEY.md s8f)
           "concatenate", "concat", "unstack", "moveaxis", "swapaxes", "permute_dims", "matrix_transpose", "squeeze",
           "expand_dims", "flip", "roll", "triu", "tril", "diagonal", "diagonalize", "pad", "repeat", "tile", "outer",
           "kron", "take", "clip", "eye", "full", "full_like", "zeros", "zeros_like", "ones", "ones_like", "empty",
           "empty_like", "asnumpy", "can_cast", "result_type", "std", "var", "abs", "reshape", "astype", "equal",
           "argmax", "argmin", "interp", "sort", "unique_values", "unique_counts", "round", "isinf", "isnan", "isposinf", "isneginf", "nonzero", "argwhere", "imag", "real", "vecdot", "diff"]

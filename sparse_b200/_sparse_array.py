"""SparseArray base: NumPy protocol dispatch and reductions.

Mirrors sparse/numba_backend/_sparse_array.py: `__array_ufunc__` (:322-370) routes ufunc calls to
`elemwise` and ufunc.reduce to `reduce` (:372-437); `__array_function__` (:282-308) routes
np.tensordot / np.matmul / np.dot / np.sum ... to this package's functions.
"""
from __future__ import annotations

import builtins
import operator
import warnings
from functools import reduce as _functools_reduce

import numpy as np
from numpy.lib.mixins import NDArrayOperatorsMixin

from ._utils import _zero_of_dtype, equivalent, normalize_axis

_reduce_super_ufunc = {np.add: np.multiply, np.multiply: np.power}


class SparseArray(NDArrayOperatorsMixin):
    __array_priority__ = 12

    def __init__(self, shape, fill_value=None):
        if not isinstance(shape, (tuple, list, np.ndarray)) and shape is not None:
            shape = (shape,)
        if not all(isinstance(s, (int, np.integer)) and int(s) >= 0 for s in shape):
            raise ValueError("shape must be an non-negative integer or a tuple of non-negative integers.")
        self.shape = tuple(int(s) for s in shape)
        if fill_value is not None:
            self.fill_value = np.asarray(fill_value)[()] if not isinstance(fill_value, np.generic) else fill_value

    # ---- shape metadata --------------------------------------------------------------------------
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return _functools_reduce(operator.mul, self.shape, 1)

    @property
    def density(self):
        return self.nnz / self.size if self.size else float("nan")  # 0 / 0 like upstream's NumPy division

    @property
    def device(self):
        return "cuda"

    def __len__(self):
        if self.ndim == 0:
            raise TypeError("len() of unsized object")
        return self.shape[0]

    def __array__(self, *args, **kwargs):
        """No silent densification (_sparse_array.py:270-280): opt in with SPARSE_AUTO_DENSIFY=1 or call todense()."""
        from . import _settings

        if not _settings.AUTO_DENSIFY:
            raise RuntimeError("Cannot convert a sparse array to dense automatically. To manually densify, use the "
                               "todense method.")
        return np.asarray(self.todense(), *args, **kwargs)

    def __repr__(self):
        return (f"<{type(self).__name__}: shape={self.shape}, dtype={self.dtype}, nnz={self.nnz}, "
                f"fill_value={self.fill_value}>")

    __str__ = __repr__

    def _repr_html_(self):
        """Jupyter summary table (same rows as upstream's `html_table`, _utils.py:472-504)."""
        def size(n):
            for unit, lim in (("", 2**10), ("K", 2**20), ("M", 2**30), ("G", 2**40), ("T", 2**50)):
                if n < lim:
                    return str(n) if not unit else f"{n / (lim >> 10):.1f}{unit}"
            return f"{n / 2**50:.1f}P"

        density = self.nnz / self.size if self.size else float("nan")
        rows = [("Format", type(self).__name__.lower()), ("Data Type", self.dtype), ("Shape", self.shape),
                ("nnz", self.nnz), ("Density", density), ("Read-only", not hasattr(self, "__setitem__"))]
        if hasattr(self, "nbytes"):
            dense = self.size * np.dtype(self.dtype).itemsize
            rows += [("Size", size(self.nbytes)), ("Storage ratio", f"{self.nbytes / dense if dense else float('nan'):.2f}")]
        if type(self).__name__ == "GCXS":
            rows.append(("Compressed Axes", self.compressed_axes))
        cells = "".join(f'<tr><th style="text-align: left">{h}</th><td style="text-align: left">{v}</td></tr>'
                        for h, v in rows)
        return f"<table><tbody>{cells}</tbody></table>"

    # ---- NumPy protocols ---------------------------------------------------------------------------
    def __array_function__(self, func, types, args, kwargs):
        import sparse_b200 as module

        name = func.__name__
        sparse_func = getattr(module, name, None)
        if sparse_func is None or sparse_func is func:
            sparse_func = getattr(type(self), name, None)
        if (sparse_func is None or not callable(sparse_func)) and len(args) == 1 and not kwargs and args[0] is self \
                and name in ("shape", "ndim", "size"):
            return getattr(self, name)  # np.shape / np.ndim / np.size: plain attributes (_sparse_array.py:300-304)
        if sparse_func is None or not callable(sparse_func):
            return NotImplemented
        return sparse_func(*args, **kwargs)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        """_sparse_array.py:322-370: `__call__` -> elemwise, `reduce` -> reduce, `outer` -> broadcast call.

        `out=` (and therefore the in-place operators `+=`, `*=` ... of NDArrayOperatorsMixin) is honoured the way
        upstream does it: NumPy's own casting check is run on one-element stand-ins (so an illegal cast raises the
        same `UFuncTypeError`), the operation is computed in `out`'s dtype, and `out` then adopts the result's device
        arrays (a shallow re-point, no copy)."""
        from ._elemwise import elemwise

        out = kwargs.pop("out", None)
        if out is not None and not isinstance(out, tuple):
            out = (out,)
        if out is not None and not all(isinstance(x, type(self)) for x in out):
            return NotImplemented
        if getattr(ufunc, "signature", None) is not None:
            return self.__array_function__(ufunc, (np.ndarray, type(self)), inputs, kwargs)
        if out is not None:
            test_args = [np.empty((1,), dtype=a.dtype) if hasattr(a, "dtype") else a for a in inputs]
            test_kwargs = kwargs.copy()
            if method == "reduce":
                test_kwargs["axis"] = None
            test_out = tuple(np.empty(() if method == "reduce" else (1,), dtype=a.dtype) for a in out)
            with np.errstate(all="ignore"):
                getattr(ufunc, method)(*test_args, out=test_out[0] if len(test_out) == 1 else test_out, **test_kwargs)
            kwargs["dtype"] = out[0].dtype
        if method == "outer":
            method = "__call__"
            cum_ndim = 0
            inputs_transformed = []
            for inp in reversed(inputs):  # the LAST operand keeps its axes, earlier ones get trailing unit axes
                inputs_transformed.append(inp[(Ellipsis,) + (None,) * cum_ndim])
                cum_ndim += inp.ndim
            inputs = tuple(reversed(inputs_transformed))
        if method == "__call__":
            result = elemwise(ufunc, *inputs, **kwargs)
        elif method == "reduce":
            result = SparseArray._reduce(ufunc, *inputs, **kwargs)
        else:
            return NotImplemented
        if out is not None and result is not NotImplemented:
            (out,) = out
            if tuple(out.shape) != tuple(result.shape):
                raise ValueError(f"non-broadcastable output operand with shape {out.shape} "
                                 f"doesn't match the broadcast shape {result.shape}")
            out._make_shallow_copy_of(result)
            return out
        return result

    # ---- reductions (reference: _sparse_array.py:310-437) ------------------------------------------
    @staticmethod
    def _reduce(method, *args, **kwargs):
        assert len(args) == 1
        self = args[0]
        if isinstance(self, np.ndarray):
            return method.reduce(self, **kwargs)
        return self.reduce(method, **kwargs)

    def reduce(self, method, axis=(0,), keepdims=False, **kwargs):
        from ._reduce import reduce_impl

        return reduce_impl(self, method, axis=axis, keepdims=keepdims, **kwargs)

    def sum(self, axis=None, keepdims=False, dtype=None, out=None):
        return np.add.reduce(self, out=out, axis=axis, keepdims=keepdims, dtype=dtype)

    def max(self, axis=None, keepdims=False, out=None):
        return np.maximum.reduce(self, out=out, axis=axis, keepdims=keepdims)

    amax = max

    def min(self, axis=None, keepdims=False, out=None):
        return np.minimum.reduce(self, out=out, axis=axis, keepdims=keepdims)

    amin = min

    def prod(self, axis=None, keepdims=False, dtype=None, out=None):
        return np.multiply.reduce(self, out=out, axis=axis, keepdims=keepdims, dtype=dtype)

    def any(self, axis=None, keepdims=False, out=None):
        return np.logical_or.reduce(self, out=out, axis=axis, keepdims=keepdims)

    def all(self, axis=None, keepdims=False, out=None):
        return np.logical_and.reduce(self, out=out, axis=axis, keepdims=keepdims)

    def mean(self, axis=None, keepdims=False, dtype=None, out=None):
        """_sparse_array.py:mean -- sum / n with NumPy's dtype rules."""
        if axis is None:
            axis = tuple(range(self.ndim))
        elif not isinstance(axis, tuple):
            axis = (axis,)
        den = _functools_reduce(operator.mul, (self.shape[i] for i in axis), 1)
        if dtype is None:
            if issubclass(self.dtype.type, (np.integer, np.bool_)):
                dtype = inter_dtype = np.dtype("f8")
            else:
                dtype = self.dtype
                inter_dtype = np.dtype("f4") if issubclass(dtype.type, np.float16) else dtype
        else:
            inter_dtype = dtype
        num = self.sum(axis=axis, keepdims=keepdims, dtype=inter_dtype)
        if num.ndim:
            out = np.true_divide(num, den)
        else:
            out = (num / den) if not isinstance(num, SparseArray) else np.true_divide(num, den)
        return out.astype(dtype) if hasattr(out, "astype") else out

    def var(self, axis=None, dtype=None, out=None, ddof=0, keepdims=False):
        """Variance along `axis` (_sparse_array.py:725-813): two-pass form, mean first, then the mean of the squared
        deviations; every step is an element-wise or reduction kernel on the device."""
        axis = normalize_axis(axis, self.ndim)
        if axis is None:
            axis = tuple(range(self.ndim))
        if not isinstance(axis, tuple):
            axis = (axis,)
        rcount = _functools_reduce(operator.mul, (self.shape[a] for a in axis), 1)
        if ddof >= rcount:
            warnings.warn("Degrees of freedom <= 0 for slice", RuntimeWarning, stacklevel=1)
        if dtype is None and issubclass(self.dtype.type, (np.integer, np.bool_)):
            dtype = np.dtype("f8")
        with np.errstate(all="ignore"):
            arrmean = np.true_divide(self.sum(axis, dtype=dtype, keepdims=True), rcount)
            x = self - arrmean
            x = np.multiply(x, x)
            ret = x.sum(axis=axis, dtype=dtype, keepdims=keepdims)
            rcount = max(rcount - ddof, 0)
            res = np.true_divide(ret, rcount)
            if hasattr(res, "astype") and res.dtype != ret.dtype:
                res = res.astype(ret.dtype)
        if out is not None:
            out._make_shallow_copy_of(res)
            return out
        return res

    def std(self, axis=None, dtype=None, out=None, ddof=0, keepdims=False):
        """Standard deviation = sqrt(var) (_sparse_array.py:815-877)."""
        ret = self.var(axis=axis, dtype=dtype, ddof=ddof, keepdims=keepdims)
        with np.errstate(all="ignore"):
            res = np.sqrt(ret)
        if out is not None:
            out._make_shallow_copy_of(res)
            return out
        return res

    def round(self, decimals=0, out=None):
        """numpy.round semantics (_sparse_array.py:592-606): round-half-even at `decimals` places, as NumPy computes
        it -- scale by 10**|decimals|, rint, scale back (integers with decimals >= 0 are returned unchanged)."""
        if out is not None and not isinstance(out, tuple):
            out = (out,)
        decimals = operator.index(decimals)
        kw = {"out": out} if out is not None else {}
        if issubclass(self.dtype.type, (np.integer, np.bool_)) and decimals >= 0:
            return np.positive(self, **kw) if self.dtype != np.bool_ else self.copy()
        if decimals == 0:
            return np.rint(self, **kw)
        with np.errstate(all="ignore"):
            if issubclass(self.dtype.type, np.integer):
                p = self.dtype.type(10 ** -decimals)
                q = np.floor_divide(self, p)
                r = np.subtract(self, np.multiply(q, p))
                # integers, negative decimals: multiples of 10**-decimals, ties to even multiple like NumPy
                half = p // 2
                up = np.logical_or(np.greater(r, half), np.logical_and(np.equal(r, half), np.not_equal(
                    np.remainder(q, 2), 0))) if p % 2 == 0 else np.greater(r, half)
                res = np.multiply(np.add(q, up.astype(self.dtype)), p)
                if out is not None:
                    out[0]._make_shallow_copy_of(res)
                    return out[0]
                return res
            p = self.dtype.type(10.0 ** abs(decimals))
            if decimals > 0:
                return np.true_divide(np.rint(np.multiply(self, p)), p, **kw)
            return np.multiply(np.rint(np.true_divide(self, p)), p, **kw)

    round_ = round

    def clip(self, min=None, max=None, out=None):
        """numpy.clip (_sparse_array.py:610-624): maximum(x, min) then minimum(., max)."""
        if min is None and max is None:
            raise ValueError("One of max or min must be given.")
        if out is not None and not isinstance(out, tuple):
            out = (out,)
        kw = {"out": out} if out is not None else {}
        if self.dtype.kind in "iu":
            # numpy.clip accepts Python integers outside the dtype's range (a bound below the type's minimum clips
            # nothing); minimum / maximum do not (NEP 50: OverflowError), so such bounds are brought into range first
            info = np.iinfo(self.dtype)
            min, max = (builtins.min(builtins.max(int(b), info.min), info.max)
                        if type(b) is int else b for b in (min, max))
        if max is None:
            return np.maximum(self, min, **kw)
        if min is None:
            return np.minimum(self, max, **kw)
        return np.minimum(np.maximum(self, min), max, **kw)

    def conj(self):
        return np.conj(self)

    def maybe_densify(self, max_size=1000, min_density=0.25):
        """Dense array if the array is small or dense enough, else ValueError (_coo/core.py:1393-1448)."""
        if self.size > max_size and self.density < min_density:
            raise ValueError("Operation would require converting large sparse array to dense")
        return self.todense()

    def todok(self):
        from ._dok import DOK

        return DOK.from_coo(self.asformat("coo"))

    # ---- structure (thin forwards to _manip; `COO.flatten/swapaxes/squeeze/nonzero`, _coo/core.py) ----------
    def flatten(self, order="C"):
        if order not in {"C", None}:
            raise NotImplementedError("The `order` parameter is not supported")
        return self.reshape(-1)

    def swapaxes(self, axis1, axis2):
        from ._manip import swapaxes

        return swapaxes(self, axis1, axis2)

    def squeeze(self, axis=None):
        from ._manip import squeeze

        return squeeze(self, axis)

    def nonzero(self):
        from ._creation import nonzero

        return nonzero(self)

    def isinf(self):
        return np.isinf(self)

    def isnan(self):
        return np.isnan(self)

    # ---- scalar conversion (_sparse_array.py:970-993) ---------------------------------------------------
    def _to_scalar(self, builtin):
        if self.size != 1 or self.shape != ():
            raise ValueError(f"{builtin} can be computed for one-element arrays only.")
        return builtin(self.todense().flatten()[0])

    def __bool__(self):
        return self._to_scalar(bool)

    def __float__(self):
        return self._to_scalar(float)

    def __int__(self):
        return self._to_scalar(int)

    def __index__(self):
        return self._to_scalar(int)

    # ---- misc -----------------------------------------------------------------------------------------
    @property
    def real(self):
        """numpy.real: the array itself for real dtypes, the pruned real plane for complex ones."""
        if self.dtype.kind != "c":
            # upstream: elemwise(np.real, x) -- the same array, in elemwise's clothes: a 0-D array comes back as its
            # value-as-fill-value form, an array with a zero-length axis as the empty COO (_umath.py:438-439, 467-477)
            if self.ndim == 0:
                return self.astype(self.dtype)
            if any(s == 0 for s in self.shape):
                return self.asformat("coo")
            return self
        from ._complex import planes

        return np.positive(planes(self)[0])  # the unary pass drops the zeros stored in the plane

    @property
    def imag(self):
        """numpy.imag: all zeros (same dtype) for real dtypes, the pruned imaginary plane for complex ones."""
        if self.dtype.kind != "c":
            # every element is +0 (x * 0 would leave -0.0 behind negative entries): no stored entry, same format
            from ._coo import COO

            out = COO(np.empty((self.ndim, 0), dtype=np.intp), np.empty(0, dtype=self.dtype), shape=self.shape,
                      has_duplicates=False, sorted=True, fill_value=self.dtype.type(0))
            if any(s == 0 for s in self.shape) or isinstance(self, COO):
                return out
            ca = getattr(self, "compressed_axes", None)
            return out.asformat(self.format, **({"compressed_axes": ca} if self.format == "gcxs" else {}))
        from ._complex import planes

        return np.positive(planes(self)[1])

    def __complex__(self):
        return self._to_scalar(complex)

    def __array_namespace__(self, *, api_version=None):
        if api_version is None:
            api_version = "2024.12"
        if api_version not in {"2021.12", "2022.12", "2023.12", "2024.12"}:
            raise ValueError(f'"{api_version}" Array API version not supported.')
        import sparse_b200

        return sparse_b200

    def _zero_fill(self):
        return bool(equivalent(self.fill_value, _zero_of_dtype(self.dtype), loose=True))

    def _norm_axis(self, axis):
        return normalize_axis(axis, self.ndim)

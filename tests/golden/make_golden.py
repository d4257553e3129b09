#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the REFERENCE itself.

Run in the authoring container only (needs /root/reference and numba):

    python tests/golden/make_golden.py

The reference tree is read-only and lacks the setuptools-scm ``_version.py``,
so a scratch copy is made under a temp dir (never inside this repo) and
imported from there.  Outputs are small ``.npz`` files holding the seeded
inputs and the reference's outputs; nothing from the reference's sources is
copied into the repo.  The GPU box never runs this script: tests read only
the committed fixtures.

Fixture layout: every ``.npz`` has a ``meta`` entry (JSON) listing cases; the
arrays of case ``i`` are stored as ``c{i}__<name>``.  Sparse values are
encoded by ``enc()`` below (kind = coo | gcxs | dense).
"""
from __future__ import annotations

import json
import os
import shutil
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SPARSE_REFERENCE", "/root/reference")


def import_reference():
    tmp = os.path.join(tempfile.gettempdir(), "sparse_b200_refcopy")
    if os.path.isdir(tmp):
        shutil.rmtree(tmp)
    shutil.copytree(os.path.join(REF, "sparse"), os.path.join(tmp, "sparse"))
    with open(os.path.join(tmp, "sparse", "_version.py"), "w") as f:
        f.write('__version__ = "0.0.0+ref"\n__version_tuple__ = (0, 0, 0)\n')
    sys.path.insert(0, tmp)
    import sparse  # noqa: E402

    assert sparse.__file__.startswith(tmp)
    return sparse


sparse = import_reference()
from sparse.numba_backend import _common as C  # noqa: E402
from sparse.numba_backend import _umath as U  # noqa: E402
from sparse.numba_backend._coo import core as coo_core  # noqa: E402
from sparse.numba_backend._compressed import convert as gconv  # noqa: E402


class Book:
    def __init__(self, name):
        self.name = name
        self.arrays = {}
        self.cases = []

    def add(self, info, **arrays):
        i = len(self.cases)
        self.cases.append(info)
        for k, v in arrays.items():
            self.arrays[f"c{i}__{k}"] = np.asarray(v)

    def save(self):
        self.arrays["meta"] = np.array(json.dumps(self.cases))
        path = os.path.join(HERE, self.name + ".npz")
        np.savez_compressed(path, **self.arrays)
        print(f"{path}: {len(self.cases)} cases, {os.path.getsize(path)} bytes")


def enc(prefix, x):
    """Encode a result (COO / GCXS / ndarray / scalar) as a dict of arrays."""
    out = {}
    if isinstance(x, sparse.COO):
        out[prefix + "kind"] = "coo"
        out[prefix + "coords"] = x.coords
        out[prefix + "data"] = x.data
        out[prefix + "shape"] = np.array(x.shape, dtype=np.int64)
        out[prefix + "fill"] = np.asarray(x.fill_value)
    elif isinstance(x, sparse.GCXS):
        out[prefix + "kind"] = "gcxs"
        out[prefix + "data"] = x.data
        out[prefix + "indices"] = x.indices
        out[prefix + "indptr"] = np.asarray(x.indptr)
        out[prefix + "shape"] = np.array(x.shape, dtype=np.int64)
        out[prefix + "ca"] = np.array(x.compressed_axes if x.compressed_axes is not None else (), dtype=np.int64)
        out[prefix + "fill"] = np.asarray(x.fill_value)
    else:
        out[prefix + "kind"] = "dense"
        out[prefix + "array"] = np.asarray(x)
    return out


def rand_sparse(rng, shape, density, dtype, fmt="coo", **kw):
    x = sparse.random(shape, density=density, random_state=rng, format=fmt, **kw)
    if np.issubdtype(np.dtype(dtype), np.integer):
        x = (x * 10).astype(dtype)  # small ints, zeros get pruned below
        x = sparse.COO(x.coords, x.data, shape=x.shape, prune=True) if fmt == "coo" else x
    else:
        x = x.astype(dtype)
    return x


def csr_of(x):
    g = x.asformat("gcxs", compressed_axes=(0,))
    return g.data, g.indices, g.indptr


def csc_of(x):
    g = x.asformat("gcxs", compressed_axes=(1,))
    return g.data, g.indices, g.indptr


# --------------------------------------------------------------------------
def gen_dot_kernels():
    bk = Book("dot_kernels")
    rng = np.random.default_rng(20260922)
    shapes = [((7, 9), (9, 5), 0.4), ((30, 40), (40, 12), 0.15), ((1, 6), (6, 1), 0.9), ((12, 8), (8, 130), 0.3),
              ((25, 25), (25, 25), 0.02)]
    for dt in ("float32", "float64", "int64", "int32"):
        for (sa, sb, dens) in shapes:
            a = rand_sparse(rng, sa, dens, dt)
            bs = rand_sparse(rng, sb, dens, dt)
            bd = bs.todense() if dens < 0.5 else (rng.random(sb) * 4 - 2).astype(dt)
            if np.issubdtype(np.dtype(dt), np.floating):
                bd = bd.copy()
                bd[rng.random(sb) < 0.3] = 0  # exercise the structural test on B
            out_shape = (sa[0], sb[1])
            ad, ai, ap = csr_of(a)
            cd, ci, cp = csc_of(a)
            bdt, bi, bp = csr_of(bs)
            info = {"dtype": dt, "a_shape": sa, "b_shape": sb}
            # csr @ dense
            bk.add({**info, "kernel": "csr_ndarray"}, a_data=ad, a_indices=ai, a_indptr=ap, b=bd,
                   out=C._dot_csr_ndarray_type(ad.dtype, bd.dtype)(out_shape, ad, ai, ap, bd))
            d, i, p = C._dot_csr_ndarray_type_sparse(ad.dtype, bd.dtype)(out_shape, ad, ai, ap, bd)
            bk.add({**info, "kernel": "csr_ndarray_sparse"}, a_data=ad, a_indices=ai, a_indptr=ap, b=bd,
                   data=d, indices=i, indptr=p)
            # csc @ dense
            bk.add({**info, "kernel": "csc_ndarray"}, a_data=cd, a_indices=ci, a_indptr=cp, b=bd,
                   out=C._dot_csc_ndarray_type(cd.dtype, bd.dtype)(sa, sb, cd, ci, cp, bd))
            d, i, p = C._dot_csc_ndarray_type_sparse(cd.dtype, bd.dtype)(sa, sb, cd, ci, cp, bd)
            # only the prefix indptr[-1]... the reference leaves a garbage tail when sums cancel; store the
            # GCXS-visible part: data/indices are np.empty-sized by the structural count.
            bk.add({**info, "kernel": "csc_ndarray_sparse"}, a_data=cd, a_indices=ci, a_indptr=cp, b=bd,
                   data=d, indices=i, indptr=p)
            # csr @ csr and coo @ coo
            d, i, p = C._dot_csr_csr_type(ad.dtype, bdt.dtype)(out_shape, ad, bdt, ai, bi, ap, bp)
            bk.add({**info, "kernel": "csr_csr"}, a_data=ad, a_indices=ai, a_indptr=ap, b_data=bdt, b_indices=bi,
                   b_indptr=bp, data=d, indices=i, indptr=p)
            a_ip = np.concatenate([[0], np.cumsum(np.bincount(a.coords[0], minlength=sa[0]))]).astype(np.intp)
            b_ip = np.concatenate([[0], np.cumsum(np.bincount(bs.coords[0], minlength=sb[0]))]).astype(np.intp)
            co, d = C._dot_coo_coo_type(a.dtype, bs.dtype)(out_shape, a.coords, bs.coords, a.data, bs.data, a_ip, b_ip)
            bk.add({**info, "kernel": "coo_coo"}, a_coords=a.coords, a_data=a.data, b_coords=bs.coords,
                   b_data=bs.data, a_indptr=a_ip, b_indptr=b_ip, coords=co, data=d)
            # coo @ dense (takes b.T view)
            bt = bd.T
            bk.add({**info, "kernel": "coo_ndarray"}, a_coords=a.coords, a_data=a.data, b=bd,
                   out=C._dot_coo_ndarray_type(a.dtype, bt.dtype)(a.coords, a.data, bt, out_shape))
            co, d = C._dot_coo_ndarray_type_sparse(a.dtype, bt.dtype)(a.coords, a.data, bt, out_shape)
            bk.add({**info, "kernel": "coo_ndarray_sparse"}, a_coords=a.coords, a_data=a.data, b=bd,
                   coords=np.asarray(co).reshape(2, -1), data=d)
            # dense @ coo
            adense = a.todense()
            bk.add({**info, "kernel": "ndarray_coo"}, a=adense, b_coords=bs.coords, b_data=bs.data,
                   out=C._dot_ndarray_coo_type(adense.dtype, bs.dtype)(adense, bs.coords, bs.data, out_shape))
            bT = bs.T
            co, d = C._dot_ndarray_coo_type_sparse(adense.dtype, bT.dtype)(adense, bT.coords, bT.data, out_shape)
            bk.add({**info, "kernel": "ndarray_coo_sparse"}, a=adense, bt_coords=bT.coords, bt_data=bT.data,
                   coords=np.asarray(co).reshape(2, -1), data=d)
    # fully dense product: exercises the row reversal at _common.py:709-714
    for dt in ("float64", "float32"):
        a = sparse.COO.from_numpy((rng.random((4, 5)) + 0.5).astype(dt))
        b = sparse.COO.from_numpy((rng.random((5, 3)) + 0.5).astype(dt))
        ad, ai, ap = csr_of(a)
        bdt, bi, bp = csr_of(b)
        d, i, p = C._dot_csr_csr_type(ad.dtype, bdt.dtype)((4, 3), ad, bdt, ai, bi, ap, bp)
        bk.add({"dtype": dt, "a_shape": (4, 5), "b_shape": (5, 3), "kernel": "csr_csr", "note": "all-dense flip"},
               a_data=ad, a_indices=ai, a_indptr=ap, b_data=bdt, b_indices=bi, b_indptr=bp, data=d, indices=i, indptr=p)
    # cancellation to +0.0 (pruned later) and -0.0 handling
    a = sparse.COO(np.array([[0, 0, 1], [0, 1, 1]]), np.array([1.0, -1.0, 2.0]), shape=(2, 2))
    b = sparse.COO(np.array([[0, 1, 1], [0, 0, 1]]), np.array([3.0, 3.0, -0.0]), shape=(2, 2))
    ad, ai, ap = csr_of(a)
    bdt, bi, bp = csr_of(b)
    d, i, p = C._dot_csr_csr_type(ad.dtype, bdt.dtype)((2, 2), ad, bdt, ai, bi, ap, bp)
    bk.add({"dtype": "float64", "a_shape": (2, 2), "b_shape": (2, 2), "kernel": "csr_csr", "note": "cancellation"},
           a_data=ad, a_indices=ai, a_indptr=ap, b_data=bdt, b_indices=bi, b_indptr=bp, data=d, indices=i, indptr=p)
    # _match_arrays with duplicate runs on both sides; _calc_counts_invidx
    for n1, n2, hi in ((50, 40, 12), (200, 7, 30), (5, 300, 9), (0, 4, 3), (64, 64, 1000)):
        x = np.sort(rng.integers(0, hi, n1)).astype(np.intp)
        y = np.sort(rng.integers(0, hi, n2)).astype(np.intp)
        ia, ib = U._match_arrays(x, y)
        bk.add({"kernel": "match_arrays"}, a=x, b=y, ia=ia, ib=ib)
        if n1:
            inv, cnt = coo_core._calc_counts_invidx(x)
            bk.add({"kernel": "counts_invidx"}, groups=x, inv_idx=inv, counts=cnt)
    ip = np.array([0, 0, 3, 3, 4, 9], dtype=np.intp)
    bk.add({"kernel": "uncompress"}, indptr=ip, rows=gconv.uncompress_dimension(ip))
    bk.save()


# --------------------------------------------------------------------------
def gen_tensordot():
    """Public-API level: sparse.tensordot / matmul / dot over formats and return types
    (shapes follow tests/test_dot.py:15-80, 114-163)."""
    bk = Book("tensordot_api")
    rng = np.random.default_rng(7)
    cases = [
        ((3, 4, 5), (4, 3), (1, 0)),
        ((3, 4, 5), (4, 5, 6), ((1, 2), (0, 1))),
        ((4, 5), (5, 4), 1),
        ((2, 3, 4), (4, 3, 2), ((0, 1), (2, 1))),
        ((5,), (5, 6), ((0,), (0,))),
        ((3, 4), (4,), 1),
        ((6, 7), (6, 7), 2),
    ]
    fmts = [("coo", "coo"), ("coo", "gcxs"), ("gcxs", "coo"), ("gcxs", "gcxs"), ("coo", "dense"), ("dense", "coo"),
            ("gcxs", "dense"), ("dense", "gcxs")]
    rts = {"none": None, "coo": sparse.COO, "gcxs": sparse.GCXS, "dense": np.ndarray}
    for dt in ("float64", "float32"):
        for sa, sb, axes in cases:
            a = rand_sparse(rng, sa, 0.5, dt)
            b = rand_sparse(rng, sb, 0.5, dt)
            for fa, fb in fmts:
                for rtn, rt in rts.items():
                    xa = a.todense() if fa == "dense" else a.asformat(fa)
                    xb = b.todense() if fb == "dense" else b.asformat(fb)
                    try:
                        r = sparse.tensordot(xa, xb, axes, return_type=rt)
                    except Exception as e:  # record the failure class too
                        bk.add({"op": "tensordot", "dtype": dt, "fa": fa, "fb": fb, "axes": axes, "rt": rtn,
                                "error": type(e).__name__},
                               **enc("a_", a), **enc("b_", b))
                        continue
                    bk.add({"op": "tensordot", "dtype": dt, "fa": fa, "fb": fb, "axes": axes, "rt": rtn},
                           **enc("a_", a), **enc("b_", b), **enc("out_", r))
    # matmul broadcasting (tests/test_dot.py:114-163)
    mm = [((1, 4, 5), (3, 5, 6)), ((3, 4, 5), (1, 5, 6)), ((3, 4, 5), (3, 5, 6)), ((3, 4, 5), (5, 6)),
          ((4, 5), (5, 6)), ((5,), (5, 6)), ((4, 5), (5,)), ((5,), (5,)), ((3, 4), (1, 2, 4, 3))]
    for sa, sb in mm:
        a = rand_sparse(rng, sa, 0.5, "float64")
        b = rand_sparse(rng, sb, 0.5, "float64")
        for fa, fb in (("coo", "coo"), ("gcxs", "gcxs"), ("coo", "dense"), ("dense", "gcxs")):
            xa = a.todense() if fa == "dense" else a.asformat(fa)
            xb = b.todense() if fb == "dense" else b.asformat(fb)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                r = sparse.matmul(xa, xb)
            bk.add({"op": "matmul", "fa": fa, "fb": fb}, **enc("a_", a), **enc("b_", b), **enc("out_", r))
    # test_small_values (tests/test_dot.py:289-300)
    a = sparse.COO(np.array([[0, 10]]), np.array([3.6e-100, 7.2e-009]), shape=(20,))
    b = sparse.COO(np.array([[0, 0], [4, 28]]), np.array([3.8e-25, 4.5e-225]), shape=(20, 50))
    for fa in ("coo", "gcxs"):
        r = sparse.dot(a.asformat(fa), b.asformat(fa))
        bk.add({"op": "dot", "fa": fa, "fb": fa, "note": "small_values"}, **enc("a_", a), **enc("b_", b),
               **enc("out_", r))
    bk.save()


# --------------------------------------------------------------------------
UFUNCS = ["add", "subtract", "multiply", "maximum", "minimum", "greater", "less_equal", "not_equal", "equal",
          "true_divide_rhs_dense"]


def gen_elemwise():
    """Binary broadcasting elemwise (tests/test_elemwise.py:143-156, 206-238, 79-111, 416-511)."""
    bk = Book("elemwise_api")
    rng = np.random.default_rng(11)
    pairs = [((4, 5), (4, 5)), ((2, 3, 4), (3, 4)), ((3, 4), (2, 3, 4)), ((3, 1, 4), (3, 2, 4)), ((3, 4), (1, 4)),
             ((4, 1), (1, 5)), ((5,), (6, 5)), ((2, 3, 4, 5), (2, 3, 4, 1)), ((1, 3, 1, 2), (4, 3, 5, 2)),
             ((6, 0), (6, 0)), ((3, 4), ()), ((2, 1, 3), (1, 4, 1))]
    for dt in ("float64", "float32", "int64"):
        for s1, s2 in pairs:
            a = rand_sparse(rng, s1, 0.4, dt)
            b = rand_sparse(rng, s2, 0.4, dt) if s2 != () else None
            for name in ("add", "subtract", "multiply", "maximum", "minimum", "greater", "less_equal", "not_equal"):
                f = getattr(np, name)
                if b is None:
                    scalar = np.dtype(dt).type(2)
                    try:
                        r = f(a, scalar)
                    except ValueError as e:
                        bk.add({"op": name, "dtype": dt, "rhs": "scalar", "error": "ValueError"}, **enc("a_", a),
                               scalar=scalar)
                        continue
                    bk.add({"op": name, "dtype": dt, "rhs": "scalar"}, **enc("a_", a), scalar=scalar, **enc("out_", r))
                    continue
                try:
                    r = f(a, b)
                except ValueError:
                    bk.add({"op": name, "dtype": dt, "rhs": "coo", "error": "ValueError"}, **enc("a_", a),
                           **enc("b_", b))
                    continue
                bk.add({"op": name, "dtype": dt, "rhs": "coo"}, **enc("a_", a), **enc("b_", b), **enc("out_", r))
    # sparse (op) dense ndarray, broadcasting both ways (tests/test_elemwise.py:79-111)
    for dt in ("float64", "float32"):
        for s1, s2 in [((3, 4), (3, 4)), ((2, 3, 4), (4,)), ((3, 4), (2, 3, 4)), ((3, 1, 4), (2, 4)), ((5, 1), (1, 6))]:
            a = rand_sparse(rng, s1, 0.4, dt)
            d = (rng.random(s2) + 0.5).astype(dt)
            d[rng.random(s2) < 0.2] = 0
            for name in ("multiply", "true_divide"):
                f = getattr(np, name)
                dd = d if name == "multiply" else (d + 1).astype(dt)
                r = f(a, dd)
                bk.add({"op": name, "dtype": dt, "rhs": "dense"}, **enc("a_", a), dense=dd, **enc("out_", r))
                if name == "multiply":
                    r = f(dd, a)
                    bk.add({"op": name, "dtype": dt, "rhs": "dense", "swap": True}, **enc("a_", a), dense=dd,
                           **enc("out_", r))
    # non-zero fill values (tests/test_elemwise.py:387-413) and pathological data (:252-305)
    a = sparse.random((3, 4), density=0.5, random_state=rng, fill_value=2.0)
    b = sparse.random((3, 4), density=0.5, random_state=rng, fill_value=-1.0)
    for name in ("add", "multiply", "maximum"):
        r = getattr(np, name)(a, b)
        bk.add({"op": name, "dtype": "float64", "rhs": "coo", "note": "fill"}, **enc("a_", a), **enc("b_", b),
               **enc("out_", r))
    a = rand_sparse(rng, (4, 5), 0.6, "float64")
    b = rand_sparse(rng, (4, 5), 0.6, "float64")
    a.data[::3] = np.nan
    b.data[1::4] = np.inf
    a.data[1::5] = -0.0 * 1.0
    for name in ("add", "multiply", "maximum", "minimum"):
        with np.errstate(all="ignore"):
            r = getattr(np, name)(a, b)
        bk.add({"op": name, "dtype": "float64", "rhs": "coo", "note": "nan-inf"}, **enc("a_", a), **enc("b_", b),
               **enc("out_", r))
    # unary (tests/test_elemwise.py:13-44)
    a = rand_sparse(rng, (5, 6), 0.5, "float64")
    a = a - 0.3 * (a != 0)  # some negatives
    for name in ("negative", "abs", "sqrt_abs", "sign", "expm1", "sin", "floor", "square"):
        if name == "sqrt_abs":
            r = np.sqrt(np.abs(a))
        else:
            r = getattr(np, name)(a)
        bk.add({"op": name, "dtype": "float64", "rhs": "unary"}, **enc("a_", a), **enc("out_", r))
    # GCXS operands keep GCXS output (_umath.py:416-427)
    a = rand_sparse(rng, (4, 5, 3), 0.4, "float64").asformat("gcxs")
    b = rand_sparse(rng, (4, 5, 3), 0.4, "float64").asformat("gcxs")
    for name in ("add", "multiply"):
        r = getattr(np, name)(a, b)
        bk.add({"op": name, "dtype": "float64", "rhs": "gcxs"}, **enc("a_", a.tocoo()), **enc("b_", b.tocoo()),
               **enc("out_", r), a_ca=np.array(a.compressed_axes), b_ca=np.array(b.compressed_axes))
    bk.save()


# --------------------------------------------------------------------------
def gen_reduce():
    """Reductions (tests/test_coo.py:44-193, tests/test_compressed.py:36-131)."""
    bk = Book("reduce_api")
    rng = np.random.default_rng(13)
    axes_list = [None, 0, 1, 2, (0, 2), (1, 2), -1, (0, 1, 2)]
    for dt in ("float64", "float32", "int64"):
        x = rand_sparse(rng, (5, 6, 7), 0.3, dt)
        for name in ("sum", "max", "min", "prod", "mean", "any", "all"):
            for axis in axes_list:
                for keepdims in (False, True):
                    if name == "mean" and dt == "int64":
                        continue
                    with np.errstate(all="ignore"):
                        r = getattr(x, name)(axis=axis, keepdims=keepdims)
                    ax = axis if axis is None or isinstance(axis, int) else list(axis)
                    bk.add({"op": name, "dtype": dt, "axis": ax, "keepdims": keepdims, "fmt": "coo"},
                           **enc("a_", x), **enc("out_", r))
        g = x.asformat("gcxs")
        for name in ("sum", "max", "min"):
            for axis in (0, (0, 2), None, 2):
                r = getattr(g, name)(axis=axis)
                ax = axis if axis is None or isinstance(axis, int) else list(axis)
                bk.add({"op": name, "dtype": dt, "axis": ax, "keepdims": False, "fmt": "gcxs"}, **enc("a_", x),
                       **enc("out_", r), a_ca=np.array(g.compressed_axes))
    # non-zero fill value (tests/test_coo.py:44-52)
    x = sparse.random((4, 5, 6), density=0.4, random_state=rng, fill_value=3.0)
    for name in ("sum", "prod", "max", "min"):
        for axis in (0, (1, 2), None):
            r = getattr(x, name)(axis=axis)
            ax = axis if axis is None or isinstance(axis, int) else list(axis)
            bk.add({"op": name, "dtype": "float64", "axis": ax, "keepdims": False, "fmt": "coo", "note": "fill"},
                   **enc("a_", x), **enc("out_", r))
    bk.save()

# --------------------------------------------------------------------------
def gen_nanreduce():
    """NaN-skipping reductions (_coo/common.py:346-531, 674-732; tests/test_coo.py:196-263 upstream)."""
    bk = Book("nanreduce_api")
    rng = np.random.default_rng(29)
    axes_list = [None, 0, 2, (0, 2), (1, 2), -1]
    arrays = []
    for dt in ("float64", "float32"):
        x = rand_sparse(rng, (5, 6, 7), 0.35, dt)
        d = x.data.copy()
        d[rng.random(d.shape[0]) < 0.25] = np.nan
        arrays.append((dt, "", sparse.COO(x.coords, d, shape=x.shape)))
    # NaN fill value: every implicit entry is skipped as well
    x = rand_sparse(rng, (4, 5, 6), 0.4, "float64")
    d = x.data.copy()
    d[::5] = np.nan
    arrays.append(("float64", "nanfill", sparse.COO(x.coords, d, shape=x.shape, fill_value=np.nan)))
    # one fully stored all-NaN slice (warns "All-NaN slice" / "Mean of empty slice")
    dn = rng.random((3, 4))
    dn[dn < 0.5] = 0.0
    dn[1, :] = np.nan
    arrays.append(("float64", "allnan", sparse.COO.from_numpy(dn)))
    # integer input: nan-functions degenerate to the plain reductions
    arrays.append(("int64", "", rand_sparse(rng, (5, 6, 7), 0.3, "int64")))
    for dt, note, x in arrays:
        for name in ("nansum", "nanprod", "nanmax", "nanmin", "nanmean"):
            for axis in axes_list:
                if x.ndim == 2 and axis in (2, (0, 2), (1, 2)):
                    continue
                for keepdims in (False, True):
                    with np.errstate(all="ignore"):
                        r = getattr(sparse, name)(x, axis=axis, keepdims=keepdims)
                    ax = axis if axis is None or isinstance(axis, int) else list(axis)
                    bk.add({"op": name, "dtype": dt, "axis": ax, "keepdims": keepdims, "fmt": "coo", "note": note},
                           **enc("a_", x), **enc("out_", r))
    # GCXS input goes through asCOO
    g = arrays[0][2].asformat("gcxs")
    for name in ("nansum", "nanmax", "nanmean"):
        for axis in (0, (0, 2), None):
            with np.errstate(all="ignore"):
                r = getattr(sparse, name)(g, axis=axis)
            ax = axis if axis is None or isinstance(axis, int) else list(axis)
            bk.add({"op": name, "dtype": "float64", "axis": ax, "keepdims": False, "fmt": "gcxs", "note": ""},
                   **enc("a_", arrays[0][2]), **enc("out_", r), a_ca=np.array(g.compressed_axes))
    bk.save()

# --------------------------------------------------------------------------
EINSUM_CASES = [
    "a,->a", "ab,->ab", ",ab,->ab", ",,->", "a,ab,abc->abc", "a,b,ab->ab", "ea,fb,gc,hd,abcd->efgh",
    "ea,fb,abcd,gc,hd->efgh", "abcd,ea,fb,gc,hd->efgh", "acdf,jbje,gihb,hfac,gfac,gifabc,hfac",
    "cd,bdhe,aidb,hgca,gc,hgibcd,hgac", "abhe,hidj,jgba,hiab,gab", "bde,cdh,agdb,hica,ibd,hgicd,hiac",
    "chd,bde,agbc,hiad,hgc,hgi,hiad", "chd,bde,agbc,hiad,bdi,cgh,agdb", "bdhe,acad,hiab,agac,hibd", "ab,ab,c->",
    "ab,ab,c->c", "ab,ab,cd,cd->", "ab,ab,cd,cd->ac", "ab,ab,cd,cd->cd", "ab,ab,cd,cd,ef,ef->", "ab,cd,ef->abcdef",
    "ab,cd,ef->acdf", "ab,cd,de->abcde", "ab,cd,de->be", "ab,bcd,cd->abcd", "ab,bcd,cd->abd", "eb,cb,fb->cef",
    "dd,fb,be,cdb->cef", "bca,cdb,dbf,afc->", "dcc,fce,ea,dbf->ab", "fdf,cdd,ccd,afe->ae", "abcd,ad",
    "ed,fcd,ff,bcf->be", "baa,dcf,af,cde->be", "bd,db,eac->ace", "fff,fae,bef,def->abd", "efc,dbc,acf,fd->abe",
    "ab,ab", "ab,ba", "abc,abc", "abc,bac", "abc,cba", "ab,bc", "ab,cb", "ba,bc", "ba,cb", "abcd,cd", "abcd,ab",
    "abcd,cdef", "abcd,cdef->feba", "abcd,efdc", "aab,bc->ac", "ab,bcc->ac", "aab,bcc->ac", "baa,bcc->ac",
    "aab,ccb->ac", "aab,fa,df,ecc->bde", "ecb,fef,bad,ed->ac", "bcf,bbb,fbf,fc->", "bb,ff,be->e", "bcb,bb,fc,fff->",
    "fbb,dfd,fc,fc->", "afd,ba,cc,dc->bf", "adb,bc,fa,cfc->d", "bbd,bda,fc,db->acf", "dba,ead,cad->bce",
    "aef,fbc,dca->bde", "abab->ba", "...ab,...ab", "...ab,...b->...a", "a...,a...", "ab->ba", "abc->", "abc->b",
    "aa", "aa->a", "aab->b",
]  # the reference's own test list (tests/test_einsum.py:7-82 upstream) + single-term forms


def gen_einsum():
    """einsum (_common.py:1150-1476; tests/test_einsum.py upstream)."""
    bk = Book("einsum_api")
    rng = np.random.default_rng(31)
    d = 4
    for sub in EINSUM_CASES:
        terms = sub.split("->")[0].split(",")
        for density in (0.3, 1.0):
            arrays = [sparse.random((d,) * len(t.replace("...", "xy")), density=density, random_state=rng)
                      for t in terms]
            r = sparse.einsum(sub, *arrays)
            ops = {}
            for i, a in enumerate(arrays):
                ops.update(enc(f"op{i}_", a))
            bk.add({"op": "einsum", "sub": sub, "n": len(arrays), "fmts": ["coo"] * len(arrays), "dtype": None,
                    "note": f"d{density}"}, **ops, **enc("out_", r))
    # operand formats -> result format (tests/test_einsum.py:143-183 upstream, without DOK)
    for fmts in (("coo",), ("gcxs",), ("coo", "coo"), ("coo", "dense"), ("dense", "coo"), ("gcxs", "dense"),
                 ("dense", "gcxs"), ("gcxs", "gcxs"), ("dense", "coo", "gcxs"), ("dense", "dense", "coo")):
        arrays = [sparse.random((3, 3, 3), density=0.5, random_state=rng) for _ in fmts]
        ins = [a.todense() if f == "dense" else a.asformat(f) for a, f in zip(arrays, fmts)]
        eq = {1: "abc->bc", 2: "abc,cda->abd", 3: "abc,cad,dea->abe"}[len(fmts)]
        r = sparse.einsum(eq, *ins)
        ops = {}
        for i, a in enumerate(arrays):
            ops.update(enc(f"op{i}_", a))
        bk.add({"op": "einsum", "sub": eq, "n": len(arrays), "fmts": list(fmts), "dtype": None, "note": "fmt"},
               **ops, **enc("out_", r))
    # dtype=
    x = (sparse.random((3, 3), density=0.5, random_state=rng) * 10.0).astype(np.float64)
    y = sparse.COO.from_numpy(np.ones((3, 1)))
    for dt in ("int64", "float32"):
        r = sparse.einsum("ij,i->j", x, y.reshape((3,)), dtype=np.dtype(dt))
        bk.add({"op": "einsum", "sub": "ij,i->j", "n": 2, "fmts": ["coo", "coo"], "dtype": dt, "note": "dtype"},
               **enc("op0_", x), **enc("op1_", y.reshape((3,))), **enc("out_", r))
    # interleaved (sublist) call form (tests/test_einsum.py:103-115 upstream); Ellipsis encoded as -1
    x = sparse.random((d, d), density=0.5, random_state=rng)
    for lists in ([[0, 0]], [[0, Ellipsis]], [[Ellipsis, 1], [Ellipsis]], [[0, 1], [0]], [[0, 1], [1, 0]]):
        r = sparse.einsum(x, *lists)
        enc_lists = [[-1 if s is Ellipsis else s for s in li] for li in lists]
        bk.add({"op": "einsum_lists", "lists": enc_lists, "n": 1, "fmts": ["coo"], "dtype": None, "note": "lists"},
               **enc("op0_", x), **enc("out_", r))
    bk.save()

# --------------------------------------------------------------------------
def gen_io():
    """Files written by the reference's save_npz (_io.py:7-66): the on-disk format load_npz must read."""
    rng = np.random.default_rng(37)
    x = sparse.random((5, 6, 7), density=0.2, random_state=rng, fill_value=0.5)
    sparse.save_npz(os.path.join(HERE, "ref_saved_coo.npz"), x)
    g = sparse.random((6, 8), density=0.3, random_state=rng).astype(np.float32).asformat("gcxs", compressed_axes=(1,))
    sparse.save_npz(os.path.join(HERE, "ref_saved_gcxs.npz"), g, compressed=False)
    print("ref_saved_coo.npz, ref_saved_gcxs.npz written")

# --------------------------------------------------------------------------
def _enc_index(index):
    """JSON form of a basic index: int | None | "..." | [start, stop, step]."""
    if not isinstance(index, tuple):
        index = (index,)
    out = []
    for i in index:
        if i is Ellipsis:
            out.append("...")
        elif isinstance(i, slice):
            out.append([i.start, i.stop, i.step])
        else:
            out.append(i)
    return out


INDEX_CASES = [
    0, 1, -1, (1, 1, 1), (slice(0, 2),), (slice(None, 2), slice(None, 2)), (slice(1, None), slice(1, None)),
    (slice(None, None),), (slice(None, None, -1),), (slice(None, 2, -1), slice(None, 2, -1)),
    (slice(1, None, 2), slice(1, None, 2)), (slice(None, None, 2),), (slice(None, 2, -1), slice(None, 2, -2)),
    (slice(1, None, 2), slice(1, None, 1)), (slice(None, None, -2),), (0, slice(0, 2)), (slice(0, 1), 0),
    (None, slice(1, 3), 0), (slice(0, 3), None, 0), (slice(1, 2), slice(2, 4)), (slice(1, 2), slice(None, None)),
    (slice(1, 2), slice(None, None), 2), (slice(1, 2, 2), slice(None, None), 2),
    (slice(1, 2, None), slice(None, None, 2), 2), (slice(1, 2, -2), slice(None, None), -2),
    (slice(1, 2, None), slice(None, None, -2), 2), (slice(1, 2, -1), slice(None, None), -1),
    (slice(1, 2, None), slice(None, None, -1), 2), (slice(2, 0, -1), slice(None, None), -1),
    (slice(-2, None, None),), (slice(-1, None, None), slice(-2, None, None)), (Ellipsis, slice(1, 3)),
    (1, Ellipsis, slice(1, 3)), (slice(0, 1), Ellipsis), (Ellipsis, None), (None, Ellipsis), (1, Ellipsis),
    (1, Ellipsis, None), (1, 1, 1, Ellipsis), (Ellipsis, 1, None), (slice(None, 1000),),
    (slice(None), slice(None, 1000)), (slice(None), slice(1000, -1000, -1)), (slice(None), slice(1000, -1000, -50)),
    (slice(5, 0),), (slice(0, 5, -1),), (slice(0, 0, None),),
    (slice(None), 2, slice(None, None, -1)), (slice(None, None, -1), slice(None, None, -1), slice(None, None, -1)),
    (4, slice(1, 6, 2), slice(None)), (slice(None), slice(None), 3), (None, None, 2), (2, None, slice(None), None, 1),
]  # tests/test_coo.py:408-469 upstream (basic forms) + a few more


def gen_indexing():
    """Basic indexing of COO and GCXS (_coo/indexing.py:12-133, _compressed/indexing.py:14-174)."""
    bk = Book("indexing_api")
    rng = np.random.default_rng(41)
    for shape, density in (((2, 3, 4), 0.5), ((5, 7, 6), 0.3)):
        x = sparse.random(shape, density=density, random_state=rng)
        for index in INDEX_CASES:
            try:
                want = x.todense()[index]
            except IndexError:
                continue
            r = x[index]
            assert np.array_equal(np.asarray(want), r.todense() if hasattr(r, "todense") else r)
            bk.add({"op": "getitem", "fmt": "coo", "index": _enc_index(index)}, **enc("a_", x), **enc("out_", r))
        for ca in ((0,), (1,), (0, 1), (2,)):
            g = x.asformat("gcxs", compressed_axes=ca)
            for index in INDEX_CASES:
                try:
                    x.todense()[index]
                    r = g[index]
                except (IndexError, ValueError):
                    continue
                check = "full"
                if isinstance(r, sparse.GCXS):
                    # canonical within-row order for comparison (negative steps leave reversed rows upstream)
                    try:
                        r = sparse.GCXS.from_coo(r.tocoo(), compressed_axes=r.compressed_axes)
                    except Exception:
                        # upstream quirk: a 1-D selection followed by None insertion keeps indptr=None and cannot be
                        # converted; pin the values and the compressed axes only
                        check = "dense"
                        want_ca = r.compressed_axes
                        r = sparse.GCXS.from_numpy(x.todense()[index], compressed_axes=want_ca)
                bk.add({"op": "getitem", "fmt": "gcxs", "index": _enc_index(index), "check": check}, **enc("a_", x),
                       **enc("out_", r), a_ca=np.array(ca))
    # non-zero fill value and an element lookup that misses
    x = sparse.random((4, 5), density=0.4, random_state=rng, fill_value=7.0)
    for index in ((slice(1, 3), slice(None, None, 2)), (0, 0), (3, 4), (2,), (Ellipsis, 1)):
        bk.add({"op": "getitem", "fmt": "coo", "index": _enc_index(index)}, **enc("a_", x), **enc("out_", x[index]))
    bk.save()

# --------------------------------------------------------------------------
def gen_elemwise_nary():
    """n-ary / user-defined functions through elemwise (tests/test_elemwise.py:252-305 upstream)."""
    sys.path.insert(0, os.path.dirname(HERE))
    import _nary_funcs as NF

    bk = Book("elemwise_nary_api")
    rng = np.random.default_rng(43)
    shape_sets = [[(2,), (3, 2), (4, 3, 2)], [(3,), (2, 3), (2, 2, 3)], [(2,), (2, 2), (2, 2, 2)],
                  [(4,), (4, 4), (4, 4, 4)], [(1, 1, 2), (1, 3, 1), (4, 1, 1)], [(2,), (2, 1), (2, 1, 1)]]
    for shapes in shape_sets:
        args = [sparse.random(sh, density=0.5, random_state=rng) for sh in shapes]
        for fi, f in enumerate(NF.TRINARY):
            r = sparse.elemwise(f, *args)
            ops = {}
            for i, a in enumerate(args):
                ops.update(enc(f"op{i}_", a))
            bk.add({"op": "trinary", "func": fi, "n": 3, "note": ""}, **ops, **enc("out_", r))
    patho = [([(2,), (3, 2), (4, 3, 2)], 0), ([(3,), (2, 3), (2, 2, 3)], 1), ([(2,), (2, 2), (2, 2, 2)], 2),
             ([(4,), (4, 4), (4, 4, 4)], 3)]
    for shapes, fi in patho:
        for value in (np.nan, np.inf, -np.inf):
            for fraction in (0.25, 0.5, 1.0):
                def rvs(n, value=value, fraction=fraction):
                    i = int(n * fraction)
                    ar = np.empty((n,), dtype=np.float64)
                    ar[:i] = value
                    ar[i:] = rng.random(n - i)
                    return ar
                args = [sparse.random(sh, density=0.5, random_state=rng, data_rvs=rvs) for sh in shapes]
                with np.errstate(all="ignore"):
                    r = sparse.elemwise(NF.TRINARY[fi], *args)
                ops = {}
                for i, a in enumerate(args):
                    ops.update(enc(f"op{i}_", a))
                bk.add({"op": "trinary", "func": fi, "n": 3, "note": f"patho-{value}-{fraction}"}, **ops,
                       **enc("out_", r))
    # three-operand where (tests/test_coo.py:955-984 upstream) + signed zeros / NaN / inf / non-zero fills / scalars
    where_shapes = [[(2,), (3, 2), (4, 3, 2)], [(3,), (2, 3), (2, 2, 3)], [(4,), (4, 4), (4, 4, 4)],
                    [(1, 1, 2), (1, 3, 1), (4, 1, 1)], [(2,), (2, 1), (2, 1, 1)], [(3,), (), (2, 3)], [(4, 4), (), ()]]
    for shapes in where_shapes:
        for dt in ("float64", "float32", "int64"):
            cs = sparse.random(shapes[0], density=0.5, random_state=rng).astype(np.bool_)
            xs = rand_sparse(rng, shapes[1], 0.5, dt)
            ys = rand_sparse(rng, shapes[2], 0.5, dt)
            r = sparse.where(cs, xs, ys)
            bk.add({"op": "where", "func": -1, "n": 3, "note": f"{dt}"}, **enc("op0_", cs), **enc("op1_", xs),
                   **enc("op2_", ys), **enc("out_", r))
    cs = sparse.random((6, 7), density=0.5, random_state=rng).astype(np.bool_)
    xs = sparse.random((6, 7), density=0.6, random_state=rng)
    ys = sparse.random((6, 7), density=0.6, random_state=rng)
    dx, dy = xs.data.copy(), ys.data.copy()
    dx[::4], dx[1::4], dy[::3], dy[1::5] = -0.0, np.nan, np.inf, -0.0
    xs2 = sparse.COO(xs.coords, dx, shape=xs.shape)
    ys2 = sparse.COO(ys.coords, dy, shape=ys.shape, fill_value=2.5)
    for note, args in (("special", (cs, xs2, ys2)), ("scalar_x", (cs, 0, ys2)), ("scalar_y", (cs, xs2, -1.5)),
                       ("float_cond", (xs2, xs, ys))):
        with np.errstate(all="ignore"):
            r = sparse.where(*args)
        ops = {}
        for i, a in enumerate(args):
            ops.update(enc(f"op{i}_", a))
        bk.add({"op": "where", "func": -1, "n": 3, "note": note}, **ops, **enc("out_", r))
    x = sparse.random((5, 6), density=0.4, random_state=rng)
    y = sparse.random((5, 6), density=0.4, random_state=rng)
    for fi, f in enumerate(NF.UNARY_BINARY):
        n = f.__code__.co_argcount
        args = [x, y][:n]
        with np.errstate(all="ignore"):
            r = sparse.elemwise(f, *args)
        ops = {}
        for i, a in enumerate(args):
            ops.update(enc(f"op{i}_", a))
        bk.add({"op": "unary_binary", "func": fi, "n": n, "note": ""}, **ops, **enc("out_", r))
    bk.save()


# --------------------------------------------------------------------------
def gen_formats():
    """COO canonicalisation, COO<->GCXS, transpose/reshape (a-1, a-2, a-19)."""
    bk = Book("formats_api")
    rng = np.random.default_rng(17)
    # unsorted, duplicated, zero-containing raw triplets -> canonical COO
    for dt in ("float64", "float32", "int64"):
        for shape in ((6, 7), (3, 4, 5), (20,), (2, 3, 2, 3)):
            n = 40
            coords = np.stack([rng.integers(0, s, n) for s in shape])
            data = (rng.integers(-3, 4, n)).astype(dt)
            data_f = data.copy()
            if dt != "int64":
                data_f = data_f * 0.5
                data_f[::7] = -0.0
            c = sparse.COO(coords, data_f, shape=shape, prune=True)
            bk.add({"op": "coo_ctor", "dtype": dt, "shape": shape}, coords=coords, data=data_f, **enc("out_", c))
            c2 = sparse.COO(coords, data_f, shape=shape)
            bk.add({"op": "coo_ctor_noprune", "dtype": dt, "shape": shape}, coords=coords, data=data_f,
                   **enc("out_", c2))
    for shape in ((6, 7), (3, 4, 5), (2, 3, 4, 5)):
        x = rand_sparse(rng, shape, 0.4, "float64")
        nd = len(shape)
        cas = [(0,), (nd - 1,), None] + ([(0, 1), (1,), (0, 2)] if nd >= 3 else [])
        for ca in cas:
            g = sparse.GCXS(x, compressed_axes=ca)
            bk.add({"op": "from_coo", "ca": None if ca is None else list(ca)}, **enc("a_", x), **enc("out_", g))
            bk.add({"op": "tocoo", "ca": None if ca is None else list(ca)}, **enc("a_", x), **enc("out_", g.tocoo()))
        perm = tuple(rng.permutation(nd).tolist())
        bk.add({"op": "transpose", "axes": perm}, **enc("a_", x), **enc("out_", x.transpose(perm)))
        bk.add({"op": "reshape", "shape": (-1, shape[-1])}, **enc("a_", x), **enc("out_", x.reshape((-1, shape[-1]))))
    bk.save()


# --------------------------------------------------------------------------
def gen_examples():
    """SDDMM / MTTKRP example paths (examples/sddmm_example.py:43-55, mttkrp_example.py:43-55)."""
    bk = Book("examples_api")
    rng = np.random.default_rng(19)
    for dt in ("float32", "float64"):
        for (L, K, dens) in ((40, 8, 0.1), (64, 256, 0.05), (33, 17, 0.3)):
            s = rand_sparse(rng, (L, L), dens, dt)
            a = rng.random((L, K)).astype(dt)
            b = rng.random((K, L)).astype(dt)
            r = s * (a @ b)
            bk.add({"op": "sddmm", "dtype": dt}, **enc("s_", s), a=a, b=b, **enc("out_", r))
    for dt in ("float64", "float32"):
        I_, K_, L_, J_ = 12, 10, 6, 5
        B = rand_sparse(rng, (I_, K_, L_), 0.1, dt).asformat("gcxs")
        D = rng.random((L_, J_)).astype(dt)
        Cm = rng.random((K_, J_)).astype(dt)
        r = sparse.sum(B[:, :, :, None] * D[None, None, :, :] * Cm[None, :, None, :], axis=(1, 2))
        bk.add({"op": "mttkrp", "dtype": dt}, **enc("B_", B.tocoo()), D=D, C=Cm, **enc("out_", r),
               B_ca=np.array(B.compressed_axes))
    bk.save()


def gen_complex():
    """Complex operands through the public API (tests/test_dot.py:303-335, tests/test_coo.py:1318-1332 upstream):
    matmul over formats / return types, element-wise arithmetic, sum."""
    bk = Book("complex_api")
    rng = np.random.default_rng(21)

    def crand(shape, dt, density=0.5):
        re = sparse.random(shape, density=density, random_state=rng)
        im = sparse.random(shape, density=density, random_state=rng)
        return (re + im * 1j).astype(dt)

    fmts = [("coo", "coo"), ("coo", "gcxs"), ("gcxs", "coo"), ("gcxs", "gcxs"), ("coo", "dense"), ("dense", "coo"),
            ("gcxs", "dense"), ("dense", "gcxs")]
    for dt1, dt2 in (("complex128", "complex128"), ("complex64", "complex64"), ("complex64", "complex128"),
                     ("complex128", "float64"), ("float32", "complex64")):
        for sa, sb in (((6, 7), (7, 5)), ((7,), (7, 5)), ((6, 7), (7,)), ((3, 4, 5), (5, 6))):
            a = crand(sa, dt1) if dt1.startswith("c") else rand_sparse(rng, sa, 0.5, dt1)
            b = crand(sb, dt2) if dt2.startswith("c") else rand_sparse(rng, sb, 0.5, dt2)
            for fa, fb in fmts:
                xa = a.todense() if fa == "dense" else a.asformat(fa)
                xb = b.todense() if fb == "dense" else b.asformat(fb)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    r = sparse.matmul(xa, xb)
                bk.add({"op": "matmul", "dt1": dt1, "dt2": dt2, "fa": fa, "fb": fb}, **enc("a_", a), **enc("b_", b),
                       **enc("out_", r))
    for dt in ("complex128", "complex64"):
        a, b = crand((4, 5, 3), dt), crand((4, 5, 3), dt)
        y = rand_sparse(rng, (5, 1), 0.6, "float64" if dt == "complex128" else "float32")
        for name, f in (("add", lambda u, v: u + v), ("subtract", lambda u, v: u - v), ("multiply", lambda u, v: u * v)):
            bk.add({"op": name, "dtype": dt, "kind": "cc"}, **enc("a_", a), **enc("b_", b), **enc("out_", f(a, b)))
            bk.add({"op": name, "dtype": dt, "kind": "cr"}, **enc("a_", a), **enc("b_", y), **enc("out_", f(a, y)))
        bk.add({"op": "scale", "dtype": dt}, **enc("a_", a), **enc("out_", a * (2 - 3j)))
        bk.add({"op": "conj", "dtype": dt}, **enc("a_", a), **enc("out_", a.conj()))
        bk.add({"op": "real", "dtype": dt}, **enc("a_", a), **enc("out_", a.real))
        bk.add({"op": "imag", "dtype": dt}, **enc("a_", a), **enc("out_", a.imag))
        bk.add({"op": "abs", "dtype": dt}, **enc("a_", a), **enc("out_", abs(a)))
        for axis in (None, 0, (1, 2)):
            bk.add({"op": "sum", "dtype": dt, "axis": axis}, **enc("a_", a), **enc("out_", a.sum(axis=axis)))
    bk.save()


if __name__ == "__main__":
    which = sys.argv[1:] or ["dot", "tensordot", "elemwise", "reduce", "nanreduce", "einsum", "io", "indexing", "nary", "formats", "examples", "complex"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if "dot" in which:
            gen_dot_kernels()
        if "tensordot" in which:
            gen_tensordot()
        if "elemwise" in which:
            gen_elemwise()
        if "reduce" in which:
            gen_reduce()
        if "nanreduce" in which:
            gen_nanreduce()
        if "einsum" in which:
            gen_einsum()
        if "io" in which:
            gen_io()
        if "indexing" in which:
            gen_indexing()
        if "nary" in which:
            gen_elemwise_nary()
        if "formats" in which:
            gen_formats()
        if "examples" in which:
            gen_examples()
        if "complex" in which:
            gen_complex()

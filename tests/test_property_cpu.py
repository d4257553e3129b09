"""Property tests of the host logic (broadcast rules, fill values, axis bookkeeping, index normalisation) against dense
NumPy, with the NumPy mock of the kernel layer: random shapes / densities / fill values / operators drawn by
hypothesis.  CPU only -- the kernels themselves are covered by the golden-vector tests on the GPU."""
import os

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import _mock_kernels

pytestmark = pytest.mark.skipif(torch.cuda.is_available(), reason="mock backend is only used on boxes without a GPU")

# The gate run (driver: `pytest -m "not gpu"`) is DERANDOMISED so that it cannot flake; the exploratory campaign is
#   B2S_HYP_EXAMPLES=3000 python -m pytest tests/test_property_cpu.py   (random seeds; failures become regular tests)
_N = int(os.environ.get("B2S_HYP_EXAMPLES", "0"))
SET = settings(max_examples=_N or 60, deadline=None, suppress_health_check=list(HealthCheck), derandomize=not _N,
               database=None)


@pytest.fixture(autouse=True)
def _mock():
    _mock_kernels.install()
    yield
    _mock_kernels.uninstall()


def _sp():
    import sparse_b200

    return sparse_b200


shapes = st.lists(st.integers(1, 5), min_size=1, max_size=4).map(tuple)


def _rand(shape, seed, density, fill=0.0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    d = np.full(shape, fill, dtype=dtype)
    mask = rng.random(shape) < density
    vals = rng.integers(-4, 5, size=shape).astype(dtype)
    d[mask] = vals[mask]
    return _sp().COO.from_numpy(d, fill_value=fill), d


def _broadcastable_pair(draw):
    base = draw(shapes)
    other = tuple(draw(st.sampled_from([s, 1])) for s in base)
    k = draw(st.integers(0, len(other)))
    return base, other[k:]


@SET
@given(st.data())
def test_binary_broadcast_matches_numpy(data):
    a_shape, b_shape = _broadcastable_pair(data.draw)
    fa, fb = data.draw(st.sampled_from([0.0, 1.0, -2.0])), data.draw(st.sampled_from([0.0, 3.0]))
    a, da = _rand(a_shape, data.draw(st.integers(0, 99)), data.draw(st.sampled_from([0.0, 0.3, 1.0])), fa)
    b, db = _rand(b_shape, data.draw(st.integers(0, 99)), data.draw(st.sampled_from([0.0, 0.5, 1.0])), fb)
    f = data.draw(st.sampled_from([np.add, np.subtract, np.multiply, np.maximum, np.minimum, np.greater, np.not_equal]))
    if data.draw(st.booleans()):
        a, da, b, db = b, db, a, da
    got = f(a, b)
    want = f(da, db)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert np.array_equal(got.todense(), want)
    # canonical: exactly the entries whose BITS differ from the fill value are stored (-0.0 is kept, like upstream)
    bits = want.view(np.uint8).reshape(want.size, -1) if want.size else np.zeros((0, 1), np.uint8)
    fbits = np.asarray(got.fill_value).reshape(1).view(np.uint8)
    assert got.nnz == int((bits != fbits).any(axis=1).sum())


@SET
@given(st.data())
def test_reduce_matches_numpy(data):
    shape = data.draw(shapes)
    fill = data.draw(st.sampled_from([0.0, 0.0, 2.0]))
    x, d = _rand(shape, data.draw(st.integers(0, 99)), data.draw(st.sampled_from([0.0, 0.4, 1.0])), fill)
    nd = len(shape)
    axis = data.draw(st.one_of(st.none(), st.integers(-nd, nd - 1),
                               st.lists(st.integers(0, nd - 1), min_size=1, max_size=nd, unique=True).map(tuple)))
    keepdims = data.draw(st.booleans())
    name = data.draw(st.sampled_from(["sum", "max", "min", "prod", "any", "all", "mean"]))
    if name in ("any", "all") and fill != 0.0:
        # the reference refuses these (logical_or.reduce([fill, fill]) is a bool, never "equivalent" to a float fill)
        with pytest.raises(ValueError):
            getattr(x, name)(axis=axis, keepdims=keepdims)
        return
    with np.errstate(all="ignore"):
        got = getattr(x, name)(axis=axis, keepdims=keepdims)
        want = getattr(d, name)(axis=axis, keepdims=keepdims)
    got_d = got.todense() if hasattr(got, "todense") else np.asarray(got)
    assert got_d.shape == np.shape(want)
    assert np.allclose(got_d, want, rtol=1e-12, atol=1e-12, equal_nan=True)


@SET
@given(st.data())
def test_getitem_matches_numpy(data):
    shape = data.draw(shapes)
    x, d = _rand(shape, data.draw(st.integers(0, 99)), data.draw(st.sampled_from([0.0, 0.5, 1.0])),
                 data.draw(st.sampled_from([0.0, 7.0])))
    index = []
    for s in shape:
        kind = data.draw(st.sampled_from(["int", "slice", "slice", "full", "none+slice"]))
        if kind == "int":
            index.append(data.draw(st.integers(-s, s - 1)))
        elif kind == "full":
            index.append(slice(None))
        else:
            if kind == "none+slice":
                index.append(None)
            index.append(slice(data.draw(st.one_of(st.none(), st.integers(-s - 1, s + 1))),
                               data.draw(st.one_of(st.none(), st.integers(-s - 1, s + 1))),
                               data.draw(st.sampled_from([None, 1, 2, 3, -1, -2]))))
    cut = data.draw(st.integers(0, len(index)))
    index = tuple(index[:cut]) + ((Ellipsis,) if data.draw(st.booleans()) else ())
    n_real = sum(1 for i in index if i is not None and i is not Ellipsis)
    want = d[index]
    got = x[index]
    if np.ndim(want) == 0 and not hasattr(got, "todense"):
        assert got == want
        return
    assert got.shape == want.shape and n_real <= len(shape)
    assert np.array_equal(got.todense(), want)
    assert got.fill_value == x.fill_value


@SET
@given(st.data())
def test_tensordot_axes_match_numpy(data):
    a_shape = data.draw(st.lists(st.integers(1, 4), min_size=1, max_size=3).map(tuple))
    n_con = data.draw(st.integers(0, len(a_shape)))
    a_axes = data.draw(st.permutations(range(len(a_shape))))[:n_con]
    b_free = data.draw(st.lists(st.integers(1, 4), min_size=0, max_size=2))
    b_shape_list = [a_shape[ax] for ax in a_axes] + b_free
    perm = data.draw(st.permutations(range(len(b_shape_list))))
    b_shape = tuple(b_shape_list[p] for p in perm)
    b_axes = [perm.index(i) for i in range(n_con)]
    if not b_shape:
        return
    a, da = _rand(a_shape, data.draw(st.integers(0, 99)), 0.6)
    b, db = _rand(b_shape, data.draw(st.integers(0, 99)), 0.6)
    fmt = data.draw(st.sampled_from(["coo", "gcxs", "dense"]))
    bb = db if fmt == "dense" else b.asformat(fmt)
    got = _sp().tensordot(a, bb, axes=(list(a_axes), b_axes))
    want = np.tensordot(da, db, axes=(list(a_axes), b_axes))
    got_d = got.todense() if hasattr(got, "todense") else np.asarray(got)
    assert got_d.shape == want.shape
    assert np.allclose(got_d, want, rtol=1e-12, atol=1e-12)


@SET
@given(st.data())
def test_transpose_reshape_roundtrip(data):
    shape = data.draw(shapes)
    x, d = _rand(shape, data.draw(st.integers(0, 99)), 0.5)
    perm = tuple(data.draw(st.permutations(range(len(shape)))))
    assert np.array_equal(x.transpose(perm).todense(), d.transpose(perm))
    g = x.asformat("gcxs")
    assert np.array_equal(g.transpose(perm).todense(), d.transpose(perm))
    size = int(np.prod(shape))
    divs = [k for k in range(1, size + 1) if size % k == 0]
    k = data.draw(st.sampled_from(divs))
    assert np.array_equal(x.reshape((k, size // k)).todense(), d.reshape((k, size // k)))
    assert np.array_equal(g.reshape((size // k, k)).todense(), d.reshape((size // k, k)))


@SET
@given(st.data())
def test_einsum_matches_numpy(data):
    labels = "abcd"
    sizes = {ch: data.draw(st.integers(1, 3)) for ch in labels}
    n_ops = data.draw(st.integers(1, 3))
    terms = ["".join(data.draw(st.lists(st.sampled_from(labels), min_size=0, max_size=3))) for _ in range(n_ops)]
    used = sorted(set("".join(terms)))
    explicit = data.draw(st.booleans())
    sub = ",".join(terms)
    if explicit:
        out = "".join(data.draw(st.permutations(used))[: data.draw(st.integers(0, len(used)))])
        sub += "->" + out
    ops, dense = [], []
    for t in terms:
        x, d = _rand(tuple(sizes[ch] for ch in t), data.draw(st.integers(0, 99)), 0.6)
        fmt = data.draw(st.sampled_from(["coo", "coo", "gcxs"]))
        ops.append(x.asformat(fmt) if x.ndim else x)
        dense.append(d)
    want = np.einsum(sub, *dense)
    got = _sp().einsum(sub, *ops)
    got_d = got.todense() if hasattr(got, "todense") else np.asarray(got)
    assert got_d.shape == want.shape
    assert np.allclose(got_d, want, rtol=1e-12, atol=1e-12)


@SET
@given(st.data())
def test_nan_reductions_match_numpy(data):
    shape = data.draw(shapes)
    rng = np.random.default_rng(data.draw(st.integers(0, 99)))
    fill = data.draw(st.sampled_from([0.0, 0.0, np.nan]))
    d = np.full(shape, fill)
    m = rng.random(shape) < 0.6
    d[m] = rng.integers(-3, 4, size=shape)[m]
    d[rng.random(shape) < 0.2] = np.nan
    x = _sp().COO.from_numpy(d, fill_value=fill)
    nd = len(shape)
    axis = data.draw(st.one_of(st.none(), st.integers(0, nd - 1)))
    keepdims = data.draw(st.booleans())
    name = data.draw(st.sampled_from(["nansum", "nanprod", "nanmax", "nanmin", "nanmean"]))
    import warnings

    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = getattr(np, name)(d, axis=axis, keepdims=keepdims)
        got = getattr(_sp(), name)(x, axis=axis, keepdims=keepdims)
    got_d = got.todense() if hasattr(got, "todense") else np.asarray(got)
    assert got_d.shape == np.shape(want)
    assert np.allclose(got_d, want, rtol=1e-12, atol=1e-12, equal_nan=True)


@SET
@given(st.data())
def test_where_and_composites_match_numpy(data):
    shape = data.draw(shapes)
    c, dc = _rand(shape, data.draw(st.integers(0, 99)), 0.5)
    x, dx = _rand(shape, data.draw(st.integers(0, 99)), 0.6, data.draw(st.sampled_from([0.0, 1.5])))
    y, dy = _rand(shape[-1:], data.draw(st.integers(0, 99)), 0.6)
    sp = _sp()
    assert np.array_equal(sp.where(c, x, y).todense(), np.where(dc, dx, dy))
    assert np.array_equal(sp.where(c != 0, 2.5, y).todense(), np.where(dc != 0, 2.5, dy))
    f = data.draw(st.sampled_from([lambda p, q, r: (p + q) * r - p, lambda p, q, r: abs(p - r) * (q > 0)]))
    assert np.array_equal(sp.elemwise(f, c, x, y).todense(), f(dc, dx, dy))


@SET
@given(st.data())
def test_sparse_dense_elemwise_and_matmul(data):
    shape = data.draw(st.lists(st.integers(1, 4), min_size=2, max_size=3).map(tuple))
    x, dx = _rand(shape, data.draw(st.integers(0, 99)), 0.5)
    rng = np.random.default_rng(data.draw(st.integers(0, 99)))
    dn = rng.integers(1, 5, size=shape[-1:]).astype(np.float64)
    got = x * dn
    assert isinstance(got, _sp().COO) and np.array_equal(got.todense(), dx * dn)
    k = data.draw(st.integers(1, 3))
    w = rng.integers(-2, 3, size=(shape[-1], k)).astype(np.float64)
    assert np.allclose(np.asarray(x @ w), dx @ w, rtol=1e-12, atol=1e-12)
    ws, _ = _rand((shape[-1], k), data.draw(st.integers(0, 99)), 0.7)
    assert np.allclose((x @ ws).todense(), dx @ ws.todense(), rtol=1e-12, atol=1e-12)


@SET
@given(st.data())
def test_matmul_and_dot_broadcasting(data):
    batch = data.draw(st.lists(st.integers(1, 3), min_size=0, max_size=2).map(tuple))
    m, k, n = (data.draw(st.integers(1, 4)) for _ in range(3))
    b_batch = tuple(data.draw(st.sampled_from([s, 1])) for s in batch)[data.draw(st.integers(0, len(batch))):]
    a, da = _rand(batch + (m, k), data.draw(st.integers(0, 99)), 0.6)
    b, db = _rand(b_batch + (k, n), data.draw(st.integers(0, 99)), 0.6)
    sp = _sp()
    fa = data.draw(st.sampled_from(["coo", "gcxs"]))
    fb = data.draw(st.sampled_from(["coo", "gcxs", "dense"]))
    A = a.asformat(fa)
    B = db if fb == "dense" else b.asformat(fb)
    got = sp.matmul(A, B)
    want = np.matmul(da, db)
    got_d = got.todense() if hasattr(got, "todense") else np.asarray(got)
    assert got_d.shape == want.shape and np.allclose(got_d, want, rtol=1e-12, atol=1e-12)
    got = sp.dot(A, B)
    want = np.dot(da, db)
    got_d = got.todense() if hasattr(got, "todense") else np.asarray(got)
    assert got_d.shape == want.shape and np.allclose(got_d, want, rtol=1e-12, atol=1e-12)


@SET
@given(st.data())
def test_gcxs_views_match_numpy(data):
    shape = data.draw(st.lists(st.integers(1, 4), min_size=2, max_size=4).map(tuple))
    x, d = _rand(shape, data.draw(st.integers(0, 99)), 0.5)
    nd = len(shape)
    ca = tuple(sorted(data.draw(st.lists(st.integers(0, nd - 1), min_size=1, max_size=nd - 1, unique=True))))
    g = x.asformat("gcxs", compressed_axes=ca)
    assert g.compressed_axes == ca and np.array_equal(g.todense(), d)
    ca2 = tuple(sorted(data.draw(st.lists(st.integers(0, nd - 1), min_size=1, max_size=nd - 1, unique=True))))
    g2 = g.change_compressed_axes(ca2)
    assert g2.compressed_axes == ca2 and np.array_equal(g2.todense(), d)
    index = tuple(data.draw(st.sampled_from([slice(None), slice(0, 1), slice(None, None, -1), 0, -1, slice(1, None, 2)]))
                  for _ in range(data.draw(st.integers(1, nd))))
    want = d[index]
    got = g[index]
    if np.ndim(want) == 0:
        assert got == want
    else:
        assert isinstance(got, _sp().GCXS) and np.array_equal(got.todense(), want)
    tgt = data.draw(st.sampled_from([shape, (2,) + shape, tuple(3 if e == 1 else e for e in shape)]))
    assert np.array_equal(x.broadcast_to(tgt).todense(), np.broadcast_to(d, tgt))
    assert np.array_equal(g.astype(np.float32).todense(), d.astype(np.float32))


# ---- array manipulation (sparse_b200/_manip.py): random shapes, axes, fill values --------------------------------------
def _check(got, want):
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got.todense(), want)
    c = got.asformat("coo")
    # stored entries = everything that differs BITWISE from the fill value (a product -3 * 0 leaves a stored -0.0,
    # as upstream's `equivalent` does)
    same = (want == c.fill_value) & (np.signbit(want) == np.signbit(c.fill_value))
    assert c.nnz == int(np.sum(~same))
    if c.ndim and c.nnz:
        assert np.all(np.diff(np.ravel_multi_index(tuple(c.coords), c.shape)) > 0)


@SET
@given(st.data())
def test_concatenate_stack_match_numpy(data):
    sp = _sp()
    shape = data.draw(shapes)
    fill = data.draw(st.sampled_from([0.0, 2.0]))
    n = data.draw(st.integers(1, 3))
    axis = data.draw(st.integers(-len(shape), len(shape) - 1))
    parts, dense = [], []
    for i in range(n):
        s = list(shape)
        s[axis] = data.draw(st.integers(0, 4))
        a, d = _rand(tuple(s), data.draw(st.integers(0, 99)), data.draw(st.sampled_from([0.0, 0.4, 1.0])), fill)
        parts.append(a if data.draw(st.booleans()) or a.ndim < 2 else sp.GCXS(a))
        dense.append(d)
    _check(sp.concatenate(parts, axis=axis), np.concatenate(dense, axis=axis))
    same = [_rand(shape, data.draw(st.integers(0, 99)), 0.5, fill) for _ in range(n)]
    sax = data.draw(st.integers(-len(shape) - 1, len(shape)))
    _check(sp.stack([a for a, _ in same], axis=sax), np.stack([d for _, d in same], axis=sax))


@SET
@given(st.data())
def test_roll_flip_pad_match_numpy(data):
    sp = _sp()
    shape = data.draw(shapes)
    a, d = _rand(shape, data.draw(st.integers(0, 99)), data.draw(st.sampled_from([0.0, 0.5, 1.0])),
                 data.draw(st.sampled_from([0.0, -1.0])))
    axis = data.draw(st.one_of(st.none(), st.integers(-len(shape), len(shape) - 1)))
    shift = data.draw(st.integers(-7, 7))
    _check(sp.roll(a, shift, axis), np.roll(d, shift, axis))
    _check(sp.flip(a, axis=axis), np.flip(d, axis=axis))
    pw = [(data.draw(st.integers(0, 2)), data.draw(st.integers(0, 2))) for _ in shape]
    _check(sp.pad(a, pw, constant_values=a.fill_value), np.pad(d, pw, constant_values=a.fill_value))


@SET
@given(st.data())
def test_diagonal_take_match_numpy(data):
    sp = _sp()
    shape = data.draw(st.lists(st.integers(1, 5), min_size=2, max_size=4).map(tuple))
    a, d = _rand(shape, data.draw(st.integers(0, 99)), data.draw(st.sampled_from([0.3, 1.0])))
    a1 = data.draw(st.integers(0, len(shape) - 1))
    a2 = data.draw(st.integers(0, len(shape) - 1).filter(lambda v: v != a1))
    k = data.draw(st.integers(-4, 4))
    _check(sp.diagonal(a, k, a1, a2), np.diagonal(d, k, a1, a2))
    _check(sp.triu(a, k), np.triu(d, k))
    _check(sp.tril(a, k), np.tril(d, k))
    ax = data.draw(st.integers(0, len(shape) - 1))
    idx = data.draw(st.lists(st.integers(-shape[ax], shape[ax] - 1), min_size=0, max_size=6))
    _check(sp.take(a, idx, axis=ax), np.take(d, np.asarray(idx, dtype=np.int64), axis=ax))
    index = tuple(np.asarray(idx, dtype=np.int64) if p == ax else
                  data.draw(st.sampled_from([slice(None), slice(None, None, -1), slice(1, None)])) for p in range(len(shape)))
    _check(a[index], d[index])


@SET
@given(st.data())
def test_kron_tile_repeat_match_numpy(data):
    sp = _sp()
    sa = data.draw(st.lists(st.integers(1, 3), min_size=1, max_size=3).map(tuple))
    sb = data.draw(st.lists(st.integers(1, 3), min_size=1, max_size=3).map(tuple))
    a, da = _rand(sa, data.draw(st.integers(0, 99)), 0.5)
    b, db = _rand(sb, data.draw(st.integers(0, 99)), 0.5)
    k = sp.kron(a, b)  # stored x stored only, as upstream (_coo/common.py:67-129): no -0.0 from negative * fill
    assert k.shape == np.kron(da, db).shape and np.array_equal(k.todense(), np.kron(da, db))
    assert k.nnz == a.nnz * b.nnz and not np.signbit(k.todense()[np.kron(da, db) == 0]).any()
    reps = tuple(data.draw(st.integers(1, 3)) for _ in range(data.draw(st.integers(1, 4))))
    _check(sp.tile(a, reps), np.tile(da, reps))
    ax = data.draw(st.one_of(st.none(), st.integers(0, len(sa) - 1)))
    r = data.draw(st.integers(1, 3))
    _check(sp.repeat(a, r, axis=ax), np.repeat(da, r, axis=ax))


@SET
@given(st.data())
def test_argreduce_sort_unique_match_numpy(data):
    sp = _sp()
    shape = data.draw(shapes)
    fill = data.draw(st.sampled_from([0.0, 2.0, -1.0]))
    a, d = _rand(shape, data.draw(st.integers(0, 99)), data.draw(st.sampled_from([0.0, 0.3, 0.8, 1.0])), fill)
    axis = data.draw(st.one_of(st.none(), st.integers(-len(shape), len(shape) - 1)))
    keepdims = data.draw(st.booleans())
    assert np.array_equal(sp.argmax(a, axis=axis, keepdims=keepdims).todense(), np.argmax(d, axis=axis, keepdims=keepdims))
    assert np.array_equal(sp.argmin(a, axis=axis, keepdims=keepdims).todense(), np.argmin(d, axis=axis, keepdims=keepdims))
    ax = data.draw(st.integers(-len(shape), len(shape) - 1))
    desc = data.draw(st.booleans())
    want = np.sort(d, axis=ax)
    _check(sp.sort(a, axis=ax, descending=desc), np.flip(want, axis=ax) if desc else want)
    vals, counts = np.unique(d, return_counts=True)
    got = sp.unique_counts(a)
    assert np.array_equal(got.values, vals) and np.array_equal(got.counts, counts)
    assert np.array_equal(sp.unique_values(a), vals)

"""Second batch of widening tests -- written after the one GPU run the last authoring session could afford, so they have
only run on the NumPy mock (whose contracts mirror the real kernel wrappers).  They live in a file that sorts after the
GPU-proven tests/test_widen_{manip,protocol,y_complex}.py: complex golden cases from the reference, narrow-integer
arithmetic, several advanced indices, np.where with dense operands, pickling."""
import numpy as np
import pytest

from _api import check_result, dec, sp  # noqa: F401
from _golden import load
from test_widen_manip import assert_eq, xy  # noqa: F401


def _dense(x):
    return x.todense() if hasattr(x, "todense") else x


@pytest.mark.parametrize("dt", [np.uint8, np.int8, np.int16, np.uint16, np.uint32])
def test_narrow_integer_arithmetic_wraps_like_numpy(sp, dt):
    """int8 ... uint32 compute in a wider signed type and are cast back: NumPy's modular arithmetic, result dtype kept,
    entries that wrap around to the fill value pruned."""
    rng = np.random.default_rng(7)
    info = np.iinfo(dt)
    a = rng.integers(info.min, int(info.max) + 1, size=(5, 6)).astype(dt)
    b = rng.integers(info.min, int(info.max) + 1, size=(5, 6)).astype(dt)
    a[rng.random((5, 6)) < 0.5] = 0
    b[rng.random((5, 6)) < 0.5] = 0
    x, y = sp.COO.from_numpy(a), sp.COO.from_numpy(b)
    with np.errstate(all="ignore"):
        for f in (np.add, np.subtract, np.multiply, np.maximum, np.minimum, np.bitwise_and, np.bitwise_or,
                  np.bitwise_xor, np.greater, np.equal, np.floor_divide, np.remainder):
            got, want = f(x, y), f(a, b)
            assert got.dtype == want.dtype and np.array_equal(got.todense(), want), f.__name__
            assert got.nnz == int(np.sum(want != got.fill_value)), f.__name__
        for f in (np.negative, np.abs, np.square, np.invert):
            got, want = f(x), f(a)
            assert got.dtype == want.dtype and np.array_equal(got.todense(), want), f.__name__
        assert np.array_equal((x + dt(3)).todense(), a + dt(3)) and (x * 2).dtype == (a * 2).dtype
        assert np.array_equal((x * 2).todense(), a * 2) and np.array_equal(_dense(x + b), a + b)
        assert np.array_equal((x.asformat("gcxs") - y.asformat("gcxs")).todense(), a - b)
        for axis in (None, 0, 1):  # reductions: NumPy's result dtype (uint64 for unsigned sums), modular like NumPy
            for name in ("sum", "max", "min", "prod", "any"):
                got, ref = getattr(x, name)(axis=axis), getattr(a, name)(axis=axis)
                assert got.dtype == np.asarray(ref).dtype and np.array_equal(got.todense(), ref), (name, axis)
        want = a @ b.T  # products: exact in the wide type, wrapped by the cast back
        for fa in ("coo", "gcxs", "dense"):
            for fb in ("coo", "gcxs", "dense"):
                if fa == fb == "dense":
                    continue
                got = (a if fa == "dense" else x.asformat(fa)) @ (b.T if fb == "dense" else y.T.asformat(fb))
                assert _dense(got).dtype == want.dtype and np.array_equal(_dense(got), want), (fa, fb)



MULTI = [([0, 1],) * 2, ([0, 1], [0, 2]), ([0, 1], [0, 2], [3, 1]), ([1, 1, 0], [2, 2, 0]), (slice(None), [0, 2], [1, 1]),
         (1, [0, 2], [1, 1]), ([1, 0], [2, 1], None), (Ellipsis, [0, 1], [3, 0]), ([0, 1], slice(None), [0, 3]),
         ([], [])]


@pytest.mark.parametrize("index", MULTI, ids=[str(i).replace(" ", "") for i in MULTI])
def test_several_advanced_indices(sp, xy, index):
    """tests/test_coo.py:456-457 upstream and more: 1-D advanced indices of one length are taken together."""
    x, _, d, _ = xy
    assert_eq(x[index], d[index])
    assert_eq(sp.GCXS(x)[index], d[index])



def test_where_with_dense_operands_and_array_properties(sp):
    """tests/test_array_function.py:59-80 upstream: np.where over every mix of dense / sparse operands; np.shape & co."""
    y = sp.random((7, 6), density=0.4, random_state=0)
    x = y.todense()
    want = np.where(x.astype(bool), x, x)
    for order in [(0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)]:
        a, b, c = [(x, y)[i] for i in order]
        got = np.where(a.astype(bool), b, c)
        assert np.array_equal(got.todense() if hasattr(got, "todense") else got, want), order
    assert np.shape(y) == (7, 6) and np.ndim(y) == 2 and np.size(y) == 42
    assert sp.asCOO(y) is y and sp.broadcast_shapes((3, 1), (1, 4)) == (3, 4)
    p, q = sp.broadcast_arrays(y[:1], np.ones((7, 1)))
    assert p.shape == q.shape == (7, 6) and np.array_equal(p.todense(), np.broadcast_to(x[:1], (7, 6)))
    with pytest.raises(ValueError):
        sp.asCOO(x)
    assert np.array_equal(sp.add(y, y).todense(), x + x) and sp.float32 is np.float32  # namespace re-exports


def test_pickle_round_trip(sp):
    """The pickled state is the host mirror (coordinates / indices and values); device arrays are rebuilt lazily."""
    import pickle
    import sys

    x = sp.random((5, 6, 3), density=0.4, random_state=0, fill_value=1.5)
    for a in (x, x.asformat("gcxs"), sp.CSR(x[0]), sp.DOK(x)):
        b = pickle.loads(pickle.dumps(a))
        assert type(b) is type(a) and b.fill_value == a.fill_value and np.array_equal(b.todense(), a.todense())
        if hasattr(a, "compressed_axes"):
            assert a.compressed_axes == b.compressed_axes
    assert np.array_equal((pickle.loads(pickle.dumps(x)) + x).todense(), 2 * x.todense())
    assert 400 < sys.getsizeof(sp.COO.from_numpy(np.eye(100))) < np.eye(100).nbytes / 10


# ---- against the REFERENCE itself: golden outputs generated by tests/golden/make_golden.py::gen_complex ---------------
GOLD = load("complex_api")


def _gid(i, c):
    return "-".join(str(v) for v in [i, c["op"], c.get("dt1", c.get("dtype")), c.get("dt2", ""), c.get("fa", ""),
                                     c.get("fb", ""), c.get("kind", ""), c.get("axis", "")] if v != "")


@pytest.mark.parametrize("c", GOLD, ids=[_gid(i, c) for i, c in enumerate(GOLD)])
def test_complex_golden(sp, c):
    """Result type, shape, fill value and COORDINATES exactly as the reference returns them; values to 2e-6 (complex64)
    / 1e-12 (complex128) -- the per-plane sums round differently from numba's complex multiply-accumulate."""
    op = c["op"]
    c64 = "complex64" in (c.get("dt1"), c.get("dt2"), c.get("dtype")) and "complex128" not in (c.get("dt1"), c.get("dt2"))
    f32 = c64 or "float32" in (c.get("dt1"), c.get("dt2"))
    tol = dict(rtol=5e-6, atol=2e-6) if f32 else dict(rtol=1e-12, atol=1e-13)
    a = dec(sp, c, "a_", c.get("fa", "coo"))
    if op == "matmul":
        got = sp.matmul(a, dec(sp, c, "b_", c["fb"]))
    elif op in ("add", "subtract", "multiply"):
        b = dec(sp, c, "b_")
        got = {"add": np.add, "subtract": np.subtract, "multiply": np.multiply}[op](a, b)
    elif op == "scale":
        got = a * (2 - 3j)
    elif op == "conj":
        got = a.conj()
    elif op == "real":
        got = a.real
    elif op == "imag":
        got = a.imag
    elif op == "abs":
        got = abs(a)
    else:
        axis = c["axis"]
        got = a.sum(axis=tuple(axis) if isinstance(axis, list) else axis)
    exact = op in ("conj", "real", "imag")
    check_result(sp, got, c, exact=exact, **({} if exact else tol))


def test_opt_in_result_cache_and_copy(sp):
    """tests/test_coo.py:754-820, 1231-1242 upstream: `cache=True` / `enable_caching()`, shallow vs deep copy."""
    x = sp.COO({(9, 9, 9): 1}, shape=(10, 10, 10))
    assert x.reshape((100, 10)).transpose().tocsr() is not x.reshape((100, 10)).transpose().tocsr()
    x = sp.COO({(9, 9, 9): 1}, shape=(10, 10, 10), cache=True)
    assert x[:].reshape((100, 10)).transpose().tocsr() is x[:].reshape((100, 10)).transpose().tocsr()
    y = sp.COO({(1, 1, 1, 1, 1, 1, 1, 2): 1}, shape=(2, 2, 2, 2, 2, 2, 2, 3), cache=True)
    for d in range(1, y.ndim):
        y.reshape((int(np.prod(y.shape[:d])), -1))
    assert len(y._cache["reshape"]) < 5
    s = sp.COO.from_numpy(np.array([1, 0, 0, 0, 0])).reshape((5, 1))
    s.enable_caching()
    assert s.T is s.T and s._cache is not None
    for deep in (True, False):
        c = s.copy(deep)
        assert c._cache is None and np.array_equal(c.todense(), s.todense())
        assert (c.data is s.data) is not deep and (c.coords is s.coords) is not deep

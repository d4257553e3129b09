"""sort / unique_values / unique_counts (sparse_b200/_sorting.py) vs NumPy.  Upstream: tests/test_coo.py:1822-1925
(`test_sort`, `test_unique_*`)."""
import numpy as np
import pytest

from _api import sp  # noqa: F401


def _rand(rng, dt):
    shape = tuple(rng.integers(1, 6, size=rng.integers(1, 4)))
    fill = float(rng.choice([0.0, 2.0, -1.5]))
    d = np.full(shape, fill).astype(dt)
    mask = rng.random(shape) < rng.choice([0.0, 0.3, 0.7, 1.0])
    vals = (rng.integers(-4, 5, size=shape) * (0.5 if np.dtype(dt).kind == "f" else 1)).astype(dt)
    d[mask] = vals[mask]
    return d, np.dtype(dt).type(fill)


@pytest.mark.parametrize("dt", [np.float64, np.float32, np.int64, np.int32])
@pytest.mark.parametrize("seed", range(6))
def test_sort_matches_numpy(sp, dt, seed):
    rng = np.random.default_rng(seed)
    d, fill = _rand(rng, dt)
    for x in (sp.COO.from_numpy(d, fill_value=fill), sp.GCXS.from_numpy(d, fill_value=fill)):
        for axis in range(-d.ndim, d.ndim):
            for descending in (False, True):
                got = sp.sort(x, axis=axis, descending=descending)
                want = np.sort(d, axis=axis)
                if descending:
                    want = np.flip(want, axis=axis)
                assert type(got) is type(x) and got.dtype == d.dtype and got.fill_value == fill
                assert np.array_equal(got.todense(), want)
                c = got.asformat("coo")
                if c.nnz:
                    assert np.all(np.diff(np.ravel_multi_index(tuple(c.coords), c.shape)) > 0)


@pytest.mark.parametrize("dt", [np.float64, np.float32, np.int64])
@pytest.mark.parametrize("seed", range(6))
def test_unique_matches_numpy(sp, dt, seed):
    rng = np.random.default_rng(100 + seed)
    d, fill = _rand(rng, dt)
    x = sp.COO.from_numpy(d, fill_value=fill)
    values, counts = np.unique(d, return_counts=True)
    assert np.array_equal(sp.unique_values(x), values)
    res = sp.unique_counts(x)
    assert np.array_equal(res.values, values) and np.array_equal(res.counts, counts)
    assert np.array_equal(np.unique_values(x), values)  # __array_function__


def test_nan_values_are_distinct_and_last(sp):
    d = np.array([1.0, np.nan, -2.0, np.nan, 0.0, 0.0, 5.0, -np.inf, np.inf])
    x = sp.COO.from_numpy(d)
    assert np.array_equal(sp.sort(x).todense(), np.sort(d), equal_nan=True)
    assert np.array_equal(sp.sort(x, descending=True).todense(), np.sort(d)[::-1], equal_nan=True)
    vals = sp.unique_values(x)
    assert np.array_equal(vals, np.unique(d, equal_nan=False), equal_nan=True) and np.isnan(vals[-2:]).all()
    res = sp.unique_counts(x)
    assert np.array_equal(res.counts, [1, 1, 2, 1, 1, 1, 1, 1])
    y = sp.COO.from_numpy(np.array([np.nan, 1.0, np.nan]), fill_value=np.nan)  # NaN fill: every fill is its own value
    assert np.isnan(sp.unique_values(y)[1:]).all() and len(sp.unique_values(y)) == 3


def test_sort_errors_and_bool(sp):
    x = sp.COO.from_numpy(np.array([[True, False, True], [False, False, True]]))
    assert np.array_equal(sp.sort(x, axis=1).todense(), np.sort(x.todense(), axis=1))
    with pytest.raises(IndexError):
        sp.sort(x, axis=2)
    with pytest.raises(ValueError):
        sp.unique_values(np.ones(3))
    with pytest.raises(ValueError):
        sp.unique_counts(np.ones(3))


@pytest.mark.parametrize("seed", range(8))
def test_interp_is_bit_identical_to_numpy(sp, seed):
    """tests/test_coo.py:2171-2213 upstream (`TestInterp`) and random tables: values, fill value and pruning."""
    rng = np.random.default_rng(seed)
    x = sp.random((6, 7, 3), density=float(rng.choice([0.0, 0.4, 1.0])), random_state=rng,
                  fill_value=float(rng.choice([0, 0.3, -2]))) * 4 - 1.5
    k = int(rng.integers(1, 6))
    xp = np.sort(rng.choice(np.arange(-4, 8) / 2.0, size=k, replace=False))
    fp = rng.standard_normal(k)
    for kw in ({}, {"left": -7.0}, {"right": 9.5}):
        for arr in (x, x.asformat("gcxs")):
            got = sp.interp(arr, xp, fp, **kw)
            want = np.interp(x.todense(), xp, fp, **kw)
            assert type(got) is type(arr) and got.fill_value == np.interp(x.fill_value, xp, fp, **kw)
            assert np.array_equal(got.todense().view(np.uint64), want.view(np.uint64))
            assert got.nnz == int(np.sum(want != got.fill_value))
    y = sp.random((10, 10, 10), random_state=seed)
    got = np.interp(y, sp.COO.from_numpy(np.array([-1, 0, 1])), [3, 2, 0])  # __array_function__, sparse table
    assert got.fill_value == 2 and np.array_equal(got.todense(), np.interp(y.todense(), [-1, 0, 1], [3, 2, 0]))
    with pytest.raises(TypeError):
        sp.interp(y, [-1, 0, 1], np.array([3, 2, 0]) + 1j)

"""NaN-skipping reductions vs golden outputs of the reference (tests/test_coo.py:196-263 upstream).
Coordinates / result types / fill values exact; values to 2e-6 (f32) / 1e-12 (f64) because NumPy's reduceat order is
unspecified (nanmax / nanmin are compared exactly)."""
import warnings

import numpy as np
import pytest

from _api import check_result, dec, sp  # noqa: F401
from _golden import load

CASES = load("nanreduce_api")


def _id(c):
    return f"{c['op']}-{c['dtype']}-{c['fmt']}-ax{c['axis']}-{'kd' if c['keepdims'] else ''}{c['note']}".replace(" ", "")


@pytest.mark.parametrize("c", CASES, ids=[f"{i}-{_id(c)}" for i, c in enumerate(CASES)])
def test_nanreduce(sp, c):
    x = dec(sp, c, "a_", c["fmt"], ca=c.arr.get("a_ca") if c["fmt"] == "gcxs" else None)
    axis = c["axis"]
    if isinstance(axis, list):
        axis = tuple(axis)
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = getattr(sp, c["op"])(x, axis=axis, keepdims=c["keepdims"])
    exact = c["op"] in ("nanmax", "nanmin") or c["dtype"] == "int64" and c["op"] != "nanmean"
    tol = 2e-6 if c["dtype"] == "float32" else 1e-12
    check_result(sp, got, c, exact=exact, rtol=tol, atol=tol)


def test_all_nan_slice_warns(sp):
    d = np.array([[1.0, 0.0, 2.0], [np.nan, np.nan, np.nan]])
    x = sp.COO.from_numpy(d)
    with pytest.warns(RuntimeWarning, match="All-NaN slice"):
        sp.nanmax(x, axis=1)
    with pytest.warns(RuntimeWarning, match="Mean of empty slice"), np.errstate(all="ignore"):
        sp.nanmean(x, axis=1)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert sp.nanmax(x, axis=0).todense().tolist() == [1.0, 0.0, 2.0]


def test_nan_functions_reject_dense(sp):
    with pytest.raises(ValueError):
        sp.nansum(np.ones((2, 2)))

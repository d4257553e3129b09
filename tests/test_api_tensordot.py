"""tensordot / matmul / dot through the public API vs golden outputs of the reference
(shapes of tests/test_dot.py:15-80, 114-163, 289-300 upstream).  Bit-exact: every kernel on this path keeps the
reference's summation order."""
import numpy as np
import pytest

from _api import check_result, dec, sp  # noqa: F401
from _golden import load

CASES = load("tensordot_api")
RT = {"none": None, "coo": "COO", "gcxs": "GCXS", "dense": np.ndarray}


def _id(c):
    return f"{c['op']}-{c.get('dtype','f8')}-{c['fa']}x{c['fb']}-{c.get('rt','')}-{c.get('axes','')}".replace(" ", "")


@pytest.mark.parametrize("c", [c for c in CASES if c["op"] == "tensordot"], ids=_id)
def test_tensordot(sp, c):
    a = dec(sp, c, "a_", c["fa"])
    b = dec(sp, c, "b_", c["fb"])
    rt = RT[c["rt"]]
    if isinstance(rt, str):
        rt = getattr(sp, rt)
    axes = c["axes"]
    if isinstance(axes, list):
        axes = tuple(tuple(x) if isinstance(x, list) else x for x in axes)
    if "error" in c:
        with pytest.raises(Exception):
            sp.tensordot(a, b, axes, return_type=rt)
        return
    got = sp.tensordot(a, b, axes, return_type=rt)
    check_result(sp, got, c)


@pytest.mark.parametrize("c", [c for c in CASES if c["op"] in ("matmul", "dot")], ids=_id)
def test_matmul_dot(sp, c):
    a = dec(sp, c, "a_", c["fa"])
    b = dec(sp, c, "b_", c["fb"])
    f = sp.matmul if c["op"] == "matmul" else sp.dot
    got = f(a, b)
    check_result(sp, got, c)


def test_errors(sp):
    x = sp.COO(np.array([[0, 1], [1, 2]]), np.array([1.0, 2.0]), shape=(3, 4), has_duplicates=False, sorted=True)
    with pytest.raises(ValueError, match="shape-mismatch"):
        sp.tensordot(x, np.ones((3, 2)), axes=1)
    y = sp.COO(np.array([[0], [1]]), np.array([1.0]), shape=(4, 2), has_duplicates=False, sorted=True, fill_value=1.0)
    with pytest.raises(ValueError, match="zero fill"):
        sp.tensordot(x, y, axes=1)
    with pytest.raises(TypeError):
        sp.matmul(x, 3)
    z = sp.COO(np.array([[0], [1]]), np.array([1.5], dtype=np.float16), shape=(4, 2), has_duplicates=False,
               sorted=True)  # float16 is storage-only: no arithmetic, no silent upcast
    with pytest.raises(TypeError, match="dtype"):
        sp.tensordot(x, z, axes=1)

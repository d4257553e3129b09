"""einsum vs golden outputs of the reference (tests/test_einsum.py upstream: the same subscripts list, operand
formats, dtype= and the interleaved call form).  Result type / shape / coordinates exact; values to 1e-12 (f64) --
the products are exact-order but the duplicate sums follow NumPy's reduceat order in the reference."""
import numpy as np
import pytest

from _api import check_result, dec, sp  # noqa: F401
from _golden import load

CASES = load("einsum_api")


def _id(i, c):
    return f"{i}-{c.get('sub', c.get('lists'))}-{c['note']}".replace(" ", "")


@pytest.mark.parametrize("c", CASES, ids=[_id(i, c) for i, c in enumerate(CASES)])
def test_einsum(sp, c):
    ops = [dec(sp, c, f"op{i}_", fmt) for i, fmt in enumerate(c["fmts"])]
    if c["op"] == "einsum_lists":
        lists = [[Ellipsis if s == -1 else s for s in li] for li in c["lists"]]
        got = sp.einsum(ops[0], *lists)
    elif c["dtype"]:
        got = sp.einsum(c["sub"], *ops, dtype=np.dtype(c["dtype"]))
    else:
        got = sp.einsum(c["sub"], *ops)
    tol = 1e-6 if c["dtype"] == "float32" else 1e-12
    check_result(sp, got, c, exact=False, rtol=tol, atol=tol)


def test_einsum_through_numpy_protocol(sp):
    rng = np.random.default_rng(5)
    a = sp.random((4, 5), density=0.5, random_state=rng)
    b = sp.random((5, 3), density=0.5, random_state=rng)
    got = np.einsum("ij,jk->ik", a, b)
    assert isinstance(got, sp.COO)
    assert np.allclose(got.todense(), a.todense() @ b.todense(), rtol=1e-12, atol=0)


def test_einsum_fill_value_and_no_input(sp):
    x = sp.random((2,), density=0.5, fill_value=2.0, random_state=1)
    with pytest.raises(ValueError):
        sp.einsum("cba", x)
    with pytest.raises(ValueError):
        sp.einsum()


@pytest.mark.parametrize("subscript", ["a+b->c", "i->&", "i->ij", "ij->jij", "a..,a...", ".i...", "a,a->->"])
def test_einsum_invalid_input(sp, subscript):
    x = sp.random((2,), density=0.5, random_state=2)
    y = sp.random((2,), density=0.5, random_state=3)
    with pytest.raises(ValueError):
        sp.einsum(subscript, x, y)


@pytest.mark.parametrize("subscript", [0, [0, 0]])
def test_einsum_type_error(sp, subscript):
    x = sp.random((2,), density=0.5, random_state=2)
    y = sp.random((2,), density=0.5, random_state=3)
    with pytest.raises(TypeError):
        sp.einsum(subscript, x, y)


def test_einsum_shape_check(sp):
    x = sp.random((2, 3, 4), density=0.5, random_state=4)
    y = sp.random((2, 3, 4), density=0.5, random_state=5)
    with pytest.raises(ValueError):
        sp.einsum("aab", x)
    with pytest.raises(ValueError):
        sp.einsum("abc,acb", x, y)

"""SDDMM and MTTKRP: fused kernels and the unfused example expressions vs golden outputs of the reference
(examples/sddmm_example.py:43-55, examples/mttkrp_example.py:43-55).  The reference value depends on BLAS /
reduceat association, so values are compared to 2e-5 (f32) / 1e-11 (f64); coordinates exactly."""
import numpy as np
import pytest

from _api import check_result, dec, sp  # noqa: F401
from _golden import load

CASES = load("examples_api")


@pytest.mark.parametrize("c", [c for c in CASES if c["op"] == "sddmm"], ids=lambda c: c["dtype"])
@pytest.mark.parametrize("fused", [True, False], ids=["fused", "expr"])
def test_sddmm(sp, c, fused):
    s = dec(sp, c, "s_")
    a, b = np.array(c.arr["a"]), np.array(c.arr["b"])
    got = sp.sddmm(s, a, b) if fused else s * (a @ b)
    tol = 2e-5 if c["dtype"] == "float32" else 1e-11
    check_result(sp, got, c, exact=False, rtol=tol, atol=tol)


@pytest.mark.parametrize("c", [c for c in CASES if c["op"] == "mttkrp"], ids=lambda c: c["dtype"])
@pytest.mark.parametrize("fused", [True, False], ids=["fused", "expr"])
def test_mttkrp(sp, c, fused):
    B = dec(sp, c, "B_", "gcxs", ca=c.arr["B_ca"])
    Dm, Cm = np.array(c.arr["D"]), np.array(c.arr["C"])
    if fused:
        got = sp.mttkrp(B, Dm, Cm)
    else:
        got = sp.sum(B[:, :, :, None] * Dm[None, None, :, :] * Cm[None, :, None, :], axis=(1, 2))
    tol = 2e-5 if c["dtype"] == "float32" else 1e-11
    w = c.sub("out_")
    assert isinstance(got, sp.GCXS)
    assert tuple(got.shape) == tuple(int(s) for s in w["shape"])
    ref = sp.GCXS((w["data"], w["indices"], w["indptr"]), shape=tuple(w["shape"]), compressed_axes=tuple(w["ca"]))
    assert np.allclose(got.todense(), ref.todense(), rtol=tol, atol=tol)
    assert got.nnz == len(w["data"])

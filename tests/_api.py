"""Shared helpers for the API-level parity tests (golden fixtures from the reference vs sparse_b200).

Every test body runs twice: with the NumPy mock of the kernel layer on a box without a GPU (host-logic
check, `-m "not gpu"`) and against the real CUDA kernels on the B200 box (`-m gpu`)."""
import numpy as np
import pytest

BACKENDS = [pytest.param("mock", id="mock"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def sp(request):
    """The sparse_b200 module wired to the requested backend."""
    import torch

    import _mock_kernels
    import sparse_b200

    if request.param == "mock":
        if torch.cuda.is_available():
            pytest.skip("mock backend is only used on boxes without a GPU")
        _mock_kernels.install()
        yield sparse_b200
        _mock_kernels.uninstall()
    else:
        _mock_kernels.uninstall()
        from sparse_b200 import _lib

        _lib.load()
        yield sparse_b200


def dec(sp, case, prefix, fmt="coo", ca=None):
    """Build a sparse_b200 operand from the fixture's canonical COO encoding."""
    a = case.sub(prefix)
    kind = str(a["kind"])
    if kind == "dense":
        return np.array(a["array"])
    assert kind == "coo", kind
    x = sp.COO(np.array(a["coords"]), np.array(a["data"]), shape=tuple(int(s) for s in a["shape"]),
               has_duplicates=False, sorted=True, fill_value=a["fill"][()])
    if fmt == "coo":
        return x
    if fmt == "gcxs":
        return x.asformat("gcxs", **({"compressed_axes": tuple(ca)} if ca is not None else {}))
    if fmt == "dense":
        return x.todense()
    raise AssertionError(fmt)


def same_bits(x, y):
    x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
    if x.shape != y.shape or x.dtype != y.dtype:
        return False
    return np.array_equal(x.view(np.uint8), y.view(np.uint8))


def check_values(got, want, exact=True, rtol=0.0, atol=0.0, what="data"):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert got.dtype == want.dtype, (what, got.dtype, want.dtype)
    if exact or got.dtype.kind not in "fc":
        if got.dtype.kind in "fc":
            assert same_bits(got, want), f"{what}: not bit-identical (max abs diff {np.nanmax(np.abs(got - want)) if got.size else 0})"
        else:
            assert np.array_equal(got, want), what
    else:
        assert np.allclose(got, want, rtol=rtol, atol=atol, equal_nan=True), (
            what, float(np.nanmax(np.abs(got - want))) if got.size else 0.0)


def check_result(sp, got, case, prefix="out_", exact=True, rtol=0.0, atol=0.0):
    """Compare a result with the reference's (kind, shape, fill value, coordinates exact, data exact or tol)."""
    w = case.sub(prefix)
    kind = str(w["kind"])
    if kind == "dense":
        assert isinstance(got, np.ndarray) or np.isscalar(got), type(got)
        check_values(np.asarray(got), w["array"], exact, rtol, atol, "dense")
        return
    shape = tuple(int(s) for s in w["shape"])
    assert tuple(got.shape) == shape, (got.shape, shape)
    fill_w = w["fill"][()]
    assert got.fill_value.dtype == fill_w.dtype, (got.fill_value.dtype, fill_w.dtype)
    if exact or fill_w.dtype.kind not in "fc":
        assert same_bits(np.asarray(got.fill_value), np.asarray(fill_w)), (got.fill_value, fill_w)
    else:
        assert np.allclose(got.fill_value, fill_w, rtol=rtol, atol=atol, equal_nan=True)
    if kind == "coo":
        assert isinstance(got, sp.COO), type(got)
        assert np.array_equal(got.coords, w["coords"]), "coords differ"
        check_values(got.data, w["data"], exact, rtol, atol)
    else:
        assert isinstance(got, sp.GCXS), type(got)
        ca = tuple(int(c) for c in w["ca"]) if len(w["ca"]) else None
        assert got.compressed_axes == ca, (got.compressed_axes, ca)
        assert np.array_equal(got.indices, w["indices"]), "indices differ"
        if len(w["indptr"]):
            assert np.array_equal(got.indptr, w["indptr"]), "indptr differ"
        check_values(got.data, w["data"], exact, rtol, atol)

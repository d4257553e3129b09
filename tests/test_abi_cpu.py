"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol
include/sparse_b200.h declares, and fails loudly without a device (no CPU fallback)."""
import ctypes
import subprocess

import pytest

from sparse_b200 import _lib


def test_library_loads_and_exports_header_symbols():
    lib = _lib.load()
    names = _lib.header_symbols()
    assert "b2s_spmm_csr_dense" in names and len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sparse_b200.h but not exported"
    assert lib.b2s_abi_version() == 1


def test_exports_match_header_exactly():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T b2s_" in l)
    assert exported == _lib.header_symbols()


def test_no_device_is_a_loud_error():
    import torch

    if torch.cuda.is_available():
        pytest.skip("device present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.require_device()
    lib = _lib.load()
    n = ctypes.c_int(-1)
    assert lib.b2s_device_count(ctypes.byref(n)) != 0 and n.value == 0
    assert lib.b2s_last_error()  # message set


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (test infrastructure)."""
    import os
    import re

    pkg = os.path.dirname(_lib.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f

"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol
include/sparse_b200.h declares, and fails loudly without a device (no CPU fallback)."""
import ctypes
import subprocess

import numpy as np
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


def test_host_side_index_narrowing_is_thread_safe():
    """b2s_host_narrow_i64_i32 (the staging step of the host-buffer product) needs no GPU: several Python threads
    calling it at once (ctypes releases the GIL) share one thread pool behind a mutex and all get exact results."""
    import threading

    from sparse_b200 import _kernels as Kn

    rng = np.random.default_rng(0)
    srcs = [rng.integers(0, 2**31 - 1, 1_000_003 + 17 * i, dtype=np.int64) for i in range(4)]
    dsts = [np.zeros(s.size, np.int32) for s in srcs]

    def work(i):
        for _ in range(3):
            Kn.host_narrow(srcs[i], dsts[i])

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    for s, d in zip(srcs, dsts):
        assert np.array_equal(d, s.astype(np.int32))
    small = np.arange(5, dtype=np.int64)
    assert np.array_equal(Kn.host_narrow(small, np.empty(5, np.int32)), small.astype(np.int32))


def test_host_staging_mode_follows_the_local_world_size():
    """One rank per GPU with 4+ ranks on a box makes host DRAM the bottleneck of the host-buffer product: the host-side
    int64 -> int32 narrowing pass is switched off there (raw upload + device narrowing); needs no GPU."""
    from sparse_b200 import _dist as SD

    assert SD.configure_host_staging(8) == "device" and SD.configure_host_staging(4) == "device"
    assert SD.configure_host_staging(2) == "host" and SD.configure_host_staging(1) == "host"


def test_numa_binding_reports_why_it_could_not_bind():
    from sparse_b200 import _dist as SD

    got = SD.bind_to_gpu_numa(0, 0, 1)
    assert got is None or "numa_node" in got
    if got and got["numa_node"] is None:
        assert got["note"]

"""Seeded plain-NumPy input recipes for the FULL-SIZE parity fixtures (shared by make_fullsize_digests.py, which feeds
them to the reference, and by the -m gpu tests / bench.py's `configs` block, which feed them to sparse_b200)."""
import hashlib

import numpy as np

C3_SHAPE_A, C3_SHAPE_B = (512, 512, 512, 64), (512, 512, 512, 1)


def coo_inputs(shape, n_draws, seed, dtype=np.float64):
    """Canonical COO entries: sorted unique C-order linear indices (about n_draws of them) and uniform[0,1) data."""
    rng = np.random.default_rng(seed)
    size = int(np.prod(shape, dtype=np.int64))
    lin = np.unique(rng.integers(0, size, size=n_draws, dtype=np.int64))
    coords = np.stack(np.unravel_index(lin, shape)).astype(np.int64)
    data = rng.random(lin.shape[0]).astype(dtype)
    return coords, data


def c3_inputs(dtype=np.float64):
    """BASELINE.json config 3: (512,512,512,64) and (512,512,512,1) at density 1e-4."""
    a = coo_inputs(C3_SHAPE_A, 858_993, 0, dtype)
    b = coo_inputs(C3_SHAPE_B, 13_421, 1, dtype)
    return a, b


def digest(*arrays):
    h = hashlib.sha256()
    for x in arrays:
        x = np.ascontiguousarray(x)
        h.update(str(x.dtype).encode() + str(x.shape).encode())
        h.update(x.tobytes())
    return h.hexdigest()


def coo_digest(coords, data):
    """Digest of a canonical COO result: coordinates as int64 [ndim, nnz], values by bit pattern."""
    return digest(np.asarray(coords, dtype=np.int64), np.asarray(data))

"""`random`: synthetic sparse arrays (host-side input generator; same signature as sparse.random,
sparse/numba_backend/_utils.py:221-346).  Positions are sampled without replacement on the host with NumPy;
this is input generation, not part of the accelerated path."""
from __future__ import annotations

import numpy as np

from ._coo import COO
from ._utils import can_store, prod


def random(shape, density=None, nnz=None, random_state=None, data_rvs=None, format="coo", fill_value=None,
           idx_dtype=None, **kwargs):
    if isinstance(shape, (int, np.integer)):
        shape = (shape,)
    shape = tuple(int(s) for s in shape)
    if density is not None and nnz is not None:
        raise ValueError("Specify at most one of `density` or `nnz`")
    if density is None:
        density = 0.01
    if not (0 <= density <= 1):
        raise ValueError(f"density {density} is not in the unit interval")
    elements = prod(shape)
    if idx_dtype is not None and shape and not can_store(idx_dtype, max(shape)):
        raise ValueError(f"cannot cast array with shape {shape} to dtype {idx_dtype}.")
    if nnz is None:
        nnz = int(elements * density)
    if not (0 <= nnz <= elements):
        raise ValueError(f"cannot generate {nnz} nonzero elements for an array with {elements} total elements")
    rng = random_state if isinstance(random_state, np.random.Generator) else np.random.default_rng(random_state)
    if data_rvs is None:
        data_rvs = rng.random
    if elements and nnz > elements // 2 and elements < 2**27:
        ind = np.sort(rng.choice(elements, size=nnz, replace=False))
    else:
        ind = np.unique(rng.integers(0, max(elements, 1), size=int(nnz * 1.05) + 16, dtype=np.int64))
        while len(ind) < nnz:
            extra = rng.integers(0, elements, size=nnz - len(ind) + 16, dtype=np.int64)
            ind = np.unique(np.concatenate([ind, extra]))
        if len(ind) > nnz:
            ind = np.sort(rng.choice(ind, size=nnz, replace=False))
    data = np.asarray(data_rvs(nnz))
    coords = np.stack(np.unravel_index(ind, shape)).astype(idx_dtype or np.intp) if shape else np.empty((0, nnz), np.intp)
    ar = COO(coords, data, shape=shape, has_duplicates=False, sorted=True, fill_value=fill_value)
    return ar.asformat(format, **kwargs)

"""save_npz / load_npz: the reference's on-disk format (sparse/numba_backend/_io.py:7-132), node for node, so files
written by either package load in the other.  Pure host I/O around one D2H / H2D of the component arrays; a loaded
COO is taken as canonical (sorted=True, has_duplicates=False), exactly like the reference does."""
from __future__ import annotations

import numpy as np

from ._coo import COO
from ._gcxs import GCXS


def save_npz(filename, matrix, compressed=True):
    """_io.py:7-66: nodes data / shape / fill_value + coords (COO) or indices / indptr / compressed_axes (GCXS)."""
    nodes = {"data": matrix.data, "shape": matrix.shape, "fill_value": matrix.fill_value}
    if isinstance(matrix, COO):
        nodes["coords"] = matrix.coords
    elif isinstance(matrix, GCXS):
        nodes["indices"] = matrix.indices
        nodes["indptr"] = matrix.indptr
        nodes["compressed_axes"] = matrix.compressed_axes
    else:
        raise TypeError(f"sparse_b200.save_npz: unsupported array type {type(matrix).__name__}")
    (np.savez_compressed if compressed else np.savez)(filename, **nodes)


def load_npz(filename):
    """_io.py:69-132."""
    with np.load(filename) as fp:
        names = set(fp.files)
        if {"coords", "data", "shape", "fill_value"} <= names:
            return COO(fp["coords"], fp["data"], shape=tuple(int(s) for s in fp["shape"]), sorted=True,
                       has_duplicates=False, fill_value=fp["fill_value"][()])
        if {"data", "indices", "indptr", "compressed_axes", "shape", "fill_value"} <= names:
            ca = fp["compressed_axes"]
            ca = None if ca.ndim == 0 and ca[()] is None else tuple(int(c) for c in np.atleast_1d(ca))
            return GCXS((fp["data"], fp["indices"], fp["indptr"]), shape=tuple(int(s) for s in fp["shape"]),
                        fill_value=fp["fill_value"][()], compressed_axes=ca)
    raise RuntimeError(f"The file {filename!s} does not contain a valid sparse matrix")

"""einsum: any combination of traces, transposes, sums and broadcast products of sparse operands.

Host-side mirror of sparse/numba_backend/_common.py:1150-1476 (`_parse_einsum_input`, `_einsum_single`, `einsum`).
The algebra is the reference's: every operand is first reduced / aligned to the ordered union of the indices that
matter (`_einsum_single`), the aligned operands are multiplied with broadcasting (the fused merge kernel of
csrc/elemwise.cu), and one last single-term einsum produces the output.

A single-term einsum of a COO array is ONE linearisation here: the trace selector (repeated labels) is a flag kernel
over the coordinate rows (`b2s_coo_diag_flags`), the transpose and the dropped (summed) axes are folded into the
strides of `b2s_coo_linearize` (stride 0 for dropped rows and for the later occurrences of a repeated label), and the
duplicate keys that result are summed by the canonicalisation primitives (sort, head flags, segmented sum) -- the
reference builds permuted coordinates and lets `COO(..., has_duplicates=True)` do the same on the host.
"""
from __future__ import annotations

import operator

import numpy as np

from . import _device as D
from . import _kernels as Kn
from ._coo import COO
from ._elemwise import elemwise
from ._sparse_array import SparseArray
from ._utils import c_strides, check_zero_fill_value, key_bits, prod

# label alphabet shared with NumPy: position in this string <-> integer label of the interleaved call form
_LABELS = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


def _implicit_output(terms_joined):
    """Labels that occur exactly once, in sorted order (classical Einstein convention)."""
    out = []
    for ch in sorted(set(terms_joined)):
        if not ch.isalpha():
            raise ValueError(f"Character {ch} is not a valid symbol.")
        if terms_joined.count(ch) == 1:
            out.append(ch)
    return "".join(out)


def _sublist_to_term(sub):
    if not isinstance(sub, (list, tuple)):
        raise TypeError("For this input type lists must contain either int or Ellipsis")
    term = ""
    for s in sub:
        if s is Ellipsis:
            term += "..."
            continue
        try:
            term += _LABELS[operator.index(s)]
        except TypeError as e:
            raise TypeError("For this input type lists must contain either int or Ellipsis") from e
    return term


def _parse_einsum_input(operands):
    """(input_subscripts, output_subscript, operands) with ellipses expanded and the output made explicit
    (_common.py:1150-1321; same accepted forms and errors as numpy.einsum's own front end)."""
    if len(operands) == 0:
        raise ValueError("No input operands")
    if isinstance(operands[0], str):
        spec = operands[0].replace(" ", "")
        operands = list(operands[1:])
        for ch in spec:
            if ch not in ".,->" and not ch.isalpha():
                raise ValueError(f"Character {ch} is not a valid symbol.")
    else:
        rest = list(operands)
        arrays, terms = [], []
        while len(rest) >= 2:
            arrays.append(rest.pop(0))
            terms.append(_sublist_to_term(rest.pop(0)))
        spec = ",".join(terms)
        if rest:
            spec += "->" + _sublist_to_term(rest[0])
        operands = arrays

    if "-" in spec or ">" in spec:
        if spec.count("-") > 1 or spec.count(">") > 1 or spec.count("->") != 1:
            raise ValueError("Subscripts can only contain one '->'.")

    if "->" in spec:
        lhs, rhs = spec.split("->")
    else:
        lhs, rhs = spec, None
    terms = lhs.split(",")

    if "." in spec:
        used = set(spec) - set(".,->")
        spare = [ch for ch in _LABELS if ch not in used]  # labels standing for the ellipsis axes
        longest = 0
        for k, term in enumerate(terms):
            if "." not in term:
                continue
            if term.count(".") != 3 or term.count("...") != 1:
                raise ValueError("Invalid Ellipses.")
            if k >= len(operands):
                raise ValueError("Number of einsum subscripts must be equal to the number of operands.")
            nd = len(np.shape(operands[k]) if not hasattr(operands[k], "shape") else operands[k].shape)
            count = 0 if nd == 0 else max(nd, 1) - (len(term) - 3)
            if count < 0:
                raise ValueError("Ellipses lengths do not match.")
            longest = max(longest, count)
            terms[k] = term.replace("...", "".join(spare[len(spare) - count:]) if count else "")
        ell = "".join(spare[len(spare) - longest:]) if longest else ""
        if rhs is not None:
            rhs = rhs.replace("...", ell)
        else:
            plain = _implicit_output("".join(terms))
            rhs = ell + "".join(sorted(set(plain) - set(ell)))
    elif rhs is None:
        rhs = _implicit_output("".join(terms))

    lhs = ",".join(terms)
    for ch in rhs:
        if not ch.isalpha():
            raise ValueError(f"Character {ch} is not a valid symbol.")
        if ch not in lhs:
            raise ValueError(f"Output character {ch} did not appear in the input")
    if len(terms) != len(operands):
        raise ValueError("Number of einsum subscripts must be equal to the number of operands.")
    return lhs, rhs, list(operands)


def _segment_sum(data, heads, pos, total):
    """Kn.segment_sum for the compute dtypes, per plane for complex values (sparse_b200/_complex.py)."""
    if data.is_complex():
        from ._complex import segment_sum

        return segment_sum(data, heads, pos, total)
    return Kn.segment_sum(data, heads, pos, total)


def _sum_all(data):
    """0-D COO holding the sum of a device vector (full contraction)."""
    n = int(data.shape[0])
    if n == 0:
        return COO.from_numpy(np.add.reduce(np.empty(0, dtype=D.np_dtype(data))))
    return COO._from_device(None, data, (n,), None, keys=Kn.iota(n)).sum()


def _einsum_single(lhs, rhs, operand):
    """Single-term einsum: traces, transposes and sums of one array (_common.py:1324-1407)."""
    from ._gcxs import GCXS

    if lhs == rhs:
        if not rhs:
            return operand.sum()  # 0-D result per the Array API
        return operand
    if not isinstance(operand, SparseArray):
        # dense operand: stays dense (the reference calls np.einsum here).  It takes the same device path as a sparse
        # one -- its non-zero entries as a COO through the trace selector / re-keying / reduction kernels -- and is
        # densified again; no library einsum on the way
        from ._coo import as_coo

        if D.is_device_tensor(operand):  # the product of two dense operands computed on the device
            if operand.ndim == 0:
                operand = D.download(operand)
            else:
                dt = D.np_dtype(operand)
                operand = COO._from_dense_device(operand.contiguous(), tuple(operand.shape), dt.type(0))
        res = _einsum_single(lhs, rhs, as_coo(operand) if not isinstance(operand, SparseArray) else operand)
        return res.todense_device() if isinstance(res, SparseArray) else res
    was_gcxs = isinstance(operand, GCXS)
    operand = operand.tocoo() if was_gcxs else operand

    first = [lhs.index(ch) for ch in lhs]  # first axis carrying each axis' label
    for d, f in enumerate(first):
        if operand.shape[d] != operand.shape[f]:
            raise ValueError("Repeated indices must have the same dimension.")
    new_shape = tuple(operand.shape[lhs.index(ch)] for ch in rhs)

    coords, data = operand._dev()
    if any(f != d for d, f in enumerate(first)) and operand.nnz:
        flags = Kn.diag_flags(coords, first)
        pos, total = Kn.scan_flags(flags)
        if total != operand.nnz:
            coords = Kn.compact_rows(coords, flags, pos, total)
            data = Kn.compact(data, flags, pos, total)
    if not rhs:
        return _sum_all(data)

    st_new = c_strides(new_shape)
    strides = [0] * len(lhs)
    for p, ch in enumerate(rhs):
        strides[lhs.index(ch)] = st_new[p]
    nnz = int(data.shape[0])
    if nnz == 0:
        out = COO(np.zeros((len(new_shape), 0), dtype=np.intp), np.empty(0, dtype=operand.dtype), shape=new_shape,
                  has_duplicates=False, sorted=True)
    else:
        keys = Kn.linearize(coords, strides)
        unsorted, dups = Kn.keys_flags(keys)
        if unsorted:
            keys, perm = Kn.sort_keys(keys, key_bits(prod(new_shape)))
            data = Kn.gather(data, perm)
            _, dups = Kn.keys_flags(keys)
        if dups:  # axes summed away / traced: COO(..., has_duplicates=True) of the reference
            heads = Kn.flag_heads(keys)
            pos, total = Kn.scan_flags(heads)
            data = _segment_sum(data, heads, pos, total)
            keys = Kn.compact(keys, heads, pos, total)
        out = COO._from_device(None, data, new_shape, None, keys=keys)
    return GCXS.from_coo(out) if was_gcxs else out


def einsum(*operands, **kwargs):
    """numpy.einsum for sparse operands (_common.py:1410-1476)."""
    lhs, rhs, operands = _parse_einsum_input(operands)
    check_zero_fill_value(*operands)
    dtype = kwargs.pop("dtype", None)
    if kwargs:
        raise TypeError(f"sparse_b200.einsum: unsupported keyword arguments {sorted(kwargs)}")
    if not any(isinstance(o, SparseArray) for o in operands):
        if all(isinstance(o, np.ndarray) for o in operands):
            return np.einsum(f"{lhs}->{rhs}", *operands, **({"dtype": dtype} if dtype is not None else {}))
        raise ValueError(f"None of the args is sparse: {operands}")
    from ._dok import DOK

    sparse_ops = [o for o in operands if isinstance(o, SparseArray)]
    if any(isinstance(o, DOK) for o in sparse_ops):  # the DOK builder computes as COO; all-DOK input gives DOK back
        all_dok = all(isinstance(o, DOK) for o in sparse_ops)
        out = einsum(f"{lhs}->{rhs}", *[o.to_coo() if isinstance(o, DOK) else o for o in operands],
                     **({"dtype": dtype} if dtype is not None else {}))
        return DOK.from_coo(out.asformat("coo")) if all_dok and isinstance(out, SparseArray) else out
    if dtype is not None:
        operands = [o.astype(dtype) if hasattr(o, "astype") else o.to(D.torch_dtype(dtype)) for o in operands]
    if len(operands) == 1:
        return _einsum_single(lhs, rhs, operands[0])

    terms = lhs.split(",")
    seen_in, sizes = {}, {}
    for t, term in enumerate(terms):
        shape = tuple(operands[t].shape)
        for ch, extent in zip(term, shape, strict=False):
            if extent != sizes.setdefault(ch, extent):
                raise ValueError(f"Inconsistent shape for index '{ch}'.")
            seen_in.setdefault(ch, set()).add(t)
    for ch in rhs:
        seen_in[ch].add(-1)
    # labels living on a single term and absent from the output are summed away before the product
    aligned = "".join(ch for ch, where in seen_in.items() if len(where) > 1)

    parts = []
    for term, array in zip(terms, operands, strict=True):
        pterm = "".join(ch for ch in aligned if ch in term)
        if pterm != term:
            array = _einsum_single(term, pterm, array)
        shape = tuple(array.shape[pterm.index(ch)] if ch in pterm else 1 for ch in aligned)
        parts.append(array.reshape(shape) if tuple(array.shape) != shape else array)
    out = _einsum_single(aligned, rhs, _product(parts))
    if D.is_device_tensor(out) and not any(D.is_device_tensor(o) for o in operands):
        # every sparse operand collapsed to a scalar factor: the result is dense, and host operands get a host array
        # back (the reference returns an ndarray here)
        out = D.download(out)
    return out


def _product(parts):
    """Left-to-right broadcasting product (reduce(mul, ...)); a dense prefix is multiplied on the device."""
    acc = parts[0]
    for nxt in parts[1:]:
        if not isinstance(acc, SparseArray) and not isinstance(nxt, SparseArray):
            from ._elemwise import dense_binary

            acc = dense_binary(np.multiply, acc if D.is_device_tensor(acc) else D.upload(np.ascontiguousarray(acc)),
                               nxt if D.is_device_tensor(nxt) else D.upload(np.ascontiguousarray(nxt)))
        else:
            acc = elemwise(np.multiply, acc, nxt)
    return acc

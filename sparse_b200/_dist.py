"""Multi-GPU execution of the hot path: one process per GPU, 1-D row blocks of the left operand.

The path shards naturally along the first (row) axis of the left operand: output rows are independent, so the
only exchange is making the right operand visible to every rank:

* ``tensordot_rowblock``   C_r = A_r @ B:  A_r is this rank's nnz-balanced row block, B arrives row-sharded and is
  all-gathered over NCCL/NVLink (``dist.all_gather_into_tensor``) before the local K1 kernel.
* ``spgemm_rowblock``      C_r = A_r @ B for sparse B: B's CSR arrays are all-gathered (variable sizes, padded).
* ``sddmm_rowblock``       mask row block x local rows of `a`; `b` (K x N) arrives column-sharded and is gathered.

No reduction collective is needed for row blocking (a reduce-scatter would only appear if the CONTRACTION axis
were split, which doubles the dense traffic).  Host logic is backend-agnostic: the `gloo` tests
run it with world_size 2 on CPU tensors.
"""
from __future__ import annotations

import numpy as np

from . import _device as D
from . import _kernels as Kn


def nnz_balanced_splits(indptr, world: int):
    """Row boundaries r_0=0 <= r_1 <= ... <= r_world=M such that every block holds ~nnz/world stored entries
    (split points are chosen on indptr, not on the row count)."""
    ip = np.asarray(indptr, dtype=np.int64)
    M = len(ip) - 1
    nnz = int(ip[-1]) if M >= 0 and len(ip) else 0
    bounds = [0]
    for r in range(1, world):
        target = nnz * r // world
        cut = int(np.searchsorted(ip, target, side="left"))
        cut = min(max(cut, bounds[-1]), M)
        bounds.append(cut)
    bounds.append(M)
    return bounds


def row_block(data, indices, indptr, r0: int, r1: int):
    """Rows [r0, r1) of a CSR triple (host arrays); indptr rebased to 0."""
    ip = np.asarray(indptr)
    lo, hi = int(ip[r0]), int(ip[r1])
    return data[lo:hi], indices[lo:hi], (ip[r0:r1 + 1] - ip[r0])


def _dist():
    import torch.distributed as dist

    return dist


def all_gather_rows(shard, group=None):
    """Concatenate equally sized row shards of a dense operand along axis 0 (one all-gather)."""
    dist = _dist()
    t = D.torch()
    world = dist.get_world_size(group)
    if world == 1:
        return shard
    shard = shard.contiguous()
    out = t.empty((shard.shape[0] * world,) + tuple(shard.shape[1:]), dtype=shard.dtype, device=shard.device)
    dist.all_gather_into_tensor(out, shard, group=group)
    return out


def all_gather_varlen(x, group=None):
    """All-gather 1-D tensors of different lengths: returns (list of per-rank tensors)."""
    dist = _dist()
    t = D.torch()
    world = dist.get_world_size(group)
    if world == 1:
        return [x]
    n = t.tensor([x.shape[0]], dtype=t.int64, device=x.device)
    sizes = [t.zeros(1, dtype=t.int64, device=x.device) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes + [1])
    pad = t.zeros(mx, dtype=x.dtype, device=x.device)
    pad[: x.shape[0]] = x
    bufs = [t.empty(mx, dtype=x.dtype, device=x.device) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return [b[:s] for b, s in zip(bufs, sizes)]


def tensordot_rowblock(a_local, b_shard, group=None, out=None):
    """Local block of ``A @ B`` (dense result rows owned by this rank).

    a_local : 2-D GCXS (compressed_axes=(0,)) holding this rank's row block of A (all K columns).
    b_shard : dense device tensor, this rank's K/world rows of B (equal shards).
    """
    from ._dot import _csr_arrays, _dot_dtype

    B = all_gather_rows(b_shard, group)
    M, K = a_local.shape
    assert B.shape[0] == K, (B.shape, K)
    dtr = _dot_dtype(a_local.dtype, D.np_dtype(B))
    ad, ai, ap = _csr_arrays(a_local, dtr)
    return Kn.spmm_csr_dense(ad, ai, ap, Kn.cast(B, dtr), M, K, int(B.shape[1]), out=out)


def gather_csr_rows(b_local, group=None):
    """All-gather a row-sharded CSR operand (each rank owns a block of consecutive rows)."""
    from ._gcxs import GCXS

    t = D.torch()
    data, indices, indptr = b_local._dev()
    datas = all_gather_varlen(data, group)
    idxs = all_gather_varlen(indices, group)
    ptrs = all_gather_varlen(indptr, group)
    out_ptr, base = [], 0
    for r, p in enumerate(ptrs):
        p = p.to(t.int64)
        out_ptr.append((p if r == 0 else p[1:]) + base)
        base += int(p[-1].item())
    full_ptr = t.cat(out_ptr)
    rows = int(full_ptr.shape[0]) - 1
    return GCXS((t.cat(datas), t.cat(idxs).to(t.int64), full_ptr), shape=(rows, b_local.shape[1]),
                compressed_axes=(0,))


def spgemm_rowblock(a_local, b_local, group=None):
    """Local row block of ``A @ B`` for sparse operands: B's row blocks are all-gathered, then K4 runs locally."""
    from ._dot import _dot

    B = gather_csr_rows(b_local, group)
    assert a_local.shape[1] == B.shape[0]
    return _dot(a_local, B)


def sddmm_rowblock(s_local, a_local, b_cols_shard, group=None):
    """Local row block of ``s * (a @ b)``: `b` arrives column-sharded (K x N/world) and is gathered as b^T rows."""
    from ._fused import sddmm

    bt_shard = Kn.transpose_dense(b_cols_shard)  # (N/world, K)
    Bt = all_gather_rows(bt_shard, group)        # (N, K)
    return sddmm(s_local, a_local, Bt, b_transposed=True)

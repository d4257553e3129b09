"""N>1 host logic on CPU: world_size 2 and 3 `gloo` runs of the row-block partition + all-gather paths
(sparse_b200/_dist.py) with the NumPy mock of the kernel layer; results must equal the single-process product."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import _mock_kernels

        _mock_kernels.install()
        import oracle
        import sparse_b200 as sp
        from _util import rand_csr, rand_dense
        from sparse_b200 import _dist as DD

        rng = np.random.default_rng(123)  # same inputs on every rank
        M, K, N = 60, 48, 16
        data, indices, indptr = rand_csr(rng, M, K, 0.2, np.float32)
        B = rand_dense(rng, (K, N), np.float32)
        bounds = DD.nnz_balanced_splits(indptr, world)
        r0, r1 = bounds[rank], bounds[rank + 1]
        d, i, p = DD.row_block(data, indices, indptr, r0, r1)
        a_local = sp.GCXS((d, i, p), shape=(r1 - r0, K), compressed_axes=(0,))
        shard = K // world
        b_shard = torch.from_numpy(B[rank * shard:(rank + 1) * shard].copy())
        c_local = DD.tensordot_rowblock(a_local, b_shard).numpy()
        want = oracle.dot_csr_ndarray((M, N), data, indices, indptr, B)
        ok1 = np.array_equal(c_local.view(np.uint32), want[r0:r1].view(np.uint32))

        # sparse x sparse: B row-sharded as CSR blocks
        bd, bi, bp = rand_csr(rng, K, 40, 0.2, np.float64)
        ad = data.astype(np.float64)
        bb = DD.nnz_balanced_splits(bp, world)
        s0, s1 = bb[rank], bb[rank + 1]
        b_local = sp.GCXS(DD.row_block(bd, bi, bp, s0, s1), shape=(s1 - s0, 40), compressed_axes=(0,))
        a_local64 = sp.GCXS((d.astype(np.float64), i, p), shape=(r1 - r0, K), compressed_axes=(0,))
        g = DD.spgemm_rowblock(a_local64, b_local)
        wd, wi, wp = oracle.dot_csr_csr((r1 - r0, 40), d.astype(np.float64), bd, i, bi, p, bp)
        ok2 = (np.array_equal(g.indptr, wp) and np.array_equal(g.indices, wi)
               and np.array_equal(g.data.view(np.uint64), wd.view(np.uint64)))

        # sddmm row block with b column-sharded
        NS = 36  # divisible by every world size the tests use (equal column shards of b)
        S = sp.random((M, NS), density=0.2, random_state=np.random.default_rng(5)).astype(np.float64)
        A = rand_dense(rng, (M, 8), np.float64)
        Bm = rand_dense(rng, (8, NS), np.float64)
        full = (S.todense() * (A @ Bm))
        sd, si, sptr = (S.asformat("gcxs", compressed_axes=(0,)).data, S.asformat("gcxs", compressed_axes=(0,)).indices,
                        S.asformat("gcxs", compressed_axes=(0,)).indptr)
        sb = DD.nnz_balanced_splits(sptr, world)
        t0, t1 = sb[rank], sb[rank + 1]
        s_local = sp.GCXS(DD.row_block(sd, si, sptr, t0, t1), shape=(t1 - t0, NS), compressed_axes=(0,))
        cs = NS // world
        out = DD.sddmm_rowblock(s_local, torch.from_numpy(A[t0:t1].copy()),
                                torch.from_numpy(np.ascontiguousarray(Bm[:, rank * cs:(rank + 1) * cs])))
        ok3 = np.allclose(out.todense(), full[t0:t1], rtol=1e-12, atol=1e-12)

        # contraction split: A[:, K_r] @ B[K_r, :], all-to-all of the partials' row blocks, local merge
        kb = [K * r // world for r in range(world + 1)]
        a_full = sp.GCXS((ad, indices, indptr), shape=(M, K), compressed_axes=(0,))
        b_full = sp.GCXS((bd, bi, bp), shape=(K, 40), compressed_axes=(0,))
        a_cols = a_full[:, kb[rank]:kb[rank + 1]]
        b_rows = b_full[kb[rank]:kb[rank + 1], :]
        a_cols = a_cols if isinstance(a_cols, sp.GCXS) else a_cols.asformat("gcxs", compressed_axes=(0,))
        b_rows = b_rows if isinstance(b_rows, sp.GCXS) else b_rows.asformat("gcxs", compressed_axes=(0,))
        blk = DD.spgemm_ksplit(a_cols, b_rows)
        m0, m1 = M * rank // world, M * (rank + 1) // world
        want_k = (a_full.todense() @ b_full.todense())[m0:m1]
        ok4 = blk.shape == (m1 - m0, 40) and np.allclose(blk.todense(), want_k, rtol=1e-12, atol=1e-13)
        q.put((rank, bool(ok1), bool(ok2), bool(ok3) and bool(ok4), bounds))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_rowblock_paths(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 2000 * (world - 2)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok1, ok2, ok3, bounds in res:
        assert ok1, f"rank {rank}: tensordot_rowblock differs from the single-process product"
        assert ok2, f"rank {rank}: spgemm_rowblock differs"
        assert ok3, f"rank {rank}: sddmm_rowblock / spgemm_ksplit differs"
        assert bounds[0] == 0 and bounds[-1] == 60 and bounds == sorted(bounds)


def test_nnz_balanced_splits():
    from sparse_b200._dist import nnz_balanced_splits

    indptr = np.array([0, 0, 10, 10, 11, 50, 100])
    b = nnz_balanced_splits(indptr, 4)
    assert b[0] == 0 and b[-1] == 6 and b == sorted(b)
    per = [indptr[b[i + 1]] - indptr[b[i]] for i in range(4)]
    assert sum(per) == 100 and max(per) <= 50
    assert nnz_balanced_splits(np.array([0]), 2) == [0, 0, 0]

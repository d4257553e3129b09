#!/usr/bin/env python
"""Multi-GPU timings of the two BASELINE.json configs that name a GPU count: C4 (SDDMM, 1 -> 8 GPUs) and C5 (CSR SpGEMM
row-partitioned across the GPUs of one box).  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_multi.py [c4] [c5] [--scale 1.0] [--steps 5]

Partitioning (sparse_b200/_dist.py): the mask / left operand is cut into consecutive row blocks (C5:
nnz-balanced on indptr), the right operand arrives sharded and is all-gathered (C4: b^T rows, 1 GB at full size; C5:
the CSR arrays of A, variable length) -- STRONG scaling: the total problem size is fixed, as the configs state.  Every
rank generates only its own block (8 seeded row panels for both configs, so the global problem is the same for every
N); every timing comes with a correctness check of the local row block (C5: bit-exact against the oracle on its first
rows; C4: sampled entries against an fp64 restatement) and an order-independent bit-pattern checksum over all ranks
that must agree between N = 1 and N = 8.  Timing: CUDA events on
the current stream around `steps` calls after 2 warm-ups, barrier on both sides, MAX over ranks; rank 0 prints one
JSON line per config and writes gpurun_out/multi_<N>.json.  `--cpu-smoke` runs the same code on the NumPy mock of the
kernel layer over gloo (tiny sizes; checks the plumbing and the result against the single-process product, no timing
value) -- that is what tests/test_dist_multi_tool.py launches with world_size 2.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("which", nargs="*", default=["c5", "c5k", "c3", "c4"])
    p.add_argument("--scale", type=float, default=1.0, help="shrinks M (and nnz with it) for quick runs")
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--cpu-smoke", action="store_true")
    return p.parse_args()


def block_csr(rng, rows, K, per_row):
    """Uniform-random CSR block (sorted unique columns per row) as host arrays, int64 indices."""
    lin = np.unique(rng.integers(0, rows * K, size=int(rows * per_row), dtype=np.int64))
    r, c = lin // K, lin % K
    indptr = np.zeros(rows + 1, dtype=np.int64)
    np.cumsum(np.bincount(r, minlength=rows), out=indptr[1:])
    return c, indptr


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    import torch.distributed as dist

    if args.cpu_smoke:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import _mock_kernels

        _mock_kernels.install()
        dev = torch.device("cpu")
        if world > 1:
            dist.init_process_group("gloo")
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench_multi.py: no CUDA device (use --cpu-smoke for the plumbing check)")
        torch.cuda.set_device(local_rank)
        dev = torch.device("cuda", local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=dev)
    import sparse_b200 as sp
    from sparse_b200 import _device as D
    from sparse_b200 import _dist as SD
    from sparse_b200 import _kernels as Kn
    from sparse_b200 import _lib

    if not args.cpu_smoke:
        _lib.load()

    def sync():
        if not args.cpu_smoke:
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn):
        """ms per call: device events (max over ranks); wall clock in the CPU smoke mode."""
        for _ in range(2):
            out = fn()
        sync()
        if args.cpu_smoke:
            t0 = time.perf_counter()
            for _ in range(args.steps):
                out = fn()
            ms = (time.perf_counter() - t0) * 1e3 / args.steps
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                out = fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        sync()
        return ms, out

    def total(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    results = {}
    tiny = args.cpu_smoke

    def bits_checksum(*tensors):
        """Order-independent exact checksum: sum of the raw 32-/64-bit patterns as int64 (mod 2^64), over all ranks.
        Equal checksums at N = 1 and N = 8 mean the row blocks together are the single-GPU result, bit for bit."""
        tot = 0
        for x in tensors:
            x = x.contiguous()
            v = x.view(torch.int32) if x.element_size() == 4 else x.view(torch.int64)
            tot += int(v.to(torch.int64).sum().item())
        tot &= (1 << 62) - 1
        if world > 1:
            t = torch.tensor([tot], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            tot = int(t.item()) & ((1 << 62) - 1)
        return tot

    def all_true(flag):
        if world == 1:
            return bool(flag)
        t = torch.tensor([1 if flag else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    blocks = max(world, 8)  # both global problems are 8 seeded row panels, whatever N is
    mine = [b for b in range(blocks) if b * world // blocks == rank]

    # ---- C5: (M x M @ density) squared, row blocks of A; the right operand (= A) arrives row-sharded --------------------
    if "c5" in args.which:
        M = 240 if tiny else int(1_000_000 * args.scale)
        per_row = 4 if tiny else 10
        rows_per = M // blocks
        M = rows_per * blocks
        cols, vals, ptrs = [], [], [np.zeros(1, dtype=np.int64)]
        for b in mine:
            rng = np.random.default_rng(1000 + b)
            c, ip = block_csr(rng, rows_per, M, per_row)
            cols.append(c)
            vals.append(rng.random(len(c), dtype=np.float32))
            ptrs.append(ip[1:] + ptrs[-1][-1])
        indices, data, indptr = np.concatenate(cols), np.concatenate(vals), np.concatenate(ptrs)
        a_local = sp.GCXS((data, indices, indptr), shape=(rows_per * len(mine), M), compressed_axes=(0,))
        a_local.to_device()
        ms, out = timed((lambda: SD.spgemm_rowblock(a_local, a_local)) if world > 1
                        else (lambda: sp.tensordot(a_local, a_local, axes=1)))
        nnz_out = total(out.nnz)
        results["C5"] = {"config": f"CSR({M}x{M}, {per_row}/row)^2 f32, {world} row blocks, CSR all-gather of the right operand",
                         "n_gpus": world, "ms_per_step": round(ms, 4), "out_nnz": int(nnz_out),
                         "Gnnz_out_s": round(nnz_out / ms / 1e6, 4), "scaling": "strong"}
        if not tiny:
            od, oi, _op = out._dev()
            results["C5"]["checksum_data_indices"] = [bits_checksum(od), bits_checksum(oi.to(torch.int64))]
            # the local row block against the oracle: its first 4000 rows x the whole gathered matrix
            import oracle

            full = SD.gather_csr_rows(a_local) if world > 1 else a_local
            fd, fi, fp = full.data, full.indices.astype(np.int64), full.indptr.astype(np.int64)
            rows = min(4000, a_local.shape[0])
            lp = a_local.indptr.astype(np.int64)
            n = int(lp[rows])
            d, i, _p = oracle.dot_csr_csr((rows, M), a_local.data[:n], fd, a_local.indices[:n].astype(np.int64), fi,
                                          lp[: rows + 1], fp)
            keep = d.view(np.uint32) != 0
            m = int(out.indptr[rows])
            ok = m == int(keep.sum()) and np.array_equal(out.indices[:m], i[keep]) and np.array_equal(
                out.data[:m].view(np.uint32), d[keep].view(np.uint32))
            results["C5"]["row_block_bit_exact_vs_oracle_first_rows"] = all_true(ok)
        if tiny:  # the sharded product equals the single-process one
            full = SD.gather_csr_rows(a_local) if world > 1 else a_local
            want = (full @ full).todense()
            lo = sum(rows_per for b in range(blocks) if b * world // blocks < rank)
            assert np.allclose(out.todense(), want[lo:lo + out.shape[0]]), "C5 row block differs"
        # the CONTRACTION-split form of the same product (BASELINE config 5's "reduce-scatter of output rows"): rank r
        # holds A[:, K_r] and B[K_r, :] = its own row panels; partial products over all rows, all-to-all of the row
        # blocks, local merge.  Measured next to the row-blocked form, not instead of it.
        if world > 1 and "c5k" in args.which:
            full = SD.gather_csr_rows(a_local)
            lo = sum(rows_per for b in range(blocks) if b * world // blocks < rank)
            a_cols = full[:, lo:lo + a_local.shape[0]]
            a_cols = a_cols if isinstance(a_cols, sp.GCXS) else a_cols.asformat("gcxs", compressed_axes=(0,))
            if a_cols.compressed_axes != (0,):
                a_cols = a_cols.change_compressed_axes((0,))
            a_cols.to_device()
            del full
            ms_k, blk = timed(lambda: SD.spgemm_ksplit(a_cols, a_local))
            oc = out.tocoo()
            same = bool(torch.equal(blk.sorted_keys(), oc.sorted_keys())) and bool(
                torch.allclose(blk._data_dev(), oc._data_dev(), rtol=1e-5, atol=0))
            results["C5 contraction-split"] = {
                "config": f"same product, A column blocks x B row blocks, all-to-all of the partials' row blocks + merge",
                "n_gpus": world, "ms_per_step": round(ms_k, 4), "Gnnz_out_s": round(nnz_out / ms_k / 1e6, 4),
                "scaling": "strong", "matches_row_blocked_result_rtol_1e-5": all_true(same)}
            del a_cols, blk, oc
        del a_local, out

    # ---- C4: mask row blocks x local rows of a; b arrives column-sharded and is gathered as b^T ------------------------
    if "c4" in args.which:
        M = N = 192 if tiny else int(1_000_000 * args.scale)
        K = 8 if tiny else 256
        per_row = 6 if tiny else 100
        M = N = (M // (8 * blocks)) * 8 * blocks
        rows_per = M // blocks
        cols, vals, ptrs, a_rows, b_parts = [], [], [np.zeros(1, dtype=np.int64)], [], []
        for b in mine:  # mask rows, rows of a and columns of b of panel b: the same global problem for every N
            rng = np.random.default_rng(3000 + b)
            c, ip = block_csr(rng, rows_per, N, per_row)
            cols.append(c)
            vals.append(rng.random(len(c), dtype=np.float32))
            ptrs.append(ip[1:] + ptrs[-1][-1])
            a_rows.append(rng.random((rows_per, K), dtype=np.float32))
            b_parts.append(np.random.default_rng(4000 + b).random((K, N // blocks), dtype=np.float32))
        c, v, ip = np.concatenate(cols), np.concatenate(vals), np.concatenate(ptrs)
        rows = rows_per * len(mine)
        s_local = sp.GCXS((v, c, ip), shape=(rows, N), compressed_axes=(0,)).tocoo()
        a_local = D.upload(np.concatenate(a_rows))
        b_cols = D.upload(np.ascontiguousarray(np.concatenate(b_parts, axis=1)))
        del cols, vals, ptrs, a_rows, b_parts, c, v, ip
        ms, out = timed((lambda: SD.sddmm_rowblock(s_local, a_local, b_cols)) if world > 1
                        else (lambda: sp.sddmm(s_local, a_local, b_cols)))
        nnz = total(s_local.nnz)
        alg = nnz * 12 + M * K * 4 + nnz * K * 4 + nnz * 4
        results["C4"] = {"config": f"SDDMM mask {M}x{N} ({per_row}/row), K={K} f32, {world} mask row blocks, b^T all-gather",
                         "n_gpus": world, "ms_per_step": round(ms, 4), "mask_nnz": int(nnz),
                         "Gnnz_s": round(nnz / ms / 1e6, 4), "alg_GBs_all_gpus": round(alg / ms / 1e6, 1),
                         "scaling": "strong"}
        if not tiny:
            oc, od = out._dev()
            results["C4"]["checksum_data"] = bits_checksum(od)
            Bt = SD.all_gather_rows(Kn.transpose_dense(b_cols)) if world > 1 else Kn.transpose_dense(b_cols)
            g = torch.Generator(device=dev).manual_seed(5 + rank)
            sel = torch.randint(0, int(od.shape[0]), (4096,), generator=g, device=dev)
            sc_, sd_ = s_local._dev()
            live = sd_ != 0  # float32 uniform[0,1) draws hit exactly 0.0 a few times in 1e8: those products are pruned
            same_coords = bool(torch.equal(oc, sc_[:, live]))
            sel = torch.randint(0, int(od.shape[0]), (4096,), generator=g, device=dev)
            dots = (a_local[oc[0, sel]].double() * Bt[oc[1, sel]].double()).sum(dim=1)
            want = sd_[live][sel].double() * dots
            ok = same_coords and bool(torch.allclose(od[sel].double(), want, rtol=2e-5, atol=0))
            results["C4"]["row_block_matches_fp64_restatement_on_samples"] = all_true(ok)
            del Bt
        if tiny:
            Bt = SD.all_gather_rows(D.torch().as_tensor(np.ascontiguousarray(D.download(b_cols).T)).to(dev)) \
                if world > 1 else D.torch().as_tensor(np.ascontiguousarray(D.download(b_cols).T))
            want = s_local.todense() * (D.download(a_local) @ D.download(Bt).T)
            assert np.allclose(out.todense(), want, rtol=1e-4), "C4 row block differs"

    # ---- C3-large: element-wise add and reductions, range-partitioned on the leading axis -----------
    if "c3" in args.which and not tiny:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs

        bench_configs.DEV = dev  # its generators default to cuda:0
        rand_coo = bench_configs.rand_coo  # seeded device generator: the same tensors on every rank

        shape_a, shape_b = (512, 512, 512, 64), (512, 512, 512, 1)
        a = rand_coo(shape_a, int(85_899_345 * args.scale), 10)
        b = rand_coo(shape_b, int(1_342_177 * args.scale), 11)
        lead = shape_a[0]
        r0, r1 = lead * rank // world, lead * (rank + 1) // world  # uniform data: equal leading ranges are balanced
        a_blk, b_blk = SD.leading_block(a, r0, r1), SD.leading_block(b, r0, r1)
        n_in = total(a_blk.nnz + b_blk.nnz)
        ms, out = timed(lambda: SD.elemwise_leading(np.add, a_blk, b_blk))
        full = (a + b)[r0:r1]  # this rank's slice of the single-GPU result
        ok = bool(torch.equal(full.sorted_keys(), out.sorted_keys())) and bool(
            torch.equal(full._data_dev().view(torch.int64), out._data_dev().view(torch.int64)))
        n_out = total(out.nnz)
        alg = n_in * 16 + n_out * 40
        results["C3-large add"] = {"config": f"COO {shape_a} + {shape_b} (broadcast on axis 3), density 1e-2, f64; "
                                             f"{world} leading-axis ranges, no collective",
                                   "n_gpus": world, "ms_per_step": round(ms, 4), "nnz_in": int(n_in),
                                   "nnz_out": int(n_out), "Gnnz_in_s": round(n_in / ms / 1e6, 4),
                                   "alg_GBs_all_gpus": round(alg / ms / 1e6, 1), "scaling": "strong",
                                   "block_bit_exact_vs_slice_of_single_gpu_result": all_true(ok)}
        del full, out
        for name, axis in (("sum axis=3 (axis 0 kept: local)", (3,)),
                           ("sum axis=(0,1) (axis 0 reduced: all-gather of sparse partials + one more reduction)", (0, 1))):
            ms, red = timed(lambda: SD.reduce_leading(a_blk, np.add, axis=axis))
            ref = a.sum(axis=axis)
            if 0 in axis:  # the same (replicated) result on every rank
                ok = bool(torch.equal(ref.sorted_keys(), red.sorted_keys())) and bool(
                    torch.allclose(ref._data_dev(), red._data_dev(), rtol=1e-12, atol=0))
            else:
                refb = ref[r0:r1]
                ok = bool(torch.equal(refb.sorted_keys(), red.sorted_keys())) and bool(
                    torch.allclose(refb._data_dev(), red._data_dev(), rtol=1e-12, atol=0))
            results["reduce-large " + name] = {"n_gpus": world, "ms_per_step": round(ms, 4),
                                               "Gnnz_s": round(total(a_blk.nnz) / ms / 1e6, 4), "scaling": "strong",
                                               "matches_single_gpu_result": all_true(ok)}
        del a, b, a_blk, b_blk

    if rank == 0:
        for k, v in results.items():
            print(json.dumps({k: v}), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        name = f"multi_cpu_smoke_{world}.json" if args.cpu_smoke else f"multi_{world}.json"
        with open(os.path.join(ROOT, "gpurun_out", name), "w") as f:
            json.dump(results, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

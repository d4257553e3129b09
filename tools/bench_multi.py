#!/usr/bin/env python
"""Multi-GPU timings of the two BASELINE.json configs that name a GPU count: C4 (SDDMM, 1 -> 8 GPUs) and C5 (CSR SpGEMM
row-partitioned across the GPUs of one box).  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_multi.py [c4] [c5] [--scale 1.0] [--steps 5]

Partitioning (sparse_b200/_dist.py): the mask / left operand is cut into consecutive row blocks (C5:
nnz-balanced on indptr), the right operand arrives sharded and is all-gathered (C4: b^T rows, 1 GB at full size; C5:
the CSR arrays of A, variable length) -- STRONG scaling: the total problem size is fixed, as the configs state.  Every
rank generates only its own block (C5: 8 seeded row panels, so the matrix is the same for every N; C4: seeded per rank,
same size and distribution for every N).  Timing: CUDA events on
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
    p.add_argument("which", nargs="*", default=["c4", "c5"])
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

    # ---- C5: (M x M @ density) squared, row blocks of A; the right operand (= A) arrives row-sharded --------------------
    if "c5" in args.which:
        M = 240 if tiny else int(1_000_000 * args.scale)
        per_row = 4 if tiny else 10
        blocks = max(world, 8)  # the global matrix is 8 seeded row panels, whatever N is
        rows_per = M // blocks
        M = rows_per * blocks
        mine = [b for b in range(blocks) if b * world // blocks == rank]
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
        if tiny:  # the sharded product equals the single-process one
            full = SD.gather_csr_rows(a_local) if world > 1 else a_local
            want = (full @ full).todense()
            lo = sum(rows_per for b in range(blocks) if b * world // blocks < rank)
            assert np.allclose(out.todense(), want[lo:lo + out.shape[0]]), "C5 row block differs"

    # ---- C4: mask row blocks x local rows of a; b arrives column-sharded and is gathered as b^T ------------------------
    if "c4" in args.which:
        M = N = 192 if tiny else int(1_000_000 * args.scale)
        K = 8 if tiny else 256
        per_row = 6 if tiny else 100
        M = N = (M // (8 * world)) * 8 * world
        rows = M // world
        rng = np.random.default_rng(3000 + rank)
        c, ip = block_csr(rng, rows, N, per_row)
        vals = rng.random(len(c), dtype=np.float32)
        s_local = sp.GCXS((vals, c, ip), shape=(rows, N), compressed_axes=(0,)).tocoo()
        a_local = D.upload(rng.random((rows, K), dtype=np.float32))
        b_cols = D.upload(np.random.default_rng(4000 + rank).random((K, N // world), dtype=np.float32))
        ms, out = timed((lambda: SD.sddmm_rowblock(s_local, a_local, b_cols)) if world > 1
                        else (lambda: sp.sddmm(s_local, a_local, b_cols)))
        nnz = total(s_local.nnz)
        alg = nnz * 12 + M * K * 4 + nnz * K * 4 + nnz * 4
        results["C4"] = {"config": f"SDDMM mask {M}x{N} ({per_row}/row), K={K} f32, {world} mask row blocks, b^T all-gather",
                         "n_gpus": world, "ms_per_step": round(ms, 4), "mask_nnz": int(nnz),
                         "Gnnz_s": round(nnz / ms / 1e6, 4), "alg_GBs_all_gpus": round(alg / ms / 1e6, 1),
                         "scaling": "strong"}
        if tiny:
            Bt = SD.all_gather_rows(D.torch().as_tensor(np.ascontiguousarray(D.download(b_cols).T)).to(dev)) \
                if world > 1 else D.torch().as_tensor(np.ascontiguousarray(D.download(b_cols).T))
            want = s_local.todense() * (D.download(a_local) @ D.download(Bt).T)
            assert np.allclose(out.todense(), want, rtol=1e-4), "C4 row block differs"

    if rank == 0:
        for k, v in results.items():
            print(json.dumps({k: v}), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"multi_{world}.json"), "w") as f:
            json.dump(results, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

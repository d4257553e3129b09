#!/usr/bin/env python
"""Sweep K1 variants on the C2 workload (device-resident, CUDA-event timed). Writes gpurun_out/tune_k1.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sparse_b200 import _kernels as Kn  # noqa: E402


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device("cuda", 0)
    M = K = int(os.environ.get("ROWS", 1_000_000))
    nnz = int(os.environ.get("NNZ", 100_000_000))
    res = []
    vals, cols, indptr, B = bench.make_workload(torch, M, K, nnz, 256, 1234, dev)
    peak, _ = bench.peaks()
    for ncols in (128, 64, 32, 256):
        Bn = B[:, :ncols].contiguous()
        C = torch.empty((M, ncols), dtype=torch.float32, device=dev)
        variants = [(1, 4), (1, 8), (1, 16), (1, 32)] + ([(2, 8)] if ncols == 128 else [])
        if ncols != 128:
            variants = [(1, 8)]
        for v, u in variants:
            Kn.spmm_set_variant(v, u)
            ms = timeit(lambda: Kn.spmm_csr_dense(vals, cols, indptr, Bn, M, K, ncols, out=C))
            alg = bench.algorithmic_bytes(int(vals.numel()), M, ncols)
            r = {"ncols": ncols, "variant": v, "unroll": u, "ms": round(ms, 4), "gnnz_s": round(vals.numel() / ms / 1e6, 3),
                 "alg_GBs": round(alg / ms / 1e6, 1), "frac": round(alg / ms / 1e6 / peak, 4)}
            print(r, flush=True)
            res.append(r)
    Kn.spmm_set_variant(1, 8)
    # power-law row lengths (same nnz budget): exercises load imbalance of the row-split
    g = torch.Generator(device=dev).manual_seed(7)
    u = torch.rand(M, generator=g, device=dev, dtype=torch.float64)
    lens = (u.pow(-1 / 1.5) * 20).clamp(max=50_000).to(torch.int64)
    lens = (lens.double() * (nnz / lens.sum().item())).to(torch.int64).clamp(min=0, max=K)
    indptr2 = torch.zeros(M + 1, dtype=torch.int64, device=dev)
    indptr2[1:] = torch.cumsum(lens, 0)
    n2 = int(indptr2[-1].item())
    cols2 = torch.randint(0, K, (n2,), generator=g, device=dev, dtype=torch.int32)
    vals2 = torch.rand(n2, generator=g, device=dev, dtype=torch.float32)
    Bn = B[:, :128].contiguous()
    C = torch.empty((M, 128), dtype=torch.float32, device=dev)
    ip32 = indptr2.to(torch.int32)
    for v, uu in [(1, 8), (2, 8)]:
        Kn.spmm_set_variant(v, uu)
        ms = timeit(lambda: Kn.spmm_csr_dense(vals2, cols2, ip32, Bn, M, K, 128, out=C))
        alg = bench.algorithmic_bytes(n2, M, 128)
        r = {"ncols": 128, "variant": v, "unroll": uu, "rows": "power-law(1.5), max %d" % int(lens.max().item()), "nnz": n2,
             "ms": round(ms, 4), "gnnz_s": round(n2 / ms / 1e6, 3), "alg_GBs": round(alg / ms / 1e6, 1),
             "frac": round(alg / ms / 1e6 / peak, 4)}
        print(r, flush=True)
        res.append(r)
    Kn.spmm_set_variant(1, 8)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "tune_k1.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

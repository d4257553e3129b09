#!/usr/bin/env python
"""Time the non-headline BASELINE.json configs (C1, C3, C4, C5, reductions, MTTKRP) through the public API.

Device-resident inputs (generated with seeded torch generators on the GPU), wall-clock around
torch.cuda.synchronize() because the public calls contain their own size-returning syncs.
Writes gpurun_out/configs.json; a per-kernel launch list comes from running this under
`ncu --metrics gpu__time_duration.sum` (see profiles/).

    python tools/bench_configs.py [c1 c3 c3big red c4 c5 mttkrp]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import sparse_b200 as sp  # noqa: E402
from sparse_b200 import _lib  # noqa: E402

DEV = torch.device("cuda", 0)
PEAK, _ = bench.peaks()


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), float(np.median(ts))


def rand_coo(shape, nnz, seed, dtype=torch.float64):
    g = torch.Generator(device=DEV).manual_seed(seed)
    size = int(np.prod(shape))
    lin = torch.unique(torch.randint(0, size, (int(nnz * 1.02) + 64,), generator=g, device=DEV, dtype=torch.int64))
    if lin.numel() > nnz:
        lin = lin[torch.randperm(lin.numel(), generator=g, device=DEV)[:nnz]].sort().values
    coords = torch.stack(torch.unravel_index(lin, shape))
    data = torch.rand(lin.numel(), generator=g, device=DEV, dtype=dtype)
    return sp.COO(coords, data, shape=shape, has_duplicates=False, sorted=True)


def rand_csr(M, K, nnz, seed, dtype=torch.float32):
    vals, cols, indptr, _ = bench.make_workload(torch, M, K, nnz, 1, seed, DEV)
    return sp.GCXS((vals.to(dtype), cols, indptr), shape=(M, K), compressed_axes=(0,))


def main():
    which = sys.argv[1:] or ["c1", "c3", "red", "c5", "c4", "mttkrp", "c3big"]
    res = {}
    _lib.load()

    def rec(name, fn, units, unit_name, alg_bytes=None, reps=5):
        n0 = _lib.launch_count()
        fn()
        launches = _lib.launch_count() - n0
        best, med = timeit(fn, reps=reps)
        r = {"ms_best": round(best, 4), "ms_median": round(med, 4), unit_name: round(units / best / 1e6, 4),
             "launches_per_call": int(launches)}
        if alg_bytes:
            r["alg_GBs"] = round(alg_bytes / best / 1e6, 1)
            r["frac_of_hbm_peak"] = round(alg_bytes / best / 1e6 / PEAK, 4)
        res[name] = r
        print(name, r, flush=True)

    if "c1" in which:
        rng = np.random.default_rng(42)
        a = sp.random((1000, 1000), density=0.01, random_state=rng)
        b = sp.random((1000, 1000), density=0.01, random_state=rng)
        a._dev(); b._dev()
        bd = b.todense()
        out = sp.tensordot(a, b, axes=1)
        rec("C1 COO(1000^2@.01) . COO -> COO (f64)", lambda: sp.tensordot(a, b, axes=1), 1e5, "Mproducts_per_ms_x1e-3")
        res["C1 COO(1000^2@.01) . COO -> COO (f64)"]["out_nnz"] = out.nnz
        bdev = torch.from_numpy(bd).to(DEV)
        rec("C1 COO . dense -> dense (device operand)", lambda: sp.tensordot(a, bdev, axes=1), a.nnz * 1000, "Gmadd_s")

    if "c3" in which or "red" in which:
        a = rand_coo((512, 512, 512, 64), 858_993, 0)
        b = rand_coo((512, 512, 512, 1), 13_421, 1)
        if "c3" in which:
            out = a + b
            alg = (a.nnz + b.nnz) * 16 + out.nnz * (32 + 8)
            rec("C3 COO add (512,512,512,64)+(512,512,512,1) f64", lambda: a + b, a.nnz + b.nnz, "Gnnz_in_s", alg)
            res["C3 COO add (512,512,512,64)+(512,512,512,1) f64"]["out_nnz"] = out.nnz
            rec("C3 multiply (same operands)", lambda: a * b, a.nnz + b.nnz, "Gnnz_in_s")
            a2 = rand_coo((512, 512, 512, 64), 858_993, 5)
            rec("C3 same-shape add", lambda: a + a2, 2 * a.nnz, "Gnnz_in_s")
        if "red" in which:
            rec("reduce sum axis=3", lambda: a.sum(axis=3), a.nnz, "Gnnz_s", a.nnz * 16)
            rec("reduce sum axis=(0,1)", lambda: a.sum(axis=(0, 1)), a.nnz, "Gnnz_s", a.nnz * 16)
            rec("reduce max axis=0", lambda: a.max(axis=0), a.nnz, "Gnnz_s", a.nnz * 16)

    if "c3big" in which:
        a = rand_coo((512, 512, 512, 64), 85_899_345, 10)
        b = rand_coo((512, 512, 512, 1), 1_342_177, 11)
        out = a + b
        alg = (a.nnz + b.nnz) * 16 + out.nnz * (32 + 8)
        rec("C3-large add density 1e-2 (8.6e7 nnz) f64", lambda: a + b, a.nnz + b.nnz, "Gnnz_in_s", alg, reps=3)
        res["C3-large add density 1e-2 (8.6e7 nnz) f64"]["out_nnz"] = out.nnz
        rec("reduce-large sum axis=3", lambda: a.sum(axis=3), a.nnz, "Gnnz_s", a.nnz * 16, reps=3)
        del a, b, out

    if "c5" in which:
        for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
            A = rand_csr(1_000_000, 1_000_000, 10_000_000, 3, dt)
            out = sp.tensordot(A, A, axes=1)
            products = None
            vb = 4 if dt == torch.float32 else 8
            alg = A.nnz * (vb + 4) + 4e6 + 1e8 * (vb + 4) + out.nnz * (vb + 8) + 8e6
            rec(f"C5 CSR(1e6^2@1e-5)^2 {name}", lambda: sp.tensordot(A, A, axes=1), out.nnz, "Gnnz_out_s", alg, reps=3)
            res[f"C5 CSR(1e6^2@1e-5)^2 {name}"]["out_nnz"] = out.nnz
            del A, out

    if "c4" in which:
        M = N = 1_000_000
        K = 256
        nnz = 100_000_000
        vals, cols, indptr, _ = bench.make_workload(torch, M, N, nnz, 1, 21, DEV)
        S = sp.GCXS((vals, cols, indptr), shape=(M, N), compressed_axes=(0,)).tocoo()
        g = torch.Generator(device=DEV).manual_seed(22)
        A = torch.rand((M, K), generator=g, device=DEV, dtype=torch.float32)
        B = torch.rand((K, N), generator=g, device=DEV, dtype=torch.float32)
        from sparse_b200 import _kernels as Kn
        from sparse_b200._dot import _coo_as_csr
        sv, sc, sip = _coo_as_csr(S, np.float32)
        Bt = Kn.transpose_dense(B)
        alg = nnz * 12 + M * K * 4 + nnz * K * 4 + nnz * 4
        rec("C4 SDDMM kernel only (1e6^2 mask nnz=1e8, K=256 f32)", lambda: Kn.sddmm(sip, sc, sv, A, Bt, M, N, K), nnz,
            "Gnnz_s", alg, reps=5)
        rec("C4 SDDMM public sddmm(s,a,b) incl. B transpose + prune", lambda: sp.sddmm(S, A, B), nnz, "Gnnz_s", alg, reps=3)
        del S, A, B, Bt

    if "mttkrp" in which:
        I_, K_, L_, J = 10_000, 10_000, 1_000, 32
        Bt = rand_coo((I_, K_, L_), 10_000_000, 31, torch.float32)
        g = torch.Generator(device=DEV).manual_seed(32)
        Dm = torch.rand((L_, J), generator=g, device=DEV, dtype=torch.float32)
        Cm = torch.rand((K_, J), generator=g, device=DEV, dtype=torch.float32)
        alg = Bt.nnz * 16 + 2 * Bt.nnz * J * 4 + I_ * J * 4
        rec("MTTKRP fused (1e4 x 1e4 x 1e3, nnz 1e7, J=32 f32)", lambda: sp.mttkrp(Bt, Dm, Cm), Bt.nnz, "Gnnz_s", alg)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

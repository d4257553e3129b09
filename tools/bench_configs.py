#!/usr/bin/env python
"""Time (and parity-check against the oracle / the reference's digests) the non-headline BASELINE.json configs (C1, C3, C4, C5, reductions, MTTKRP) through the public API.

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

DEV = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
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


def run(which=None, compact=False, parity=True):
    """Time the configs; returns {name: {ms_best, ..., frac_of_hbm_peak, parity}}.  `compact` = the subset bench.py
    embeds in its JSON line (one entry per BASELINE.json config + reductions + MTTKRP)."""
    import oracle  # the checker (tests / bench only)

    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import fullsize_inputs as FI

    which = which or ["c1", "c3", "red", "c5", "c4", "mttkrp", "c3big"]
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
        if not compact:
            print(name, r, flush=True)
        return r

    if "c1" in which:
        rng = np.random.default_rng(42)
        a = sp.random((1000, 1000), density=0.01, random_state=rng)
        b = sp.random((1000, 1000), density=0.01, random_state=rng)
        a._dev(); b._dev()
        bd = b.todense()
        out = sp.tensordot(a, b, axes=1)
        r = rec("C1 COO(1000^2@.01) . COO -> COO (f64)", lambda: sp.tensordot(a, b, axes=1), 1e5,
                "Mproducts_per_ms_x1e-3")
        r["out_nnz"] = out.nnz
        if parity:
            ip_a = np.searchsorted(a.coords[0], np.arange(1001))
            ip_b = np.searchsorted(b.coords[0], np.arange(1001))
            co, d = oracle.dot_coo_coo((1000, 1000), a.coords, b.coords, a.data, b.data, ip_a, ip_b)
            order = np.lexsort((co[1], co[0]))
            r["parity"] = {"vs": "oracle _dot_coo_coo + canonical sort", "coords_exact": bool(
                np.array_equal(out.coords, co[:, order])), "data_bit_exact": bool(
                np.array_equal(out.data.view(np.uint64), d[order].view(np.uint64)))}
        if not compact:
            bdev = torch.from_numpy(bd).to(DEV)
            rec("C1 COO . dense -> dense (device operand)", lambda: sp.tensordot(a, bdev, axes=1), a.nnz * 1000,
                "Gmadd_s")

    if "c3" in which or "red" in which:
        (ca, da), (cb, db) = FI.c3_inputs(np.float64)
        a = sp.COO(ca, da, shape=FI.C3_SHAPE_A, has_duplicates=False, sorted=True)
        b = sp.COO(cb, db, shape=FI.C3_SHAPE_B, has_duplicates=False, sorted=True)
        a._dev(); b._dev()
        want = {}
        try:
            with open(os.path.join(ROOT, "tests", "golden", "fullsize_digests.json")) as f:
                want = json.load(f)["cases"]
        except OSError:
            pass
        if "c3" in which:
            out = a + b
            alg = (a.nnz + b.nnz) * 16 + out.nnz * (32 + 8)
            r = rec("C3 COO add (512,512,512,64)+(512,512,512,1) f64", lambda: a + b, a.nnz + b.nnz, "Gnnz_in_s", alg)
            r["out_nnz"] = out.nnz
            if parity and want:
                w = want["c3_add_float64"]
                r["parity"] = {"vs": "the reference's own result (SHA-256 of coords + value bits, "
                                     "tests/golden/fullsize_digests.json)",
                               "nnz_equal": bool(out.nnz == w["nnz"]),
                               "digest_equal": bool(FI.coo_digest(out.coords, out.data) == w["sha256"])}
            if not compact:
                rec("C3 multiply (same operands)", lambda: a * b, a.nnz + b.nnz, "Gnnz_in_s")
        if "red" in which:
            for name, axis in (("reduce sum axis=3", (3,)), ("reduce sum axis=(0,1)", (0, 1)),
                               ("reduce max axis=0", (0,))):
                red = "max" if "max" in name else "sum"
                r = rec(name + " (C3 tensor)", lambda: getattr(a, red)(axis=axis), a.nnz, "Gnnz_s", a.nnz * 16)
                key = f"c3_{red}_axis{''.join(map(str, axis))}_float64"
                if parity and key in want:
                    got = getattr(a, red)(axis=axis)
                    w = want[key]
                    r["parity"] = {"vs": "the reference's own result (coords digest + sampled values, rtol 1e-12)",
                                   "coords_equal": bool(got.nnz == w["nnz"] and FI.digest(
                                       np.asarray(got.coords, dtype=np.int64)) == w["coords_sha256"]),
                                   "values_close": bool(np.allclose(got.data[:: w["sample_step"]], w["values_sample"],
                                                                    rtol=1e-12, atol=0))}
        del a, b

    if "c3big" in which:
        a = rand_coo((512, 512, 512, 64), 85_899_345, 10)
        b = rand_coo((512, 512, 512, 1), 1_342_177, 11)
        out = a + b
        alg = (a.nnz + b.nnz) * 16 + out.nnz * (32 + 8)
        r = rec("C3-large add density 1e-2 (8.6e7 nnz) f64", lambda: a + b, a.nnz + b.nnz, "Gnnz_in_s", alg, reps=3)
        r["out_nnz"] = out.nnz
        if parity:
            # size-independent properties: |union| = |a| + 64|b| - |matches|, every matched position holds a + b,
            # and the value total is the sum of the operands' totals (fp64, tolerance)
            ka, kb = a.sorted_keys(), b.sorted_keys()
            hit = torch.isin(torch.div(ka, 64, rounding_mode="floor"), kb)
            n_match = int(hit.sum().item())
            tot = float(out._data_dev().sum().item())
            exp = float(a._data_dev().sum().item()) + 64.0 * float(b._data_dev().sum().item())
            ko = out.sorted_keys()
            r["parity"] = {"vs": "properties (union size, strictly increasing keys, total)",
                           "union_size": bool(out.nnz == a.nnz + 64 * b.nnz - n_match),
                           "keys_strictly_increasing": bool((ko[1:] > ko[:-1]).all().item()),
                           "total_close": bool(abs(tot - exp) <= 1e-9 * abs(exp))}
            del ka, kb, hit, ko
        del out
        r = rec("reduce-large sum axis=3 (8.6e7 nnz)", lambda: a.sum(axis=3), a.nnz, "Gnnz_s", a.nnz * 16, reps=3)
        if parity:
            got = a.sum(axis=3)
            gk = torch.unique_consecutive(torch.div(a.sorted_keys(), 64, rounding_mode="floor"))
            tot, exp = float(got._data_dev().sum().item()), float(a._data_dev().sum().item())
            r["parity"] = {"vs": "properties (group ids = distinct key // 64, total)",
                           "groups_equal": bool(got.nnz == gk.numel() and torch.equal(got.sorted_keys(), gk)),
                           "total_close": bool(abs(tot - exp) <= 1e-9 * abs(exp))}
            del got, gk
        del a, b

    if "c5" in which:
        for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
            if compact and name == "f64":
                continue
            A = rand_csr(1_000_000, 1_000_000, 10_000_000, 3, dt)
            out = sp.tensordot(A, A, axes=1)
            vb = 4 if dt == torch.float32 else 8
            alg = A.nnz * (vb + 4) + 4e6 + 1e8 * (vb + 4) + out.nnz * (vb + 8) + 8e6
            r = rec(f"C5 CSR(1e6^2@1e-5)^2 {name}", lambda: sp.tensordot(A, A, axes=1), out.nnz, "Gnnz_out_s", alg,
                    reps=3)
            r["out_nnz"] = out.nnz
            if parity:
                # rows are independent: the first 20000 rows of the product against the oracle's _dot_csr_csr
                # (reverse-first-touch column order, then prune) on the same arrays
                rows = 20_000
                ad, ai, ap = A.data, A.indices.astype(np.int64), A.indptr.astype(np.int64)
                n = int(ap[rows])
                d, i, p = oracle.dot_csr_csr((rows, 1_000_000), ad[:n], ad, ai[:n], ai, ap[: rows + 1], ap)
                keep = d.view(np.uint32 if vb == 4 else np.uint64) != 0
                m = int(out.indptr[rows])
                r["parity"] = {"vs": f"oracle _dot_csr_csr on the first {rows} rows (all of B)",
                               "indices_exact": bool(m == int(keep.sum()) and np.array_equal(out.indices[:m], i[keep])),
                               "data_bit_exact": bool(m == int(keep.sum()) and np.array_equal(
                                   out.data[:m].view(np.uint8), d[keep].view(np.uint8)))}
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
        r = rec("C4 SDDMM kernel only (1e6^2 mask nnz=1e8, K=256 f32)", lambda: Kn.sddmm(sip, sc, sv, A, Bt, M, N, K),
                nnz, "Gnnz_s", alg, reps=5)
        if parity:
            ov = Kn.sddmm(sip, sc, sv, A, Bt, M, N, K)
            gsel = torch.Generator(device=DEV).manual_seed(23)
            sel = torch.randint(0, nnz, (4096,), generator=gsel, device=DEV)
            rows_of = torch.searchsorted(sip.to(torch.int64), sel, right=True) - 1
            dots = (A[rows_of].double() * Bt[sc[sel].to(torch.int64)].double()).sum(dim=1)  # fp64 restatement
            wantv = sv[sel].double() * dots
            r["parity"] = {"vs": "fp64 restatement of s * dot(A[i,:], B[:,j]) on 4096 sampled entries, rtol 2e-5",
                           "values_close": bool(torch.allclose(ov[sel].double(), wantv, rtol=2e-5, atol=0))}
            del ov
        if not compact:
            rec("C4 SDDMM public sddmm(s,a,b) incl. B transpose + prune", lambda: sp.sddmm(S, A, B), nnz, "Gnnz_s",
                alg, reps=3)
        del S, A, B, Bt, sv, sc, sip, vals, cols, indptr

    if "mttkrp" in which:
        I_, K_, L_, J = 10_000, 10_000, 1_000, 32
        Bt = rand_coo((I_, K_, L_), 10_000_000, 31, torch.float32)
        g = torch.Generator(device=DEV).manual_seed(32)
        Dm = torch.rand((L_, J), generator=g, device=DEV, dtype=torch.float32)
        Cm = torch.rand((K_, J), generator=g, device=DEV, dtype=torch.float32)
        alg = Bt.nnz * 16 + 2 * Bt.nnz * J * 4 + I_ * J * 4
        r = rec("MTTKRP fused (1e4 x 1e4 x 1e3, nnz 1e7, J=32 f32)", lambda: sp.mttkrp(Bt, Dm, Cm), Bt.nnz, "Gnnz_s",
                alg)
        if parity:
            got = sp.mttkrp(Bt, Dm, Cm)
            got = got.todense_device() if hasattr(got, "todense_device") else torch.as_tensor(np.asarray(got)).to(DEV)
            co, dv = Bt._dev()
            contrib = dv.double()[:, None] * Dm[co[2]].double() * Cm[co[1]].double()
            wantm = torch.zeros((I_, J), dtype=torch.float64, device=DEV).index_add_(0, co[0], contrib)
            r["parity"] = {"vs": "fp64 restatement sum_kl B[i,k,l] D[l,j] C[k,j], rtol 2e-5",
                           "values_close": bool(torch.allclose(got.double(), wantm, rtol=2e-5, atol=1e-9))}
            del contrib, wantm
    torch.cuda.empty_cache()
    return res


def main():
    res = run(sys.argv[1:] or None)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

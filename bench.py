#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: configs[1] = C2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload C2: GCXS/CSR A (1e6 x 1e6, nnz 1e8, fp32, uniform-random positions) times dense
B (1e6 x 128 fp32) -> dense C, i.e. ``sparse.tensordot(A, B, axes=1)`` -> _dot_csr_ndarray
(reference: sparse/numba_backend/_common.py:720-755).  A "step" is one full product.

* ``value``    GNNZ/s, device-resident inputs, CUDA-event timed, max over ranks.
* ``e2e``      the same product through the host-buffer C-ABI call (b2s_spmm_csr_dense_host):
               pinned host arrays in, host array out, H2D/D2H inside the timed region.
* ``roofline`` algorithmic bytes (gather model, SURVEY.md s8(d): 525 B/nnz with int32 indices)
               / measured kernel time, against MEASURED_PEAKS.json's HBM copy bandwidth.
* ``cpu_baseline``  the oracle port (oracle/dot_oracle.c, gcc -O3, OpenMP over rows) timed on this
               box's host cores on a bounded row-sample of the same workload.
* N > 1: weak scaling -- every rank owns a 1e6-row block of A (nnz 1e8) and a K/N row shard of B;
  each step all-gathers B over NCCL/NVLink and runs K1 on the local row block (no other collective).

Inputs are synthetic (seeded torch generators on the device); 2.2 GB of operands per rank is far
larger than the 126 MB L2, so no explicit L2 flush is needed between steps ("l2": "inputs>L2").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--rows", type=int, default=1_000_000, help="M (= K) of the C2 workload")
    p.add_argument("--nnz", type=int, default=100_000_000)
    p.add_argument("--ncols", type=int, default=128)
    p.add_argument("--variant", type=int, default=0, help="K1 variant override (1 LDG, 2 bulk-TMA)")
    p.add_argument("--unroll", type=int, default=0)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--cpu-rows", type=int, default=0, help="rows of A in the CPU-baseline sample (0 = auto)")
    return p.parse_args()


# ----------------------------------------------------------------------------------------------
def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def known_traffic(ncols):
    """Per-launch DRAM bytes of K1 from the committed ncu --set full capture (profiles/), if any."""
    path = os.path.join(ROOT, "profiles", "k1_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        if int(d.get("ncols", 128)) == ncols:
            return float(d["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        # one nvidia-smi process in loop mode (-lms 20): dense samples even for a timed region of ~100 ms
        try:
            self._proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self._proc.stdout:
                line = line.strip()
                if line:
                    self.samples.append([x.strip() for x in line.split(",")])
                if self._stop.is_set():
                    break
        except Exception:
            pass

    def __enter__(self):
        self._proc = None
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        time.sleep(0.15)  # let the sampler start before the timed region
        return self

    def __exit__(self, *a):
        time.sleep(0.05)
        self._stop.set()
        if self._proc is not None:
            try:
                self._proc.terminate()
            except Exception:
                pass
        self._t.join(timeout=3)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
def make_workload(torch, M, K, nnz, ncols, seed, device):
    """Uniform-random CSR (sorted unique columns per row) + dense B, generated on the device."""
    g = torch.Generator(device=device).manual_seed(seed)
    lin = torch.randint(0, M * K, (int(nnz * 1.0006) + 1024,), generator=g, device=device, dtype=torch.int64)
    lin = torch.unique(lin)  # sorted
    extra = lin.numel() - nnz
    if extra > 0:  # drop `extra` entries at random positions -> exactly nnz
        keep = torch.ones(lin.numel(), dtype=torch.bool, device=device)
        while extra > 0:
            idx = torch.randint(0, lin.numel(), (extra,), generator=g, device=device)
            keep[idx] = False
            extra = int(keep.sum().item()) - nnz
            if extra < 0:  # dropped too many duplicates impossible; re-add is not needed
                break
        lin = lin[keep]
    rows = torch.div(lin, K, rounding_mode="floor")
    cols = (lin - rows * K).to(torch.int32)
    counts = torch.bincount(rows, minlength=M)
    indptr = torch.zeros(M + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(counts, 0)
    del lin, rows, counts
    vals = torch.rand(cols.numel(), generator=g, device=device, dtype=torch.float32)
    B = torch.rand((K, ncols), generator=g, device=device, dtype=torch.float32)
    return vals, cols, indptr.to(torch.int32), B


def algorithmic_bytes(nnz, M, ncols, vb=4, ib=4):
    """SURVEY.md s8(d) gather model: A stream + indptr + one B row per nnz + C written once."""
    return nnz * (vb + ib) + (M + 1) * ib + nnz * ncols * vb + M * ncols * vb


def cpu_baseline(vals, cols, indptr, B, M, K, ncols, rows_sample, threads=None):
    """Time the oracle port on the first `rows_sample` rows of the same A (bounded sample)."""
    import oracle

    ip = indptr[: rows_sample + 1].cpu().numpy().astype(np.int64)
    n = int(ip[-1])
    a_data = vals[:n].cpu().numpy()
    a_idx = cols[:n].cpu().numpy().astype(np.int64)
    Bh = B.cpu().numpy()
    oracle.dot_csr_ndarray((min(rows_sample, 64), ncols), a_data[: ip[min(rows_sample, 64)]],
                           a_idx[: ip[min(rows_sample, 64)]], ip[: min(rows_sample, 64) + 1], Bh)  # warm-up
    best = None
    reps = 0
    t_all = time.perf_counter()
    while reps < 3 and (time.perf_counter() - t_all) < 25:
        t0 = time.perf_counter()
        out = oracle.dot_csr_ndarray((rows_sample, ncols), a_data, a_idx, ip, Bh)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        reps += 1
    return {"value": n / best / 1e9, "unit": "GNNZ/s", "cores": oracle.max_threads(), "kind": "port",
            "sample": f"first {rows_sample} rows of A ({n} nnz) x full B, best of {reps}, "
                      f"oracle/dot_oracle.c (gcc -O3, OpenMP {oracle.max_threads()} threads)",
            "seconds": best}, out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    M = K = args.rows
    ncols = args.ncols

    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch

    from sparse_b200 import _kernels as Kn
    from sparse_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sparse_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    if args.variant or args.unroll:
        Kn.spmm_set_variant(args.variant or 1, args.unroll or 8)

    vals, cols, indptr, B = make_workload(torch, M, K, args.nnz, ncols, seed=1234 + rank, device=dev)
    nnz = int(vals.numel())
    C = torch.empty((M, ncols), dtype=torch.float32, device=dev)
    if world > 1:
        assert K % world == 0
        shard = K // world
        B_shard = B[rank * shard:(rank + 1) * shard].clone()
        B_full = torch.empty_like(B)

    # N > 1: every step needs its own all-gather of the row-sharded B.  The gather of step i+1 runs on a side
    # stream into the other half of a double buffer while K1 of step i runs, so the NVLink transfer is hidden behind
    # the kernel (events order buffer reuse: gather(i+2) waits for K1(i), K1(i) waits for gather(i)).
    if world > 1:
        comm = torch.cuda.Stream(device=dev)
        B_bufs = [B_full, torch.empty_like(B)]
        ev_gathered = [torch.cuda.Event(), torch.cuda.Event()]
        ev_consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def issue_gather(i):
            slot = i % 2
            with torch.cuda.stream(comm):
                comm.wait_event(ev_consumed[slot])  # K1 that last read this buffer has finished
                dist.all_gather_into_tensor(B_bufs[slot], B_shard)
                ev_gathered[slot].record(comm)

        def run_steps(n, kev=None):
            main = torch.cuda.current_stream()
            for slot in (0, 1):
                ev_consumed[slot].record(main)
            issue_gather(0)
            for i in range(n):
                slot = i % 2
                if i + 1 < n:
                    issue_gather(i + 1)
                main.wait_event(ev_gathered[slot])
                if kev is not None:
                    kev[i][0].record()
                Kn.spmm_csr_dense(vals, cols, indptr, B_bufs[slot], M, K, ncols, out=C)
                if kev is not None:
                    kev[i][1].record()
                ev_consumed[slot].record(main)
    else:
        def run_steps(n, kev=None):
            for i in range(n):
                if kev is not None:
                    kev[i][0].record()
                Kn.spmm_csr_dense(vals, cols, indptr, B, M, K, ncols, out=C)
                if kev is not None:
                    kev[i][1].record()

    run_steps(max(args.warmup, 3))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = _lib.launch_count()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clk:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        t_start.record()
        run_steps(args.steps, kev)
        t_end.record()
        torch.cuda.synchronize()
    launches = _lib.launch_count() - launches0
    total_ms = t_start.elapsed_time(t_end)
    kern_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    if world > 1:
        tt = torch.tensor([total_ms, float(nnz)], device=dev, dtype=torch.float64)
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms = float(tmax[0].item())
        nnz_all = float(tsum[1].item())
        dist.barrier()
    else:
        nnz_all = float(nnz)
    ms_per_step = total_ms / args.steps
    value = nnz_all / (ms_per_step * 1e-3) / 1e9

    peak, peak_src = peaks()
    alg = algorithmic_bytes(nnz, M, ncols)
    achieved = alg / (kern_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": known_traffic(ncols),
            "kernel": "spmm_csr_dense_kernel<f32,i32,VEC4,G32>" if not args.variant or args.variant == 1
            else "spmm_csr_dense_tma_kernel<f32,i32>",
            "kernel_ms": round(kern_ms, 4), "algorithmic_bytes_per_launch": alg,
            "bytes_per_nnz_model": round(alg / nnz, 2), "peak_source": peak_src,
            "compulsory_bytes_per_launch": nnz * 8 + (M + 1) * 4 + K * ncols * 4 + M * ncols * 4}

    # ---- e2e through the host-buffer C-ABI call (rank-local; N>1 reports the aggregate) ------------
    e2e = None
    if not args.no_e2e:
        h_vals = vals.cpu().pin_memory()
        h_cols = cols.cpu().to(torch.int64).pin_memory()
        h_ptr = indptr.cpu().to(torch.int64).pin_memory()
        h_B = B.cpu().pin_memory()
        h_C = torch.empty((M, ncols), dtype=torch.float32).pin_memory()
        import sparse_b200 as sp

        npv = (h_vals.numpy(), h_cols.numpy(), h_ptr.numpy(), h_B.numpy())
        esteps = max(3, min(args.steps, 10))

        def e2e_step():
            # the call a user of the reference makes: host arrays in, np.ndarray out (H2D + K1 + D2H inside)
            A = sp.GCXS((npv[0], npv[1], npv[2]), shape=(M, K), compressed_axes=(0,))
            return sp.tensordot(A, npv[3], axes=1)

        for _ in range(2):  # warm-up (pinned result pool, scratch pool, page mapping)
            h_res = e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            h_res = e2e_step()
        torch.cuda.synchronize()
        e_ms = (time.perf_counter() - t0) * 1e3 / esteps
        h_C.copy_(torch.from_numpy(h_res))
        if world > 1:
            tt = torch.tensor([e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e_ms = float(tt.item())
        # bytes that cross PCIe per step: values, column indices (int64 on the host, narrowed to int32 by the
        # library's host thread pool into pinned staging before the copy), int64 indptr, B; and C coming back
        host_in = int(h_vals.numel() * 4 + h_cols.numel() * 8 + h_ptr.numel() * 8 + h_B.numel() * 4)
        h2d = int(h_vals.numel() * 4 + h_cols.numel() * 4 + h_ptr.numel() * 8 + h_B.numel() * 4)
        d2h = int(h_C.numel() * 4)
        e2e = {"value": round(nnz_all / (e_ms * 1e-3) / 1e9, 4), "unit": "GNNZ/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "host_input_bytes_per_step": host_in, "ms_per_step": round(e_ms, 3),
               "steps": esteps,
               "api": "sparse_b200.tensordot(GCXS(host arrays, int64 indices), np.ndarray) -> np.ndarray "
                      "(b2s_spmm_csr_dense_host: host-side int64->int32 index narrowing + 3-stream H2D/K1/D2H "
                      "pipeline, pinned buffers)"}
        C_ref = C if world == 1 else Kn.spmm_csr_dense(vals, cols, indptr, B, M, K, ncols)
        same = bool(torch.equal(h_C.to(dev), C_ref))
        e2e["matches_device_path"] = same

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rows_sample = args.cpu_rows or min(M, 200_000)
        cpu, ref_out = cpu_baseline(vals, cols, indptr, B, M, K, ncols, rows_sample)
        got = C[:rows_sample].cpu().numpy()
        cpu["parity_bit_exact_vs_gpu"] = bool(np.array_equal(got.view(np.uint32), ref_out.view(np.uint32)))
        cpu.pop("seconds", None)

    if rank == 0:
        line = {
            "metric": "CSR x dense tensordot throughput (GNNZ/s)", "value": round(value, 4), "unit": "GNNZ/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(M, K, nnz, ncols, world),
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clk.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def workload_config(M, K, nnz, ncols, world):
    """The `config` object of the JSON line -- identical for the product arm and the reference arm."""
    return {"workload": f"C2: GCXS/CSR({M}x{K}, nnz={nnz} per GPU, uniform) @ dense({K}x{ncols}) fp32 -> dense",
            "index_dtype_device": "int32", "l2": "inputs>L2 (2.2 GB operands vs 126 MB L2), no flush",
            "parallelism": "1-D row blocks of A per GPU; B row-sharded, one NCCL all-gather per step, "
                           "double-buffered on a side stream (overlaps the previous step's K1)"
            if world > 1 else "single GPU",
            "exact_order": True}


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores.

    /root/reference (Python + numba) cannot travel to the GPU box, so this arm times the oracle
    port of _dot_csr_ndarray (oracle/dot_oracle.c) with all host threads, on a bounded sample of
    the same workload (first `cpu_rows` rows of A x full B per step)."""
    if rank != 0:
        return
    import oracle

    M = K = args.rows
    ncols = args.ncols
    rows_sample = args.cpu_rows or min(M, 100_000)
    rng = np.random.default_rng(1234)
    per_row = max(1, args.nnz // M)
    nnz_s = rows_sample * per_row
    # same distribution as the GPU arm (uniform positions, ~nnz/M per row); generated on the host
    lin = np.unique(rng.integers(0, rows_sample * K, size=int(nnz_s), dtype=np.int64))
    rows, cols = lin // K, lin % K
    indptr = np.zeros(rows_sample + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=rows_sample), out=indptr[1:])
    data = rng.random(len(lin), dtype=np.float32)
    B = rng.random((K, ncols), dtype=np.float32)
    n = len(lin)
    for _ in range(max(args.warmup, 1)):
        oracle.dot_csr_ndarray((rows_sample, ncols), data, cols, indptr, B)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.dot_csr_ndarray((rows_sample, ncols), data, cols, indptr, B)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt / 1e9
    sample = f"{rows_sample} rows of A ({n} nnz) x full B({K}x{ncols}) per step"
    line = {
        "impl": "reference", "metric": "CSR x dense tensordot throughput (GNNZ/s)", "value": round(val, 5),
        "unit": "GNNZ/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
        "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(M, K, args.nnz, ncols, world),
        "cpu_baseline": {"value": round(val, 5), "unit": "GNNZ/s", "cores": oracle.max_threads(), "kind": "port",
                         "sample": sample + "; oracle/dot_oracle.c = C restatement of _dot_csr_ndarray "
                                            "(the numba reference itself is single-threaded and cannot travel)"},
        "e2e": {"value": round(val, 5), "unit": "GNNZ/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Collect the multi-GPU runs of a round (tools/gpu_multi.sh N -> gpurun_out/) into profiles/r02_scaling.md and copy the
JSON lines next to it.  Efficiencies are computed here from the per-N values (the bench never reports one)."""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"


def last_json(path):
    try:
        lines = [x for x in open(path).read().splitlines() if x.startswith("{")]
        return json.loads(lines[-1]) if lines else None
    except OSError:
        return None


def multi(path):
    try:
        return json.load(open(path))
    except (OSError, ValueError):
        return None


def main():
    ns = [1, 2, 4, 8]
    ce = {n: last_json(os.path.join(OUT, f"bench_n{n}_ce.json")) for n in ns}
    if ce[1] is None:
        ce[1] = last_json(os.path.join(PROF, f"{TAG}_bench_n1.json"))
    nc = {n: last_json(os.path.join(OUT, f"bench_n{n}_nccl.json")) for n in ns}
    mu = {n: multi(os.path.join(OUT, f"multi_{n}.json")) for n in ns}
    for n in ns:
        for src, name in ((f"bench_n{n}_ce.json", f"{TAG}_bench_n{n}_copy_engine_gather.json"),
                          (f"bench_n{n}_nccl.json", f"{TAG}_bench_n{n}_nccl_gather.json"),
                          (f"multi_{n}.json", f"{TAG}_multi_{n}.json")):
            p = os.path.join(OUT, src)
            if os.path.exists(p) and n > 1 or (os.path.exists(p) and src.startswith("multi")):
                shutil.copy(p, os.path.join(PROF, name))
    base = ce[1]["value"] if ce[1] else None
    # runs made before the dynamic-row K1 became the default are compared with the N = 1 line of THEIR build
    first = last_json(os.path.join(PROF, f"{TAG}_bench_n1_first.json"))
    base_static = first["value"] if first else None

    def base_for(d):
        static = "dyn" not in d["roofline"].get("kernel", "dyn")
        return (base_static if static and base_static else base), static
    L = [f"# Round 2 — scaling over the GPUs of one B200 box ({TAG})\n",
         "One process per GPU (`torchrun`), CUDA events on the launching stream, barrier + synchronize on both sides, MAX",
         "over ranks; `tools/gpu_multi.sh N` under `gpurun --gpus N`.  Efficiencies are computed in this table from the",
         "per-N values.  The GPU budget of the round ended after the last N = 8 call: N = 1 and N = 8 are the FINAL build (K1 with",
         "dynamic row assignment), N = 2 and N = 4 were measured before that change (static-grid K1) and are compared with",
         "the N = 1 line of their own build; the bench_multi rows (C5 / C4 / C3-large / reductions) do not involve K1.\n",
         "## C2 weak scaling (1e8 nnz per GPU; B row-sharded, gathered EVERY step, double-buffered behind K1)\n",
         "| N | gather | ms/step | K1 ms (rank 0) | GNNZ/s (all GPUs) | efficiency vs N x (N=1) | distributed product bit-exact |",
         "|---|---|---|---|---|---|---|"]
    for n in ns:
        for kind, d in (("copy engines (CUDA IPC + cudaMemcpyAsync pull)", ce[n]), ("NCCL all-gather (SM kernels)", nc[n])):
            if d is None or (n == 1 and kind.startswith("NCCL")):
                continue
            b_, static = base_for(d)
            eff = f"{d['value'] / (n * b_):.3f}" + (" (static-grid K1 build, vs its N = 1: " + str(b_) + ")" if static else "") if b_ else ""
            L.append(f"| {n} | {'—' if n == 1 else kind} | {d['ms_per_step']} | {d['roofline']['kernel_ms']} | {d['value']} | "
                     f"{eff} | {d['config'].get('distributed_product_bit_exact_vs_local_B', '—')} |")
    L += ["", "## C2 strong scaling (the named problem: 1e8 nnz in total, nnz-balanced row blocks)\n",
          "| N | B replicated (no collective): ms, GNNZ/s, speed-up | B row-sharded (gather every step): ms, GNNZ/s, speed-up | row blocks bit-exact |",
          "|---|---|---|---|"]
    t1 = ce[1]["ms_per_step"] if ce[1] else None
    t1_static = first["ms_per_step"] if first else None
    if ce[1]:
        L.append(f"| 1 | {t1} ms, {ce[1]['value']}, 1.00 | = | — |")
    for n in ns[1:]:
        d = ce[n]
        if d is None or "strong" not in d:
            continue
        s = d["strong"]
        _, static = base_for(d)
        tb = t1_static if static and t1_static else t1
        tag = f" (static-grid K1 build, vs its N = 1: {tb} ms)" if static else ""
        L.append(f"| {n} | {s['replicated_B']['ms_per_step']} ms, {s['replicated_B']['GNNZ/s']}, "
                 f"{tb / s['replicated_B']['ms_per_step']:.2f}{tag} | {s['sharded_B']['ms_per_step']} ms, "
                 f"{s['sharded_B']['GNNZ/s']}, {tb / s['sharded_B']['ms_per_step']:.2f} | "
                 f"{s['row_blocks_bit_exact_vs_single_gpu_product']} |")
    L += ["", "## C2 end to end (host buffers in, host array out; per-rank PCIe; 1e8 nnz per GPU)\n",
          "| N | ms/step (max over ranks) | GNNZ/s (all GPUs) | index narrowing | the other staging mode |", "|---|---|---|---|---|"]
    for n in ns:
        d = ce[n]
        if d is None or not d.get("e2e"):
            continue
        e = d["e2e"]
        o = e.get("other_staging_mode")
        L.append(f"| {n} | {e['ms_per_step']} | {e['value']} | {e.get('index_narrowing', 'host')} | "
                 f"{(o['index_narrowing'] + ': ' + str(o['ms_per_step']) + ' ms') if o else '—'} |")
    L += ["", "## C5 / C4 / C3-large over N row blocks (`tools/bench_multi.py`, STRONG scaling: the same global problem for every N)\n"]
    keys = []
    for n in ns:
        if mu[n]:
            for k in mu[n]:
                if k not in keys:
                    keys.append(k)
    for k in keys:
        L += [f"### {k}\n", "| N | ms/step | rate | speed-up vs N=1 | check |", "|---|---|---|---|---|"]
        t_1 = mu[1][k]["ms_per_step"] if mu[1] and k in mu[1] else None
        for n in ns:
            if not mu[n] or k not in mu[n]:
                continue
            r = mu[n][k]
            rate = next((f"{r[x]} {x}" for x in ("Gnnz_out_s", "Gnnz_s", "Gnnz_in_s") if x in r), "")
            chk = {x: r[x] for x in r if x.startswith(("row_block", "block_bit", "matches", "checksum"))}
            L.append(f"| {n} | {r['ms_per_step']} | {rate} | {(t_1 / r['ms_per_step']):.2f} | {chk} |" if t_1 else
                     f"| {n} | {r['ms_per_step']} | {rate} | — | {chk} |")
        L.append("")
    with open(os.path.join(PROF, f"{TAG}_scaling.md"), "w") as f:
        f.write("\n".join(L) + "\n")
    print("\n".join(L))


if __name__ == "__main__":
    main()

"""tools/bench_multi.py (C4 SDDMM / C5 SpGEMM over row blocks) in its CPU smoke mode: world_size 2 over gloo on the
NumPy mock of the kernel layer; the script itself asserts that every rank's block equals the single-process product."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_multi_cpu_smoke_world2():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29537", os.path.join(ROOT, "tools", "bench_multi.py"), "--cpu-smoke"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [json.loads(line) for line in r.stdout.splitlines() if line.startswith("{")]
    got = {k: v for d in lines for k, v in d.items()}
    assert set(got) == {"C4", "C5", "C5 contraction-split"}
    assert got["C5 contraction-split"]["matches_row_blocked_result_rtol_1e-5"] is True
    assert got["C5"]["n_gpus"] == 2 and got["C5"]["out_nnz"] > 0 and got["C4"]["mask_nnz"] > 0

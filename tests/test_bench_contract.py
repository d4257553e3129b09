"""bench.py contract checks that do not need a GPU: the reference arm prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "4000",
                          "--nnz", "80000", "--ncols", "128", "--steps", "2", "--warmup", "1", "--cpu-rows", "2000"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "GNNZ/s" and line["value"] > 0
    cb = line["cpu_baseline"]
    if os.path.exists(os.path.join(ROOT, "baseline", "_ref", "sparse", "__init__.py")):
        # the reference itself (numba, single-threaded by construction) + the labelled all-cores port figure
        assert cb["kind"] == "reference" and cb["cores"] == 1 and "numba" in cb and cb["value"] == line["value"]
        assert cb["all_cores_port"]["kind"] == "port" and cb["all_cores_port"]["cores"] >= 1
    else:  # baseline/_ref not installed (tools/make_ref.sh): the C port stands in and says so
        assert cb["kind"] == "port" and cb["cores"] >= 1 and "baseline/_ref missing" in cb["sample"]
    assert "same arrays" in cb["sample"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_other_ranks_of_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--rows", "1000", "--nnz", "10000", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""

"""`bench.py --impl reference` on the CPU: the JSON contract of the reference arm (same metric / unit / config keys as the
product arm, `impl`, `cpu_baseline`, an `e2e` that moves no bytes), no product library in the process, and ranks other than
0 leaving without output.  Run on BASELINE config 2 with a 3-second budget."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RUNNER = r"""
import runpy, sys
sys.argv = ["bench.py", "--impl", "reference", "--config", "2", "--steps", "2", "--warmup", "1"]
runpy.run_path(%r, run_name="__main__")
maps = open("/proc/self/maps").read()
print("MAPS_LIBMNEMCU=%%d" %% int("libmnemcu" in maps))
print("MAPS_ORACLE=%%d" %% int("libmnem_oracle" in maps))
""" % os.path.join(ROOT, "bench.py")


def run(extra_env):
    env = dict(os.environ, MNEM_REF_BUDGET_S="3")
    env.update(extra_env)
    r = subprocess.run([sys.executable, "-c", RUNNER], capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_line_and_loaded_libraries():
    out = run({})
    lines = [l for l in out if l.startswith("{")]
    assert len(lines) == 1                                   # ONE JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "em_iterations_per_s" and d["unit"] == "EM iterations/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("cfg2:") and "model" not in d["config"]
    assert d["value"] > 0 and len(d["values_per_step"]) == 2
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "oracle port" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the arm times the CPU restatement and nothing of the product
    assert "MAPS_LIBMNEMCU=0" in out and "MAPS_ORACLE=1" in out


def test_reference_arm_other_ranks_leave_quietly():
    out = run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert not [l for l in out if l.startswith("{")]
    assert "MAPS_LIBMNEMCU=0" in out


def test_product_arm_refuses_to_run_without_a_device():
    """No CPU fallback: without a CUDA device the product arm exits non-zero with a message and prints no result line."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode != 0
    assert "no CPU path" in (r.stdout + r.stderr)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]

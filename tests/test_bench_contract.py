"""CPU-only checks of bench.py's contract: the reference arm (`--impl reference`: the oracle port of the reference's CPU
path on the host cores) prints the agreed JSON line, and the GPU arm refuses to run without a CUDA device (the product has
no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "poseidon_hashes_per_s_depth24_tree_build" and line["unit"] == "hashes/s"
    assert line["higher_is_better"] is True and line["steps"] == 1 and line["n_gpus"] == 1 and line["vs_baseline"] is None
    assert line["config"]["depth"] == 24 and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] > 1000 and "sample" in cb
    assert cb["single_thread"]["value"] > 1000 and "1 thread" in cb["single_thread"]["sample"]   # the reference is sequential (utils.rs:21)
    assert line["e2e"] == {"value": line["value"], "unit": "hashes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    if _has_gpu():
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)

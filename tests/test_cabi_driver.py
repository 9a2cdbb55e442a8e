"""The C-ABI from a plain C client (tests/cabi_driver.c, gcc, no CUDA headers): proves the boundary is what a Rust / cgo
binding would link — `extern "C"`, pointers and sizes only. Without a GPU the driver must fail cleanly (no CPU fallback);
with one it runs the reference's deterministic insert scenario and prints values this test compares with the fixtures."""
import json
import os
import subprocess

import pytest

from imt_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


def _build():
    _ffi.load()
    lib = _ffi.library_path()
    out = os.path.join(ROOT, "tests", "_build", "cabi_driver")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cabi_driver.c")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", out,
                    "-L", os.path.dirname(lib), "-limt_b200", f"-Wl,-rpath,{os.path.dirname(lib)}"], check=True,
                   env={k: v for k, v in os.environ.items() if k not in ("CC", "CXX")})
    return out


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_c_client_compiles_links_and_fails_cleanly_without_a_gpu():
    exe = _build()
    if _has_gpu():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)          # IMT_ERR_CUDA from imt_ctx_create, nothing else ran
    assert "no-gpu status=100" in r.stdout


@pytest.mark.gpu
def test_c_client_runs_the_reference_scenario():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    lines = r.stdout.strip().splitlines()
    fe = lambda s: int(s.split()[1], 16)
    assert fe(lines[0]) == int(GOLD["kat_h3_zero"])
    assert lines[1] == "empty: 1 Cannot create Merkle Tree with no leaves" and lines[2] == "odd: 2 Leaves must be even"
    assert fe(lines[3]) == int(GOLD["empty_depth3_root"])
    rounds = lines[4:16]
    assert [int(l.split()[1]) for l in rounds[0::2]] == GOLD["scenario_low_idx"]
    assert [str(fe(l)) for l in rounds[1::2]] == GOLD["scenario_roots"]
    assert [int(l.split()[3]) for l in rounds[0::2]] == [1, 0, 0, 0, 1, 0]      # is_new_leaf_largest per round (IMT:736-741)
    assert lines[16] == "non_inclusion low_idx 3 val 20 next 30 matched 1 largest 0"
    assert lines[17] == "trace_ends_in_root 1"
    assert lines[18] == "limbs nl_r 25 ll_r 30 llv_r 20 flags 111"
    assert fe(lines[19]) == int(GOLD["spec"]["published_perm_x5_254_5_input_0_to_4"][0], 16)   # published Poseidon test vector
    assert fe(lines[20]) == int(GOLD["spec"]["hashes"]["5,8,60"]["6"])
    assert lines[21] == f"trace_fe {(6 // 4 + 1) * (1 + 8 + 60) * 5}"
    assert lines[22] == "checkpoint n 8 depth 3 same_root 1 same_leaves 1 header_root 1"
    assert lines[23].startswith("corrupt checkpoint: 6 ") and "corrupt" in lines[23]      # IMT_ERR_INVALID_ARG, root mismatch
    assert lines[24] == "insert_trace_ok 1"
    assert lines[25] == "non_inclusion_trace_ok 1"

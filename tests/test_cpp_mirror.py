"""The C++ host mirror of the reference's Rust API (include/imt_b200.hpp) running the native halves of the reference's own
tests (tests/reference_tests.cpp = indexed_merkle_tree.rs:360-810 with the reference's names and flow). g++, no CUDA
headers. Without a GPU the binary must stop at Poseidon::new_ (no CPU fallback); with one, every printed value is compared
with the committed fixtures (which come from the oracle pinned on indexed_merkle_tree.rs:247-251)."""
import json
import os
import subprocess

import pytest

from imt_b200 import _ffi
import poseidon_ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


def _build():
    _ffi.load()
    lib = _ffi.library_path()
    out = os.path.join(ROOT, "tests", "_build", "reference_tests")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "reference_tests.cpp"), "-o", out, "-L", os.path.dirname(lib), "-limt_b200",
                    f"-Wl,-rpath,{os.path.dirname(lib)}"], check=True, env={k: v for k, v in os.environ.items() if k not in ("CC", "CXX")})
    return out


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_cpp_mirror_compiles_and_stops_cleanly_without_a_gpu():
    exe = _build()
    if _has_gpu():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 3 and r.stdout.startswith("no-gpu:"), (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_cpp_mirror_runs_the_reference_tests():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    out = {}
    for line in r.stdout.strip().splitlines():
        k, _, v = line.rpartition(" ")
        out[k] = v
    lines = r.stdout.strip().splitlines()
    assert lines[-1] == "all reference tests passed"
    assert int(out["test_hash_zero"], 16) == int(GOLD["kat_h3_zero"])                      # indexed_merkle_tree.rs:247-251
    assert int(out["multiple_round empty_root"], 16) == int(GOLD["empty_depth3_root"])
    rounds = [l.split() for l in lines if l.startswith("multiple_round ") and " low_idx " in l]
    assert [int(f[3]) for f in rounds] == GOLD["scenario_low_idx"]
    assert [str(int(f[7], 16)) for f in rounds] == GOLD["scenario_roots"]
    assert [int(f[5]) for f in rounds] == [1, 0, 0, 0, 1, 0]
    final = [l.split()[3:] for l in lines if l.startswith("multiple_round final")]
    assert final == GOLD["scenario_final_preimages"]
    assert "new_errors empty: Cannot create Merkle Tree with no leaves" in lines            # utils.rs:25
    assert "new_errors odd: Leaves must be even" in lines                                   # utils.rs:35
    # test_insert_leaf with the fixed value: the oracle replays the two inserts
    v = 0x2a3b4c5d6e7f80910f1e2d3c4b5a6978fedcba98765432100123456789abcdef
    assert v < R.P
    rnds, _ = R.insert_rounds(3, [v, 42])
    assert int(out["test_insert_leaf root1"], 16) == rnds[0]["new_root"] and int(out["test_insert_leaf root2"], 16) == rnds[1]["new_root"]
    assert int(out["other_instance h3"], 16) == R.hash_n([1, 2, 3], R.Spec(8, 56, 4))

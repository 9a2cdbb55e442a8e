"""CPU-only checks of the drop-in boundary: the library builds, loads, and exports every symbol include/imt_b200.h
declares (no compute calls without a GPU); the product never touches oracle/."""
import os
import re

import pytest

import imt_b200
from imt_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "indexed-merkle-tree-halo2_b200")


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "imt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(imt_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    lib = _ffi.load()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/imt_b200.h but not exported"
    assert sorted(_ffi.SIGNATURES) == declared, "ctypes signature table out of sync with the header"


def test_status_strings_are_the_reference_messages():
    lib = _ffi.load()
    assert lib.imt_status_string(_ffi.ERR_EMPTY) == b"Cannot create Merkle Tree with no leaves"  # utils.rs:25
    assert lib.imt_status_string(_ffi.ERR_ODD) == b"Leaves must be even"  # utils.rs:35


def test_null_handles_do_not_crash():
    lib = _ffi.load()
    assert lib.imt_tree_num_leaves(None) == 0
    assert lib.imt_tree_depth(None) == 0
    assert lib.imt_tree_root(None, None) == _ffi.ERR_INVALID_ARG
    assert lib.imt_poseidon_hash2(None, None, 0, None) == _ffi.ERR_INVALID_ARG
    lib.imt_tree_destroy(None)
    lib.imt_ctx_destroy(None)


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(imt_b200.ImtError):
        imt_b200.Engine(0)


def test_product_never_references_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inl")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "poseidon_ref" not in text and "libimt_oracle" not in text, f

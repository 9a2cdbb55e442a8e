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


def test_checkpoint_header_is_readable_without_a_device(tmp_path):
    """imt_checkpoint_read_info touches no device: a hand-made file in the documented layout (include/imt_b200.h, checkpoints)
    parses; a wrong magic, a truncated file and a zero leaf count are refused."""
    import ctypes
    import struct
    lib = _ffi.load()
    n = 4
    leaves = bytes(range(96)) * n
    root = bytes(range(100, 132))
    good = b"IMTB200\0" + struct.pack("<6I", 1, 3, 2, 8, 57, 2) + struct.pack("<Q", n) + bytes(24) + leaves + root
    assert len(good) == 64 + 96 * n + 32
    p = tmp_path / "t.imt"
    p.write_bytes(good)
    info = _ffi.CheckpointInfo()
    assert lib.imt_checkpoint_read_info(os.fsencode(str(p)), ctypes.byref(info)) == _ffi.OK
    assert (info.num_leaves, info.version, info.t, info.rate, info.r_f, info.r_p, info.depth) == (n, 1, 3, 2, 8, 57, 2)
    assert bytes(info.root) == root
    for bad in (b"IMTB201\0" + good[8:], good[:-5], good[:32] + struct.pack("<Q", 0) + good[40:]):
        p.write_bytes(bad)
        assert lib.imt_checkpoint_read_info(os.fsencode(str(p)), ctypes.byref(info)) == _ffi.ERR_INVALID_ARG
    assert lib.imt_checkpoint_read_info(os.fsencode(str(tmp_path / "missing")), ctypes.byref(info)) == _ffi.ERR_INVALID_ARG


def test_multi_gpu_entry_points_reject_bad_arguments_without_a_device():
    import ctypes
    lib = _ffi.load()
    assert lib.imt_comm_create(None, 0, 2, None) == _ffi.ERR_INVALID_ARG
    assert lib.imt_tree_exchange_roots(None) == _ffi.ERR_INVALID_ARG
    assert lib.imt_multi_size(None) == 0 and lib.imt_mtree_num_leaves(None) == 0 and lib.imt_mtree_depth(None) == 0
    h = ctypes.c_void_p()
    devs = (ctypes.c_int * 3)(0, 1, 2)
    assert lib.imt_multi_create(devs, 3, 0, ctypes.byref(h)) == _ffi.ERR_INVALID_ARG      # not a power of two
    assert lib.imt_multi_create(None, 2, 0, ctypes.byref(h)) == _ffi.ERR_INVALID_ARG
    assert lib.imt_insert_trace_hashes(24) == 99
    lib.imt_multi_destroy(None)
    lib.imt_mtree_destroy(None)
    buf = (ctypes.c_uint8 * 128)()
    st = lib.imt_comm_unique_id(buf)       # NCCL is bound at run time: OK where libnccl.so.2 exists (this image), IMT_ERR_CUDA where not
    assert st in (_ffi.OK, _ffi.ERR_CUDA)
    if st == _ffi.OK:
        assert any(buf)

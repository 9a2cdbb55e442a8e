"""CPU-only: the any-width Poseidon instances. Pins the oracle's parameter generation and permutation on the
PUBLISHED Poseidon test vectors (the paper's reference implementation, files poseidonperm_x5_254_3 and _5: the
permutation of [0, 1, 2] / [0, 1, 2, 3, 4] over the BN254 scalar field with R_F = 8 and R_P = 57 / 60) — an anchor that
is independent of the reference's own known-answer (indexed_merkle_tree.rs:247-251) — and checks the library's HOST
parameter derivation (imt_spec_params_host: no device work) against the oracle's for every supported width."""
import ctypes
import json
import os

import numpy as np
import pytest

import imt_b200
from imt_b200 import _ffi
import poseidon_ref as R
import oracle as O

P = R.P
# Poseidon reference implementation, test vectors for x^5, 254-bit prime (BN254 Fr): input = [0, 1, ..., t-1]
KAT_T3 = [0x115cc0f5e7d690413df64c6b9662e9cf2a3617f2743245519e19607a4417189a,
          0x0fca49b798923ab0239de1c9e7a4a9a2210312b6a2f616d18b5a87f9b628ae29,
          0x0e7ae82e40091e63cbd4f16a6d16310b3729d4b6e138fcf54110e2867045a30c]
KAT_T5 = [0x299c867db6c1fdd79dcefa40e4510b9837e60ebb1ce0663dbaa525df65250465,
          0x1148aaef609aa338b27dafd89bb98862d8bb2b429aceac47d86206154ffe053d,
          0x24febb87fed7462e23f6665ff9a0111f4044c38ee1672c1ac6b0637d34f24907,
          0x0eb08f6d809668a981c186beaf6110060707059576406b248e5d9cf6e78b3d3e,
          0x07748bc6877c9b82c8b98666ee9d0626ec7f5be4205f79ee8528ef1c4a376fc7]
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))["spec"]
INSTANCES = [(2, 8, 56), (3, 8, 57), (4, 8, 56), (5, 8, 60), (3, 6, 10), (4, 2, 0)]


def test_published_permutation_vectors_pin_the_python_oracle():
    assert R.permute_naive([0, 1, 2]) == KAT_T3
    assert R.permute([0, 1, 2]) == KAT_T3                       # optimized schedule
    sp5 = R.Spec(8, 60, 5)
    assert R.permute_naive([0, 1, 2, 3, 4], sp5) == KAT_T5
    assert R.permute([0, 1, 2, 3, 4], sp5) == KAT_T5


def test_golden_fixture_holds_the_published_vectors_and_the_oracle_digests():
    assert [int(x, 16) for x in GOLD["published_perm_x5_254_3_input_0_1_2"]] == KAT_T3
    assert [int(x, 16) for x in GOLD["published_perm_x5_254_5_input_0_to_4"]] == KAT_T5
    for key, digests in GOLD["hashes"].items():
        t, r_f, r_p = (int(x) for x in key.split(","))
        sp = R.Spec(r_f, r_p, t)
        for k, d in digests.items():
            assert R.hash_n(list(range(1, int(k) + 1)), sp) == int(d)
    assert int(GOLD["hashes"]["3,8,57"]["3"]) == R.hash3(1, 2, 3) and int(GOLD["hashes"]["3,8,57"]["2"]) == R.hash2(1, 2)


def test_published_permutation_vector_pins_the_c_oracle():
    out = O.permute(O.fes([0, 1, 2]), naive=False)
    assert O.to_ints(out) == KAT_T3
    assert O.to_ints(O.permute(O.fes([0, 1, 2]), naive=True)) == KAT_T3


@pytest.mark.parametrize("t,r_f,r_p", INSTANCES)
def test_naive_and_optimized_schedules_agree_for_every_width(t, r_f, r_p):
    sp = R.Spec(r_f, r_p, t)
    rng = np.random.default_rng(t * 1000 + r_p)
    for _ in range(3):
        s = [int.from_bytes(rng.bytes(32), "little") % P for _ in range(t)]
        assert R.permute(s, sp) == R.permute_naive(s, sp)


def _host_params(t, r_f, r_p):
    lib = _ffi.load()
    cnt = ctypes.c_size_t()
    assert lib.imt_spec_params_host(t, t - 1, r_f, r_p, None, 0, ctypes.byref(cnt)) == _ffi.OK
    buf = np.zeros((cnt.value, 4), np.uint64)
    assert lib.imt_spec_params_host(t, t - 1, r_f, r_p, ctypes.c_void_p(buf.ctypes.data), cnt.value, None) == _ffi.OK
    rinv = pow(1 << 256, -1, P)
    return [imt_b200.fe_to_int(r) * rinv % P for r in buf]     # Montgomery -> integers


@pytest.mark.parametrize("t,r_f,r_p", INSTANCES)
def test_host_parameter_derivation_matches_the_oracle(t, r_f, r_p):
    got = _host_params(t, r_f, r_p)
    sp = R.Spec(r_f, r_p, t)
    half = r_f // 2
    want = [1 << 64, 1] + list(sp.start[0])
    full = [sp.start[i + 1] for i in range(half)] + [sp.end[i] for i in range(half - 1)] + [[0] * t]
    for row in full:
        want += list(row)
    for row in sp.mds:
        want += list(row)
    for row in sp.pre_sparse:
        want += list(row)
    for k in range(r_p):
        row, col = sp.sparse[k]
        want += [sp.partial[k]] + list(row) + list(col)
    assert len(got) == len(want) == 2 + t + r_f * t + 2 * t * t + r_p * 2 * t
    assert got == want


def test_unsupported_instances_are_rejected():
    lib = _ffi.load()
    cnt = ctypes.c_size_t()
    for t, rate, r_f, r_p in [(1, 0, 8, 57), (6, 5, 8, 57), (3, 1, 8, 57), (3, 2, 7, 57), (3, 2, 0, 57), (3, 2, 8, 300)]:
        assert lib.imt_spec_params_host(t, rate, r_f, r_p, None, 0, ctypes.byref(cnt)) == _ffi.ERR_INVALID_ARG
    h = ctypes.c_void_p()
    assert lib.imt_ctx_create_spec(0, 0, 6, 5, 8, 57, ctypes.byref(h)) == _ffi.ERR_INVALID_ARG
    assert lib.imt_ctx_spec(None, None, None, None, None, None) == _ffi.ERR_INVALID_ARG


def test_sponge_generalisation_reduces_to_the_reference_instance():
    assert R.hash_n([0, 0, 0]) == R.KAT_H3_ZERO               # indexed_merkle_tree.rs:247-251
    for inputs in ([1, 2], [1, 2, 3]):
        d, st = R.hash_trace_n(inputs)
        d2, st2 = R.hash_trace(inputs)
        assert (d, st) == (d2, st2)
    sp = R.Spec(8, 56, 4)
    h = R.Poseidon(sp=sp)
    h.update([5])
    h.update([6, 7, 8])                                         # buffered update: same as one update of 4 elements
    assert h.squeeze_and_reset() == R.hash_n([5, 6, 7, 8], sp)

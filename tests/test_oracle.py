"""Pins the oracle: C restatement (oracle/imt_oracle.c) vs the reference's known-answer, vs the independent big-int
restatement (oracle/poseidon_ref.py), and vs the committed golden fixtures. CPU only."""
import json
import os
import random

import numpy as np
import pytest

import oracle as O
import poseidon_ref as R

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


def test_reference_known_answer():
    # /root/reference/src/indexed_merkle_tree.rs:247-251 — the only numeric vector the reference holds
    kat = 1960587138944869480785025106734196872454309951825657414575195034687326603497
    assert O.to_int(O.hash3(O.fes([0, 0, 0]))[0]) == kat
    assert R.hash3(0, 0, 0) == kat
    assert int(GOLD["kat_h3_zero"]) == kat


def test_modulus_literal():
    # indexed_merkle_tree.rs:382-385
    assert R.P == 21888242871839275222246405745257275088548364400416034343698204186575808495617 == O.P


def test_constants_agree_between_oracles():
    sp = R.spec()
    flat = lambda rows: [x for r in rows for x in (r if isinstance(r, (list, tuple)) else [r])]  # noqa: E731
    assert O.to_ints(O.constants(0)) == flat(sp.round_constants)
    assert O.to_ints(O.constants(1)) == flat(sp.mds)
    assert O.to_ints(O.constants(2)) == flat(sp.start)
    assert O.to_ints(O.constants(3)) == flat(sp.partial)
    assert O.to_ints(O.constants(4)) == flat(sp.end)
    assert O.to_ints(O.constants(5)) == flat(sp.pre_sparse)
    assert O.to_ints(O.constants(6)) == flat([s[0] for s in sp.sparse])
    assert O.to_ints(O.constants(7)) == flat([s[1] for s in sp.sparse])
    # the well-known Poseidon-128 n=254 t=3 alpha=5 parameter set starts like this
    assert hex(sp.round_constants[0][0]) == "0xee9a592ba9a9518d05986d656f40c2114c4993c11bb29938d21d47304cd8e6e"
    assert hex(sp.mds[0][0]) == "0x109b7f411ba0e4c9b2b70caf5c36a7b194be7c11ad24378bfedb68592ba8118b"


def test_naive_and_optimized_schedules_agree():
    rng = random.Random(5)
    for _ in range(25):
        s = [rng.randrange(R.P) for _ in range(3)]
        a = O.to_ints(O.permute(O.fes(s)))
        assert a == O.to_ints(O.permute(O.fes(s), naive=True)) == R.permute(s) == R.permute_naive(s)


def test_hashes_and_traces_agree():
    rng = random.Random(6)
    for _ in range(10):
        x = [rng.randrange(R.P) for _ in range(3)]
        assert O.to_int(O.hash3(O.fes(x))[0]) == R.hash3(*x)
        assert O.to_int(O.hash2(O.fes(x[:2]))[0]) == R.hash2(*x[:2])
        for ar in (2, 3):
            dg, st = O.hash_trace(O.fes(x[:ar]))
            rd, rt = R.hash_trace(x[:ar])
            assert O.to_int(dg) == rd
            assert O.to_ints(st.reshape(-1, 4)) == [v for s_ in rt for v in s_]


def test_golden_small_vectors():
    assert R.hash2(0, 0) == int(GOLD["h2_0_0"])
    assert R.hash2(1, 2) == int(GOLD["h2_1_2"])
    assert R.hash3(1, 2, 3) == int(GOLD["h3_1_2_3"])
    assert O.to_int(O.InsertState(np.zeros((8, 3, 4), np.uint64)).root()) == int(GOLD["empty_depth3_root"])


def test_tree_paths_and_errors():
    rng = random.Random(8)
    leaves = [rng.randrange(R.P) for _ in range(16)]
    t = O.tree_build(O.fes(leaves))
    rt = R.IndexedMerkleTree(leaves)
    lv = O.levels(t, 16)
    assert [O.to_ints(l) for l in lv] == rt.tree
    for i in range(16):
        sib, hel = O.get_proof(t, 16, i)
        ps, ph = rt.get_proof(i)
        assert O.to_ints(sib) == ps and list(hel) == ph
        assert O.verify_proof(O.fe(leaves[i]), i, t[-1], sib)
        assert not O.verify_proof(O.fe(leaves[i] ^ 1), i, t[-1], sib)
    with pytest.raises(ValueError, match="Cannot create Merkle Tree with no leaves"):
        O.tree_build(np.zeros((0, 4), np.uint64))
    with pytest.raises(ValueError, match="Leaves must be even"):
        O.tree_build(O.fes([1, 2, 3]))
    with pytest.raises(ValueError, match="power of two"):
        O.tree_build(O.fes([1, 2, 3, 4, 5, 6]))
    one = O.tree_build(O.fes([7]))
    assert O.to_int(one[0]) == 7  # single leaf: root = leaf (utils.rs:27-33)
    with pytest.raises(IndexError):
        O.get_proof(t, 16, 16)


def test_insert_scenario_of_the_reference_test():
    # src/indexed_merkle_tree.rs:679-803 (deterministic): inserts 30,10,20,5,50,35 into slots 1..6 of a depth-3 tree
    vals = GOLD["scenario_inserts"]
    full = O.InsertState(np.zeros((8, 3, 4), np.uint64))
    inc = O.InsertState(np.zeros((8, 3, 4), np.uint64))
    ref_rounds, ref_pre = R.insert_rounds(3, vals)
    for i, v in enumerate(vals):
        a = full.insert(O.fe(v), i + 1, incremental=False)
        b = inc.insert(O.fe(v), i + 1, incremental=True)
        for k in a:
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k
        r = ref_rounds[i]
        assert a["low_idx"] == r["low_idx"] == GOLD["scenario_low_idx"][i]
        assert O.to_int(a["new_root"]) == r["new_root"] == int(GOLD["scenario_roots"][i])
        assert O.to_int(a["old_root"]) == r["old_root"]
        assert O.to_ints(a["low_proof"]) == r["low_proof"] and list(a["low_helper"]) == r["low_helper"]
        assert O.to_ints(a["new_proof"]) == r["new_proof"] and list(a["new_helper"]) == r["new_helper"]
        assert O.to_ints(a["low_leaf"]) == r["low_leaf"] and O.to_ints(a["new_leaf"]) == r["new_leaf"]
        assert a["is_largest"] == r["is_largest"]
    assert [O.to_ints(x) for x in full.pre] == ref_pre
    assert [[str(v) for v in leaf] for leaf in ref_pre] == GOLD["scenario_final_preimages"]


def test_incremental_inserts_match_rebuild_random():
    rng = random.Random(12)
    n = 32
    full = O.InsertState(np.zeros((n, 3, 4), np.uint64))
    inc = O.InsertState(np.zeros((n, 3, 4), np.uint64))
    vals = rng.sample(range(1, 10_000), 20) + [R.P - 1]
    for i, v in enumerate(vals):
        a = full.insert(O.fe(v), i + 1, incremental=False)
        b = inc.insert(O.fe(v), i + 1, incremental=True)
        for k in a:
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k
    assert np.array_equal(full.tree, inc.tree)


def test_low_leaf_scan_edge_cases():
    # duplicates and zero fall through exactly as update_idx_leaf does (indexed_merkle_tree.rs:639-658)
    st = O.InsertState(np.zeros((8, 3, 4), np.uint64))
    assert O.low_leaf(st.pre, O.fe(0)) == (0, True)  # first branch: head.next_val == 0 && i == 0, for ANY value
    for i, v in enumerate([30, 10, 20]):
        st.insert(O.fe(v), i + 1)
    assert O.low_leaf(st.pre, O.fe(15)) == (2, True)
    assert O.low_leaf(st.pre, O.fe(31)) == (1, True)
    assert O.low_leaf(st.pre, O.fe(5)) == (0, True)
    assert O.low_leaf(st.pre, O.fe(20)) == (4, True)   # duplicate: first empty slot (val 0 < 20, next_val == 0)
    assert O.low_leaf(st.pre, O.fe(0)) == (0, False)   # nothing satisfies val < 0


def test_golden_build_roots_small():
    import imt_b200
    from imt_b200 import synth
    for depth in ("3", "10"):
        n = 1 << int(depth)
        assert O.to_int(O.build_from_preimages(synth.random_preimages(n), 4)) == int(GOLD["build_roots"][depth]["random"])
        assert O.to_int(O.build_from_preimages(synth.indexed_preimages(n), 4)) == int(GOLD["build_roots"][depth]["indexed"])


def test_synthetic_generator_matches_oracle_definition():
    import imt_b200
    from imt_b200 import synth
    assert np.array_equal(synth.field_elements(4097, 0x494D54, 3), O.synth_fe(0x494D54, 3, 4097))
    pre = synth.indexed_preimages(256, 100, seed=9)
    ints = [O.to_ints(x) for x in pre]
    cur, count, last = 0, 0, -1
    while True:
        v, nv, ni = ints[cur]
        assert v > last
        last, count = v, count + 1
        if nv == 0:
            assert ni == 0
            break
        assert ints[ni][0] == nv
        cur = ni
    assert count == 100 and all(x == [0, 0, 0] for x in ints[100:])


def test_torch_synthetic_generators_match_numpy():
    """bench.py and the depth-24 GPU test synthesise leaves with torch on the device; same streams as the numpy versions"""
    import imt_b200
    from imt_b200 import synth
    a = synth.field_elements(1000, seed=77, first=5)
    b = synth.field_elements_torch(1000, seed=77, first=5, device="cpu").numpy().view(np.uint64)
    assert np.array_equal(a, b)
    for n, m in [(64, 40), (256, 256), (16, 1), (16, 2)]:
        assert np.array_equal(synth.indexed_preimages(n, m, seed=5), synth.indexed_preimages_torch(n, m, seed=5, device="cpu").numpy().view(np.uint64))

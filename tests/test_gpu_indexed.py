"""GPU parity tests of the indexed-leaf logic: low-leaf lookups, non-inclusion witnesses and
batched inserts through the C-ABI, against the oracle's restatement of the reference's test helpers
(update_idx_leaf IMT:632-660, insert orchestration IMT:710-741) — bit-exact. Run with `-m gpu`."""
import json
import os
import random

import numpy as np
import pytest

import imt_b200
from imt_b200 import synth, _ffi
import oracle as O

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
P = imt_b200.P
WIT = ["old_root", "low_idx", "low_leaf", "low_proof", "low_helper", "new_root", "new_leaf", "new_proof", "new_helper", "is_largest"]
GPU = ["old_roots", "low_idx", "low_leaves", "low_siblings", "low_helpers", "new_roots", "new_leaves", "new_siblings", "new_helpers", "is_largest"]


@pytest.fixture(scope="module")
def eng():
    e = imt_b200.Engine(0, "canonical")
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng_mont():
    e = imt_b200.Engine(0, "montgomery")
    yield e
    e.close()


def to_mont(a):
    return O.fes([x * (1 << 256) % P for x in O.to_ints(a)]).reshape(np.asarray(a).shape)


def from_mont(a):
    rinv = pow(1 << 256, -1, P)
    return O.fes([x * rinv % P for x in O.to_ints(a)]).reshape(np.asarray(a).shape)


def assert_witness_equal(got, k, want):
    for g, w in zip(GPU, WIT):
        a, b = got[g][k], want[w]
        assert np.array_equal(np.asarray(a), np.asarray(b)), f"insert {k}: {g} differs"


def oracle_fold_nodes(w, k, new_idx, dec=lambda a: a):
    """chain values of the four folds of insert k (IMT:196-204, 277-294, 305-313), recomputed by the oracle from the witnesses alone:
    (4, depth, 4) canonical — level 0 is the leaf hash, level l + 1 = H2 of level l with the path's sibling"""
    low, new = dec(w["low_leaves"][k]), dec(w["new_leaves"][k])
    new_low = np.stack([low[0], new[0], O.fe(new_idx)])
    lsib, nsib = dec(w["low_siblings"][k]), dec(w["new_siblings"][k])
    depth = lsib.shape[0]
    out = np.zeros((4, depth, 4), np.uint64)
    for f, (leaf, idx, sib) in enumerate(((low, int(w["low_idx"][k]), lsib), (new_low, int(w["low_idx"][k]), lsib),
                                          (np.zeros((3, 4), np.uint64), new_idx, nsib), (new, new_idx, nsib))):
        h = O.hash3(leaf.reshape(1, 3, 4), 1)[0]
        for lvl in range(depth):
            out[f, lvl] = h
            pair = np.stack([h, sib[lvl]]) if idx % 2 == 0 else np.stack([sib[lvl], h])
            h = O.hash2(pair.reshape(1, 2, 4), 1)[0]
            idx //= 2
    return out


def test_reference_scenario_as_one_batch(eng):
    """test_insert_leaf_multiple_round (IMT:679-741): 6 inserts into the empty depth-3 tree, in ONE device batch.
    Roots / low-leaf indices are the committed fixtures; every other witness field is compared with the oracle running
    the reference's own re-hash + rebuild per round."""
    pre = np.zeros((8, 3, 4), np.uint64)
    tree = eng.build_from_leaves(pre)
    assert imt_b200.fe_to_int(tree.root()) == int(GOLD["empty_depth3_root"])
    assert tree.occupied == 1
    vals = O.fes(GOLD["scenario_inserts"])
    got = tree.insert_batch(vals, 1)
    assert [int(v) for v in got["low_idx"]] == GOLD["scenario_low_idx"]
    assert [str(v) for v in O.to_ints(got["new_roots"])] == GOLD["scenario_roots"]
    st = O.InsertState(pre)
    for k in range(6):
        assert_witness_equal(got, k, st.insert(vals[k], k + 1, incremental=False))
    final = tree.preimages(8)
    assert [[str(v) for v in O.to_ints(leaf)] for leaf in final] == GOLD["scenario_final_preimages"]
    assert np.array_equal(final, st.pre)
    assert np.array_equal(tree.root(), st.root()) and tree.occupied == 7
    for lvl in range(4):
        assert np.array_equal(tree.level(lvl, 8 >> lvl), O.levels(st.tree, 8)[lvl])


def test_reference_scenario_one_insert_at_a_time(eng, eng_mont):
    for e, enc, dec in ((eng, lambda a: a, lambda a: a), (eng_mont, to_mont, from_mont)):
        tree = e.build_from_leaves(enc(np.zeros((8, 3, 4), np.uint64)))
        st = O.InsertState(np.zeros((8, 3, 4), np.uint64))
        for k, v in enumerate(GOLD["scenario_inserts"]):
            got = tree.insert_batch(enc(O.fes([v])))
            want = st.insert(O.fe(v), k + 1, incremental=False)
            for g, w in zip(GPU, WIT):
                a = got[g][0]
                if g in ("old_roots", "low_leaves", "low_siblings", "new_roots", "new_leaves", "new_siblings"):
                    a = dec(a)
                assert np.array_equal(np.asarray(a), np.asarray(want[w])), (k, g)
            assert str(O.to_int(dec(tree.root()))) == GOLD["scenario_roots"][k]


def _queries(pre, m, rng, count):
    vals = [O.to_int(pre[i, 0]) for i in range(1, m)]
    qs = [0, 1, P - 1, P - 2]
    if vals:
        qs += [min(vals) - 1, min(vals) + 1, max(vals) - 1, max(vals) + 1 if max(vals) + 1 < P else 1]
        qs += rng.sample(vals, min(len(vals), 12))                      # present values: the scan falls through
    qs += [rng.randrange(P) for _ in range(count)]
    return O.fes(qs)


@pytest.mark.parametrize("n,m", [(256, 200), (256, 256), (64, 1), (64, 2), (2, 1), (2, 2)])
def test_low_leaf_lookup_matches_the_linear_scan(eng, n, m):
    rng = random.Random(n * 1000 + m)
    pre = synth.indexed_preimages(n, m, seed=n + m)
    tree = eng.build_from_leaves(pre)
    assert tree.occupied == m
    qs = _queries(pre, m, rng, 150)
    low, matched = tree.low_leaf_lookup(qs)
    for k in range(len(qs)):
        wl, wm = O.low_leaf(pre, qs[k])
        assert (int(low[k]), bool(matched[k])) == (wl, wm), (k, O.to_int(qs[k]))


def test_low_leaf_lookup_montgomery_context(eng_mont):
    n, m = 128, 100
    pre = synth.indexed_preimages(n, m, seed=9)
    tree = eng_mont.build_from_leaves(to_mont(pre))
    qs = _queries(pre, m, random.Random(4), 60)
    low, matched = tree.low_leaf_lookup(to_mont(qs))
    for k in range(len(qs)):
        assert (int(low[k]), bool(matched[k])) == O.low_leaf(pre, qs[k])


def test_non_inclusion_witnesses(eng):
    """Everything verify_non_inclusion loads (IMT:127-137): low leaf, its path, is_largest — and the chip's own
    predicates hold on them (IMT:180-191, 226-228)."""
    n, m = 1024, 700
    pre = synth.indexed_preimages(n, m, seed=31)
    tree = eng.build_from_leaves(pre)
    top = max(O.to_int(pre[i, 0]) for i in range(m))
    qs = O.fes([random.Random(8).randrange(1, P) for _ in range(300)] + [top + 1, P - 1])
    o = tree.non_inclusion_paths(qs)
    assert o["matched"].all()
    sib, hel = tree.get_proofs(o["low_idx"])
    assert np.array_equal(o["siblings"], sib) and np.array_equal(o["helpers"], hel)
    assert np.array_equal(o["low_leaves"], pre[o["low_idx"].astype(np.int64)])
    for k in range(len(qs)):
        wl, wm = O.low_leaf(pre, qs[k])
        assert int(o["low_idx"][k]) == wl and wm
        val, nxt, v = O.to_int(o["low_leaves"][k, 0]), O.to_int(o["low_leaves"][k, 1]), O.to_int(qs[k])
        assert val < v and (v < nxt if not o["is_largest"][k] else nxt == 0)
    assert o["is_largest"][-1] == 1 and o["is_largest"][-2] == 1
    ok = eng.verify_proofs(eng.hash3(o["low_leaves"]), o["low_idx"], tree.root(), o["siblings"])
    assert ok.all()


def test_non_inclusion_witness_trace_one_call(eng, eng_mont):
    """imt_non_inclusion_witness_trace: the whole witness of verify_non_inclusion (IMT:127-229) in one call — the values of
    imt_non_inclusion_paths, the hi/lo limbs (IMT:143-172, 206-222) and the Poseidon states of H3(low leaf) (IMT:193-194) and of its
    fold up the path (IMT:196-204), every state against the oracle's trace; absent, present (unmatched) and zero values; both formats."""
    depth, m = 7, 90
    n = 1 << depth
    pre = synth.indexed_preimages(n, m, seed=77)
    vals = np.concatenate([synth.field_elements(40, seed=78), pre[3:6, 0], np.zeros((1, 4), np.uint64)])   # absent, present, zero
    for e, enc, dec in ((eng, lambda a: a, lambda a: a), (eng_mont, to_mont, from_mont)):
        tree = e.build_from_leaves(enc(pre))
        o = tree.trace_non_inclusion(enc(vals))
        ref = tree.non_inclusion_paths(enc(vals))
        for k in ("low_idx", "matched", "low_leaves", "siblings", "helpers", "is_largest"):
            assert np.array_equal(o[k], ref[k]), k
        limbs, flags = e.non_inclusion_limbs(o["low_leaves"], enc(vals))
        assert np.array_equal(o["limbs"], limbs) and np.array_equal(o["limb_flags"], flags)
        for k in range(len(vals)):                                                 # the reference's literal scan, quirks included (IMT:632-660)
            assert (int(o["low_idx"][k]), bool(o["matched"][k])) == O.low_leaf(pre, vals[k]), k
        assert o["matched"][:40].all() and o["limb_flags"][:40, 2].all() and not o["matched"][-1] and not o["limb_flags"][-1, 2]
        root = dec(tree.root())
        for k in list(range(0, 40, 7)) + [40, 43]:
            wl = int(o["low_idx"][k])
            h, ws = O.hash_trace(dec(o["low_leaves"][k]))
            assert np.array_equal(dec(o["leaf_hash"][k]), ws), k
            idx, sib = wl, dec(o["siblings"][k])
            for lvl in range(depth):
                pair = np.stack([h, sib[lvl]]) if idx % 2 == 0 else np.stack([sib[lvl], h])
                h, ws = O.hash_trace(pair)
                assert np.array_equal(dec(o["path"][k, lvl]), ws), (k, lvl)
                idx //= 2
            assert np.array_equal(h, root)
        # values only (no states), and the same bytes as the tree-path trace + the leaf-hash trace composed by hand
        lite = tree.trace_non_inclusion(enc(vals), states=False)
        assert "states" not in lite and np.array_equal(lite["limbs"], o["limbs"]) and np.array_equal(lite["siblings"], o["siblings"])
        assert np.array_equal(o["path"], tree.trace_proofs(o["low_idx"]))
    with pytest.raises(imt_b200.ImtError) as err:                                  # a value >= p
        tree.trace_non_inclusion(np.full((1, 4), 0xFFFFFFFFFFFFFFFF, np.uint64))
    assert err.value.status == _ffi.ERR_NON_CANONICAL


def test_non_inclusion_witness_trace_depth20_chunked_pipeline(eng_mont):
    """the host call streams (1 + depth) x 12 672 B per query out in chunks behind the hashing: 20 000 queries at depth 20 (5.3 GB, three
    chunks) equal the device-resident call, and sampled queries equal the oracle's trace"""
    import torch
    depth, q = 20, 20000
    n = 1 << depth
    dev = torch.device("cuda", 0)
    e = eng_mont
    d_pre = synth.indexed_preimages_torch(n, n - 5, device=dev)
    d_pre_m = torch.empty_like(d_pre)
    e.convert_dev(d_pre, 3 * n, d_pre_m, to_montgomery=True)
    tree = e.build_from_leaves_dev(d_pre_m, n)
    d_vals = synth.field_elements_torch(q, seed=4321, device=dev)
    d_vals_m = torch.empty_like(d_vals)
    e.convert_dev(d_vals, q, d_vals_m, to_montgomery=True)
    vals_m = d_vals_m.cpu().numpy().view(np.uint64)
    o = tree.trace_non_inclusion(vals_m)
    d_low = torch.empty(q, dtype=torch.int64, device=dev)
    d_leaves = torch.empty((q, 3, 4), dtype=torch.int64, device=dev)
    d_states = torch.empty((q, 1 + depth, 132, 3, 4), dtype=torch.int64, device=dev)
    launches = e.launches
    tree.trace_non_inclusion_dev(d_vals_m, q, d_low, d_leaves, d_states)
    assert e.launches - launches <= 4                                              # lookup, leaf gather, leaf-hash trace, path trace
    assert np.array_equal(d_low.cpu().numpy().astype(np.uint64), o["low_idx"])
    assert np.array_equal(d_leaves.cpu().numpy().view(np.uint64), o["low_leaves"])
    for lo in range(0, q, 4000):                                                   # compare in slices: 5.3 GB each side
        assert np.array_equal(d_states[lo:lo + 4000].cpu().numpy().view(np.uint64), o["states"][lo:lo + 4000]), lo
    root = from_mont(tree.root())
    for k in (0, 8191, 8192, q - 1):                                               # chunk edges
        h, ws = O.hash_trace(from_mont(o["low_leaves"][k]))
        assert np.array_equal(from_mont(o["leaf_hash"][k].reshape(-1, 4)).reshape(132, 3, 4), ws), k
        idx, sib = int(o["low_idx"][k]), from_mont(o["siblings"][k])
        for lvl in range(depth):
            pair = np.stack([h, sib[lvl]]) if idx % 2 == 0 else np.stack([sib[lvl], h])
            if lvl in (0, 9, depth - 1):
                h, ws = O.hash_trace(pair)
                assert np.array_equal(from_mont(o["path"][k, lvl].reshape(-1, 4)).reshape(132, 3, 4), ws), (k, lvl)
            else:
                h = O.hash2(pair.reshape(1, 2, 4), 1)[0]
            idx //= 2
        assert np.array_equal(h, root)


@pytest.mark.parametrize("depth,m,b,chunk", [(6, 10, 40, None), (10, 512, 300, 128), (13, 3000, 4200, None), (13, 3000, 4200, 4096)])
def test_insert_batch_matches_sequential_oracle(eng, depth, m, b, chunk, monkeypatch):
    """b inserts in one call against the oracle advancing one insert at a time; then the device state (preimages, every
    level, index) against the oracle's final state. `chunk` forces the internal chunking (default 16384 inserts) so that
    the batch crosses chunk boundaries."""
    if chunk:
        monkeypatch.setenv("IMT_INSERT_CHUNK", str(chunk))
    n = 1 << depth
    pre = synth.indexed_preimages(n, m, seed=depth)
    tree = eng.build_from_leaves(pre)
    vals = synth.field_elements(b, seed=1000 + depth)
    got = tree.insert_batch(vals, m)
    st = O.InsertState(pre, threads=8)
    step = 1 if b <= 400 else 7
    for k in range(b):
        want = st.insert(vals[k], m + k, incremental=True)
        if k % step == 0 or k >= b - 3:
            assert_witness_equal(got, k, want)
            if k % (5 * step) == 0 or k >= b - 3:   # the chain values of the four folds (imt_insert_witness::fold_nodes), across chunks
                assert np.array_equal(got["fold_nodes"][k], oracle_fold_nodes(got, k, m + k)), k
        else:
            assert np.array_equal(got["new_roots"][k], want["new_root"]) and int(got["low_idx"][k]) == want["low_idx"]
    assert np.array_equal(tree.preimages(n), st.pre)
    want_levels = O.levels(st.tree, n)
    for lvl in range(depth + 1):
        assert np.array_equal(tree.level(lvl, n >> lvl), want_levels[lvl]), lvl
    assert tree.occupied == m + b
    # the merged index answers like a linear scan over the new preimages
    qs = _queries(st.pre, m + b, random.Random(depth), 80)
    low, matched = tree.low_leaf_lookup(qs)
    for k in range(len(qs)):
        assert (int(low[k]), bool(matched[k])) == O.low_leaf(st.pre, qs[k])
    # idempotence: a from-scratch build of the updated preimages reproduces the incrementally maintained root
    assert np.array_equal(eng.build_from_leaves(st.pre).root(), tree.root())


def test_insert_batch_matches_full_rebuild_oracle(eng):
    """small enough for the reference's literal O(n)-per-insert sequence (re-hash all + rebuild all, IMT:724-730)"""
    n, m, b = 32, 5, 20
    pre = synth.indexed_preimages(n, m, seed=77)
    tree = eng.build_from_leaves(pre)
    vals = synth.field_elements(b, seed=78)
    got = tree.insert_batch(vals)
    st = O.InsertState(pre)
    for k in range(b):
        assert_witness_equal(got, k, st.insert(vals[k], m + k, incremental=False))


def test_insert_leaf_witness_traces(eng, eng_mont):
    """the per-round Poseidon states of all (3 + 4 depth) hashes insert_leaf constrains (IMT:253-313), in its
    call order, against the oracle's trace of the same hashes; the four folds end in old / interim / interim / new root."""
    depth, m, b = 5, 7, 6
    n = 1 << depth
    pre = synth.indexed_preimages(n, m, seed=21)
    vals = synth.field_elements(b, seed=22)
    for e, enc, dec in ((eng, lambda a: a, lambda a: a), (eng_mont, to_mont, from_mont)):
        tree = e.build_from_leaves(enc(pre))
        w = tree.insert_batch(enc(vals), m)
        tr = e.trace_insert_witness(w, m)                                        # one launch: fold_nodes came with the batch
        loop = e.trace_insert_witness({k: v for k, v in w.items() if k != "fold_nodes"}, m)   # 1 + depth dependent launches
        for name in ("states", "limbs", "limb_flags", "old_root", "interim_root", "zero_leaf_root", "new_root", "new_low_leaf_preimage"):
            assert np.array_equal(tr[name], loop[name]), name
        for k in range(b):
            assert np.array_equal(dec(w["fold_nodes"][k]), oracle_fold_nodes(w, k, m + k, dec)), k
        bad = dict(w, fold_nodes=w["fold_nodes"].copy())                         # chain values that are not these witnesses'
        bad["fold_nodes"][2, 1, 3] = w["fold_nodes"][2, 1, 2]
        with pytest.raises(imt_b200.ImtError) as err:
            e.trace_insert_witness(bad, m)
        assert err.value.status == _ffi.ERR_INVALID_ARG and "fold_nodes" in str(err.value)
        bad["fold_nodes"][2, 1, 3] = w["fold_nodes"][2, 1, 3]
        bad["fold_nodes"][4, 2, 0] = w["fold_nodes"][4, 0, 0]                    # a fold that does not start from the empty leaf
        with pytest.raises(imt_b200.ImtError):
            e.trace_insert_witness(bad, m)
        assert np.array_equal(tr["old_root"], w["old_roots"]) and np.array_equal(tr["new_root"], w["new_roots"])
        assert np.array_equal(tr["zero_leaf_root"], tr["interim_root"])          # IMT:286-294 holds on GPU-made witnesses
        assert tr["limb_flags"][:, 2].all()                                      # every witness passes the chip's assertions
        for k in range(b):                                                       # hi/lo limb splits, IMT:143-172, 206-222
            low, new = O.to_ints(dec(w["low_leaves"][k])), O.to_ints(dec(w["new_leaves"][k]))
            want = [x for v in (new[0], low[1], low[0]) for x in (v >> 128, v & ((1 << 128) - 1))]
            assert O.to_ints(dec(tr["limbs"][k])) == want
            assert list(tr["limb_flags"][k][:2]) == [new[0] < low[1], low[0] < new[0]]
        for k in (0, 3, b - 1):
            low, new = dec(w["low_leaves"][k]), dec(w["new_leaves"][k])
            new_low = np.stack([low[0], new[0], O.fe(m + k)])
            assert np.array_equal(dec(tr["new_low_leaf_preimage"][k]), new_low)
            for name, leaf, path_name, idx, sib in (("low_leaf", low, "low_path", int(w["low_idx"][k]), dec(w["low_siblings"][k])),
                                                    ("new_low_leaf", new_low, "interim_path", int(w["low_idx"][k]), dec(w["low_siblings"][k])),
                                                    (None, np.zeros((3, 4), np.uint64), "zero_path", m + k, dec(w["new_siblings"][k])),
                                                    ("new_leaf", new, "new_path", m + k, dec(w["new_siblings"][k]))):
                h, ws = O.hash_trace(leaf)
                if name:
                    assert np.array_equal(dec(tr[name][k]), ws), (k, name)
                for lvl in range(depth):
                    pair = np.stack([h, sib[lvl]]) if idx % 2 == 0 else np.stack([sib[lvl], h])
                    h, ws = O.hash_trace(pair)
                    assert np.array_equal(dec(tr[path_name][k, lvl]), ws), (k, path_name, lvl)
                    idx //= 2


def test_non_inclusion_limbs_edge_values(eng, eng_mont):
    """imt_non_inclusion_limbs on hand-picked pairs: equal high limbs, equal values, the largest-leaf case (next_val = 0)
    and a pair that violates the chip's ordering assertion (IMT:180-191, 226-228)"""
    M = (1 << 128)
    cases = [  # (low.val, low.next_val, new value)      passes?
        (5, 9, 7, True),
        (5 * M + 1, 5 * M + 9, 5 * M + 3, True),       # same high limb: decided by the low limbs
        (3 * M + 7, 0, P - 1, True),                   # low leaf is the largest: next_val == 0
        (5, 9, 9, False),                              # new == next_val
        (5, 9, 5, False),                              # new == low.val
        (7 * M, 9 * M, 6 * M + (M - 1), False),        # new < low.val
        (2, P - 1, P - 2, True),
    ]
    low = O.fes([x for lv, nv, _, _ in cases for x in (lv, nv, 3)]).reshape(-1, 3, 4)
    new = O.fes([c[2] for c in cases])
    for e, enc, dec in ((eng, lambda a: a, lambda a: a), (eng_mont, to_mont, from_mont)):
        limbs, flags = e.non_inclusion_limbs(enc(low), enc(new))
        for k, (lv, nv, v, ok) in enumerate(cases):
            assert O.to_ints(dec(limbs[k])) == [v >> 128, v % M, nv >> 128, nv % M, lv >> 128, lv % M]
            assert list(flags[k]) == [v < nv, lv < v, ok]
    with pytest.raises(imt_b200.ImtError) as ex:
        bad = new.copy()
        bad[0] = O.fe(0)
        bad[0, 3] = np.uint64(0xFFFFFFFFFFFFFFFF)                                # >= p
        eng.non_inclusion_limbs(low, bad)
    assert ex.value.status == _ffi.ERR_NON_CANONICAL


def test_insert_batch_is_independent_of_the_chunking(eng, monkeypatch):
    """the same 9000 inserts with chunks of 16384 (one chunk), 4096 and 1000: identical witnesses and final state"""
    depth, m, b = 15, 20000, 9000
    n = 1 << depth
    pre = synth.indexed_preimages(n, m, seed=61)
    vals = synth.field_elements(b, seed=62)
    results = []
    for chunk in (None, 4096, 1000):
        if chunk:
            monkeypatch.setenv("IMT_INSERT_CHUNK", str(chunk))
        else:
            monkeypatch.delenv("IMT_INSERT_CHUNK", raising=False)
        tree = eng.build_from_leaves(pre)
        w = tree.insert_batch(vals, m)
        results.append((w, tree.root().copy(), tree.preimages(n)))
        tree.close()
    w0, r0, p0 = results[0]
    for w, r, p in results[1:]:
        assert np.array_equal(r, r0) and np.array_equal(p, p0)
        for k in w0:
            assert np.array_equal(np.asarray(w[k]), np.asarray(w0[k])), k
    st = O.InsertState(pre, threads=8)                       # and against the oracle, one insert at a time
    for k in range(b):
        want = st.insert(vals[k], m + k, incremental=True)
        if k % 251 == 0 or k == b - 1:
            assert_witness_equal(w0, k, want)
    assert np.array_equal(r0, st.root())


def test_insert_errors_leave_the_tree_untouched(eng):
    n, m = 16, 6
    pre = synth.indexed_preimages(n, m, seed=5)
    tree = eng.build_from_leaves(pre)
    root = tree.root()
    present = pre[3, 0]
    fresh = synth.field_elements(3, seed=99)
    for bad in (np.stack([fresh[0], present]), np.stack([fresh[0], fresh[1], fresh[0]]), np.stack([fresh[0], O.fe(0)])):
        with pytest.raises(imt_b200.ImtError) as e:
            tree.insert_batch(bad, m)
        assert e.value.status == _ffi.ERR_INVALID_ARG
    with pytest.raises(imt_b200.ImtError) as e:
        tree.insert_batch(fresh, m + 1)                                  # not the next free slot
    assert e.value.status == _ffi.ERR_INVALID_ARG
    with pytest.raises(imt_b200.ImtError) as e:
        tree.insert_batch(synth.field_elements(n - m + 1, seed=3), m)
    assert e.value.status == _ffi.ERR_TREE_FULL
    assert np.array_equal(tree.root(), root) and np.array_equal(tree.preimages(n), pre) and tree.occupied == m
    # a tree built from hashes has no preimages to index
    with pytest.raises(imt_b200.ImtError) as e:
        eng.build_from_hashes(O.hash3(pre, 4)).low_leaf_lookup(fresh)
    assert e.value.status == _ffi.ERR_INVALID_ARG
    # preimages that are not a sorted linked list
    broken = pre.copy()
    broken[2, 1] = fresh[2]
    with pytest.raises(imt_b200.ImtError) as e:
        eng.build_from_leaves(broken).low_leaf_lookup(fresh)
    assert e.value.status == _ffi.ERR_NOT_WELL_FORMED
    hole = pre.copy()
    hole[m + 2] = pre[2]
    with pytest.raises(imt_b200.ImtError) as e:
        eng.build_from_leaves(hole).low_leaf_lookup(fresh)
    assert e.value.status == _ffi.ERR_NOT_WELL_FORMED


def test_config5_one_million_lookups_then_4096_inserts(eng):
    """BASELINE config[4] at depth 20: 1M low-leaf lookups + non-inclusion paths, then a batched insert of 4096
    leaves with per-insert roots — checked through size-independent properties and spot checks against the oracle."""
    depth, b = 20, 4096
    n = 1 << depth
    m = n - b
    pre = synth.indexed_preimages(n, m)
    tree = eng.build_from_leaves(pre)
    q = 1 << 20
    qs = synth.field_elements(q, seed=555)
    low, matched = tree.low_leaf_lookup(qs)
    assert matched.all()
    lo = low.astype(np.int64)

    def lt(a, b):  # a < b on (k, 4) little-endian words
        res = np.zeros(a.shape[0], bool)
        dec = np.zeros(a.shape[0], bool)
        for w in (3, 2, 1, 0):
            res |= ~dec & (a[:, w] < b[:, w])
            dec |= a[:, w] != b[:, w]
        return res

    nxt = pre[lo, 1]
    assert lt(pre[lo, 0], qs).all() and (lt(qs, nxt) | ~nxt.any(axis=1)).all()      # val < v < next_val (or next_val == 0)
    for k in random.Random(1).sample(range(q), 3):
        assert (int(low[k]), True) == O.low_leaf(pre, qs[k])
    part = tree.non_inclusion_paths(qs[:4096])
    assert np.array_equal(part["low_idx"], low[:4096])
    assert eng.verify_proofs(eng.hash3(part["low_leaves"]), part["low_idx"], tree.root(), part["siblings"]).all()

    vals = synth.field_elements(b, seed=556)
    old_root = tree.root()
    got = tree.insert_batch(vals, m)
    assert np.array_equal(got["old_roots"][0], old_root)
    assert np.array_equal(got["old_roots"][1:], got["new_roots"][:-1])              # the roots chain
    assert np.array_equal(got["new_roots"][-1], tree.root())
    # every witness verifies the way the chip checks it: low leaf under old_root (IMT:253-263), new leaf under new_root (IMT:305-313)
    idx_new = np.arange(m, m + b, dtype=np.uint64)
    assert eng.verify_proofs(eng.hash3(got["low_leaves"]), got["low_idx"], got["old_roots"], got["low_siblings"]).all()
    assert eng.verify_proofs(eng.hash3(got["new_leaves"]), idx_new, got["new_roots"], got["new_siblings"]).all()
    # ... and the empty leaf sits under the interim root the chip derives itself (IMT:277-294)
    low_after = got["low_leaves"].copy()
    low_after[:, 1] = vals
    low_after[:, 2] = 0
    low_after[:, 2, 0] = idx_new
    roots_i, _ = eng.trace_merkle_proofs(eng.hash3(low_after), got["low_idx"], got["low_siblings"], want_states=False)
    zero_leaf = np.broadcast_to(eng.hash3(np.zeros((1, 3, 4), np.uint64)), (b, 4))
    assert eng.verify_proofs(zero_leaf, idx_new, roots_i, got["new_siblings"]).all()
    # the first inserts against the sequential oracle, the final root against a from-scratch oracle build
    st = O.InsertState(pre, threads=O.max_threads())
    for k in range(3):
        assert_witness_equal(got, k, st.insert(vals[k], m + k, incremental=True))
    final = tree.preimages(n)
    assert np.array_equal(O.build_from_preimages(final, O.max_threads()), tree.root())


def test_checkpoint_round_trip(eng, tmp_path):
    n, m = 256, 100
    pre = synth.indexed_preimages(n, m, seed=12)
    tree = eng.build_from_leaves(pre)
    tree.insert_batch(synth.field_elements(20, seed=13))
    path = str(tmp_path / "tree.imt")
    tree.save(path)
    back = eng.load_tree(path)
    assert np.array_equal(back.root(), tree.root()) and np.array_equal(back.preimages(n), tree.preimages(n)) and back.occupied == m + 20
    raw = open(path, "rb").read()
    assert len(raw) == 64 + 96 * n + 32 and raw[:8] == b"IMTB200\0"
    assert raw[64:64 + 96] == tree.preimages(n)[0].tobytes()       # 3 x 32-byte little-endian canonical FE per leaf (utils.rs:12-17 order)
    assert raw[-32:] == tree.root().tobytes()
    info = eng.checkpoint_info(path)
    assert info["num_leaves"] == n and info["instance"] == (3, 2, 8, 57) and np.array_equal(info["root"], tree.root())
    # a Montgomery-format engine reads and writes the SAME bytes (the file is canonical whatever the context format)
    em = imt_b200.Engine(0, "montgomery")
    tm = em.load_tree(path)
    path2 = str(tmp_path / "tree_m.imt")
    tm.save(path2)
    assert open(path2, "rb").read() == raw
    em.close()
    bad = bytearray(raw)
    bad[64 + 96 * 5] ^= 1                                           # one bit of leaf 5's val
    open(path, "wb").write(bytes(bad))
    with pytest.raises(imt_b200.ImtError, match="corrupt"):
        eng.load_tree(path)
    bad = bytearray(raw)
    bad[64 + 31] = 0xFF                                             # leaf 0's val >= p
    open(path, "wb").write(bytes(bad))
    with pytest.raises(imt_b200.ImtError) as ei:
        eng.load_tree(path)
    assert ei.value.status == imt_b200._ffi.ERR_NON_CANONICAL
    open(path, "wb").write(raw[: 64 + 96 * 7])                      # truncated
    with pytest.raises(imt_b200.ImtError):
        eng.load_tree(path)
    e5 = imt_b200.Engine(0, "canonical", t=5, rate=4, r_f=8, r_p=60)
    open(path, "wb").write(raw)
    with pytest.raises(imt_b200.ImtError, match="another Poseidon instance"):
        e5.load_tree(path)
    e5.close()


def test_reference_test_insert_leaf_mirror(eng):
    """test_insert_leaf (IMT:360-596), native half, written with the reference's names: insert a (seeded) random 254-bit
    value into the empty depth-3 tree as the largest, then 42 between 0 and it — by hand with re-hash + rebuild as the
    reference does, and in one insert_batch; both must agree, and every verify_proof must hold."""
    from imt_b200 import Poseidon, IndexedMerkleTree, IndexedMerkleTreeLeaf as IMTLeaf
    native_hasher = Poseidon(8, 57, engine=eng)
    leaves = []
    for _ in range(8):                                                       # IMT:372-376
        native_hasher.update([0, 0, 0])
        leaves.append(native_hasher.squeeze_and_reset())
    tree = IndexedMerkleTree.new(native_hasher, list(leaves))
    a = random.Random(2024).getrandbits(254) % P                             # IMT:380-388, seeded
    old_root = tree.get_root()
    low_leaf_proof, low_helper = tree.get_proof(0)
    assert tree.verify_proof(leaves[0], 0, tree.get_root(), low_leaf_proof)  # IMT:397-400
    new_low_leaf = IMTLeaf(0, a, 1)
    native_hasher.update(new_low_leaf.as_tuple())
    leaves[0] = native_hasher.squeeze_and_reset()
    native_hasher.update([a, 0, 0])
    leaves[1] = native_hasher.squeeze_and_reset()
    tree = IndexedMerkleTree.new(native_hasher, list(leaves))                # IMT:417
    new_leaf_proof, new_helper = tree.get_proof(1)
    assert tree.verify_proof(leaves[1], 1, tree.get_root(), new_leaf_proof)  # IMT:420-423
    root1 = tree.get_root()
    # second insert: 42, low leaf = slot 0 {0, a, 1} (IMT:495-525)
    native_hasher.update([0, 42, 2])
    leaves[0] = native_hasher.squeeze_and_reset()
    native_hasher.update([42, a, 1])
    leaves[2] = native_hasher.squeeze_and_reset()
    tree = IndexedMerkleTree.new(native_hasher, list(leaves))
    proof2, _ = tree.get_proof(2)
    assert tree.verify_proof(leaves[2], 2, tree.get_root(), proof2)          # IMT:522-525
    root2 = tree.get_root()
    # the same two inserts as one device batch
    dev = eng.build_from_leaves(np.zeros((8, 3, 4), np.uint64))
    w = dev.insert_batch(O.fes([a, 42]), 1)
    assert O.to_ints(w["old_roots"]) == [old_root, root1] and O.to_ints(w["new_roots"]) == [root1, root2]
    assert [int(v) for v in w["low_idx"]] == [0, 0] and [int(v) for v in w["is_largest"]] == [1, 0]
    assert O.to_ints(w["low_leaves"][0]) == [0, 0, 0] and O.to_ints(w["low_leaves"][1]) == [0, a, 1]
    assert O.to_ints(w["new_leaves"][0]) == [a, 0, 0] and O.to_ints(w["new_leaves"][1]) == [42, a, 1]
    assert O.to_ints(w["low_siblings"][0]) == low_leaf_proof and O.to_ints(w["new_siblings"][0]) == new_leaf_proof
    assert O.to_ints(w["new_siblings"][1]) == proof2
    assert [int(h) for h in w["low_helpers"][0]] == low_helper and [int(h) for h in w["new_helpers"][0]] == new_helper


@pytest.mark.parametrize("pattern", ["ascending", "descending", "zigzag", "dense_chain"])
def test_insert_batch_adversarial_orders(eng, pattern):
    """Orders that stress the sequential semantics inside one batch: every insert's low leaf is the previous insert
    (ascending), always the head (descending), alternates between the two ends (zigzag), or 300 consecutive integers
    hanging off one existing leaf (dense_chain) — against the oracle advancing one insert at a time."""
    depth, m = 10, 3
    n = 1 << depth
    pre = synth.indexed_preimages(n, m, seed=3)
    b = 300
    base = sorted(O.to_int(pre[i, 0]) for i in range(1, m))
    if pattern == "ascending":
        vals = [base[-1] + 1 + 7 * k for k in range(b)]
    elif pattern == "descending":
        vals = [base[0] - 1 - 7 * k for k in range(b)]
    elif pattern == "zigzag":
        vals = [(base[-1] + 1 + k) if k % 2 == 0 else (base[0] - 1 - k) for k in range(b)]
    else:
        mid = (base[0] + base[1]) // 2
        order = list(range(b))
        random.Random(5).shuffle(order)
        vals = [mid + k for k in order]
    assert all(0 < v < P for v in vals) and len(set(vals)) == b
    v = O.fes(vals)
    tree = eng.build_from_leaves(pre)
    got = tree.insert_batch(v, m)
    st = O.InsertState(pre, threads=4)
    for k in range(b):
        assert_witness_equal(got, k, st.insert(v[k], m + k, incremental=True))
    assert np.array_equal(tree.preimages(n), st.pre) and np.array_equal(tree.root(), st.root())


def test_insert_batch_chunk_boundary_and_single_inserts(eng, monkeypatch):
    """4097 inserts = one full internal chunk + 1 (chunks forced to 4096); then single-insert calls; then the tree filled to
    the last slot."""
    monkeypatch.setenv("IMT_INSERT_CHUNK", "4096")
    depth = 13
    n = 1 << depth
    m = n - 4097 - 3
    pre = synth.indexed_preimages(n, m, seed=44)
    tree = eng.build_from_leaves(pre)
    st = O.InsertState(pre, threads=8)
    vals = synth.field_elements(4100, seed=45)
    got = tree.insert_batch(vals[:4097], m)
    for k in range(4097):
        want = st.insert(vals[k], m + k, incremental=True)
        if k in (0, 1, 4095, 4096) or k % 97 == 0:
            assert_witness_equal(got, k, want)
    assert np.array_equal(tree.root(), st.root())
    for k in range(4097, 4100):                                          # B = 1, three times, up to a completely full tree
        got = tree.insert_batch(vals[k:k + 1])
        assert_witness_equal(got, 0, st.insert(vals[k], m + k, incremental=True))
    assert tree.occupied == n and np.array_equal(tree.preimages(n), st.pre)
    with pytest.raises(imt_b200.ImtError) as e:
        tree.insert_batch(synth.field_elements(1, seed=46))
    assert e.value.status == _ffi.ERR_TREE_FULL
    low, matched = tree.low_leaf_lookup(np.stack([vals[7], synth.field_elements(1, seed=47)[0]]))   # present value, full tree
    assert (int(low[0]), bool(matched[0])) == O.low_leaf(st.pre, vals[7]) == (0, False)
    assert (int(low[1]), bool(matched[1])) == O.low_leaf(st.pre, synth.field_elements(1, seed=47)[0])


def test_insert_witness_trace_depth24_batch_4096_sampled_against_the_oracle(eng_mont):
    """imt_insert_witness_trace at the size BASELINE config 5 names: 4096 inserts into the depth-24 tree = 405 504 traced
    hashes (5.1 GB of states, kept on the device). Sampled (insert, slot) pairs are recomputed by the oracle from the
    witnesses alone: fold with O.hash_trace up to the sampled level, compare all 132 x 3 states."""
    import torch
    depth, b = 24, 4096
    n = 1 << depth
    m = n - 3 * b
    dev = torch.device("cuda", 0)
    d_pre = synth.indexed_preimages_torch(n, m, device=dev)                    # canonical integers
    e = eng_mont
    d_pre_m = torch.empty_like(d_pre)
    e.convert_dev(d_pre, 3 * n, d_pre_m, to_montgomery=True)
    del d_pre
    tree = e.build_from_leaves_dev(d_pre_m, n)
    vals = to_mont(synth.field_elements(b, seed=4242))
    w = tree.insert_batch(vals, m)
    dw = {k: torch.from_numpy(np.ascontiguousarray(w[k]).view(np.int64) if w[k].dtype == np.uint64 else w[k]).to(dev)
          for k in ("low_idx", "low_leaves", "low_siblings", "new_leaves", "new_siblings", "fold_nodes")}
    S = 3 + 4 * depth
    d_states = torch.empty((b, S, 132, 3, 4), dtype=torch.int64, device=dev)
    d_roots = torch.empty((b, 4, 4), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()                                                   # the engine runs on its own stream: inputs must be complete
    e.trace_insert_witness_dev({k: v for k, v in dw.items() if k != "fold_nodes"}, b, depth, m, d_states, d_roots)   # level loop
    d_states_1, d_roots_1 = torch.zeros_like(d_states), torch.zeros_like(d_roots)
    torch.cuda.synchronize()                                                   # the fills run on torch's stream, the library on its own
    launches = e.launches
    e.trace_insert_witness_dev(dw, b, depth, m, d_states_1, d_roots_1)         # one launch of 4 b depth independent traced hashes (+ leaves)
    assert e.launches - launches == 2
    assert torch.equal(d_states_1, d_states) and torch.equal(d_roots_1, d_roots)
    del d_states_1
    roots = d_roots.cpu().numpy().view(np.uint64)
    assert np.array_equal(roots[:, 0], w["old_roots"]) and np.array_equal(roots[:, 3], w["new_roots"])
    assert np.array_equal(roots[:, 1], roots[:, 2])                            # both routes to the interim root (IMT:277-294)
    assert np.array_equal(roots[1:, 0], roots[:-1, 3])                         # insert k starts where k - 1 ended
    rng = np.random.default_rng(7)
    for k in [0, b - 1] + [int(x) for x in rng.integers(0, b, 6)]:
        low, new = from_mont(w["low_leaves"][k]), from_mont(w["new_leaves"][k])
        new_low = np.stack([low[0], new[0], O.fe(m + k)])
        lsib, nsib = from_mont(w["low_siblings"][k]), from_mont(w["new_siblings"][k])
        got = d_states[k].cpu().numpy().view(np.uint64)
        folds = ((low, int(w["low_idx"][k]), lsib, 0, 1), (new_low, int(w["low_idx"][k]), lsib, depth + 1, depth + 2),
                 (np.zeros((3, 4), np.uint64), m + k, nsib, None, 2 * depth + 2), (new, m + k, nsib, 3 * depth + 2, 3 * depth + 3))
        for leaf, idx, sib, leaf_slot, path_slot in folds:
            h, ws = O.hash_trace(leaf)
            if leaf_slot is not None:
                assert np.array_equal(from_mont(got[leaf_slot].reshape(-1, 4)).reshape(132, 3, 4), ws), (k, leaf_slot)
            sample = set(int(x) for x in rng.integers(0, depth, 3)) | {0, depth - 1}
            for lvl in range(depth):
                pair = np.stack([h, sib[lvl]]) if idx % 2 == 0 else np.stack([sib[lvl], h])
                if lvl in sample:
                    h, ws = O.hash_trace(pair)
                    assert np.array_equal(from_mont(got[path_slot + lvl].reshape(-1, 4)).reshape(132, 3, 4), ws), (k, path_slot, lvl)
                else:
                    h = O.hash2(pair.reshape(1, 2, 4), 1)[0]
                idx //= 2


def test_fast_lookup_equals_plain_search_and_the_linear_scan(monkeypatch):
    """The prefix / shared-memory-top lookup (k_low_leaf_lookup_fast) against the plain binary search and the oracle's literal
    scan (IMT:632-660) — on random keys, on queries equal to keys, and on ADVERSARIAL keys: long runs that share their top 64
    bits (ties are resolved on the full keys), keys that differ only in the top limb, the smallest and largest values."""
    e = imt_b200.Engine(0, "canonical")
    rng = random.Random(99)
    n = 1 << 12
    C = (0x1234567890ABCDEF >> 2) << 192                                           # one top-64-bit prefix ...
    run = [C + (i << 64) + rng.getrandbits(60) for i in range(300)]                # ... shared by 300 keys
    top_only = [(k << 224) + 5 for k in range(1, 200)]                             # differ only in the top limb
    others = [rng.randrange(1, P) for _ in range(2000)] + [1, 2, P - 1, P - 2]
    vals = sorted(set(run + top_only + others))
    m = len(vals) + 1
    order = list(range(1, m))
    rng.shuffle(order)                                                             # slots in random (insertion) order
    slot_of = {v: s for v, s in zip(vals, order)}
    pre = np.zeros((n, 3, 4), np.uint64)
    chain = [0] + vals
    for a, b in zip(chain, chain[1:] + [None]):
        s = slot_of.get(a, 0)
        pre[s, 0] = O.fe(a)
        if b is not None:
            pre[s, 1], pre[s, 2] = O.fe(b), O.fe(slot_of[b])
    queries = ([v for v in run[::7]] + [v + 1 for v in run[::5]] + [v - 1 for v in run[::5]] + [C - 1, C + (301 << 64), C + (1 << 63)]
               + top_only[::9] + [v + 1 for v in top_only[::9]] + [0, 1, 2, 3, P - 1, P - 2, P - 3] + [rng.randrange(P) for _ in range(1500)])
    qv = O.fes(queries)
    want = [O.low_leaf(pre, qv[k]) for k in range(len(queries))]
    for fast_min in ("0", "1000000000"):                                           # the fast kernel, then the plain one
        monkeypatch.setenv("IMT_FAST_LOOKUP_MIN", fast_min)
        tree = e.build_from_leaves(pre)
        low, matched = tree.low_leaf_lookup(qv)
        for k in range(len(queries)):
            assert (int(low[k]) if matched[k] else 0, bool(matched[k])) == (want[k][0] if want[k][1] else 0, want[k][1]), (fast_min, k, hex(queries[k]))
        new = O.fes([C + (i << 64) + (1 << 62) for i in range(0, 300, 3)])        # inserts INTO the run: the prefix array is rebuilt
        tree.insert_batch(new)
        low2, matched2 = tree.low_leaf_lookup(qv)
        cur = tree.preimages(n)
        for k in range(0, len(queries), 11):
            w = O.low_leaf(cur, qv[k])
            assert (int(low2[k]) if matched2[k] else 0, bool(matched2[k])) == (w[0] if w[1] else 0, w[1]), (fast_min, k)
        tree.close()
    e.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_operation_sequences_against_the_oracle(eng, eng_mont, seed, tmp_path, monkeypatch):
    """Model-based test of the persistent device state: a random sequence of insert batches (random sizes, internal
    chunks forced small so that batches cross chunk boundaries), lookups, traced non-inclusion witnesses, checkpoint save + load,
    rebuilds from the current leaves and insert-trace calls — after every step the device tree equals the oracle advancing one
    insert at a time (IMT:710-741), and every answer equals the reference's literal scan (IMT:632-660)."""
    rng = random.Random(seed)
    depth = 8
    n = 1 << depth
    m = rng.randrange(1, 6)
    pre = synth.indexed_preimages(n, m, seed=100 + seed)
    e, enc, dec = (eng, lambda a: a, lambda a: a) if seed % 2 else (eng_mont, to_mont, from_mont)
    monkeypatch.setenv("IMT_INSERT_CHUNK", str(rng.choice([3, 8, 64])))
    tree = e.build_from_leaves(enc(pre))
    st = O.InsertState(pre, threads=2)
    occupied = m
    fresh = iter(synth.field_elements(800, seed=500 + seed))
    for step in range(14):
        op = rng.choice(["insert", "insert", "lookup", "trace", "checkpoint", "rebuild", "insert_trace"])
        if op in ("insert", "insert_trace") and occupied < n - 40:
            b = rng.randrange(1, 33)
            vals = np.stack([next(fresh) for _ in range(b)])
            got = tree.insert_batch(enc(vals), occupied)
            for k in range(b):
                want = st.insert(vals[k], occupied + k, incremental=True)
                for g, wk in zip(GPU, WIT):
                    a = got[g][k]
                    assert np.array_equal(dec(a) if np.asarray(a).dtype == np.uint64 and g != "low_idx" else a, want[wk]), (step, k, g)
            if op == "insert_trace":
                tr = e.trace_insert_witness(got, occupied)
                assert np.array_equal(tr["new_root"], got["new_roots"]) and np.array_equal(tr["zero_leaf_root"], tr["interim_root"])
                k = rng.randrange(b)
                assert np.array_equal(dec(got["fold_nodes"][k]), oracle_fold_nodes(got, k, occupied + k, dec))
            occupied += b
        elif op == "lookup":
            present = [st.pre[rng.randrange(1, occupied), 0] for _ in range(3)] if occupied > 1 else []
            qs = np.stack([next(fresh) for _ in range(5)] + present + [np.zeros(4, np.uint64)])
            low, matched = tree.low_leaf_lookup(enc(qs))
            for k in range(len(qs)):
                assert (int(low[k]), bool(matched[k])) == O.low_leaf(st.pre, qs[k]), (step, k)
        elif op == "trace":
            qs = np.stack([next(fresh) for _ in range(3)])
            o = tree.trace_non_inclusion(enc(qs))
            for k in range(3):
                wl, wm = O.low_leaf(st.pre, qs[k])
                assert (int(o["low_idx"][k]), bool(o["matched"][k])) == (wl, wm)
                assert np.array_equal(dec(o["low_leaves"][k]), st.pre[wl])
                h, ws = O.hash_trace(st.pre[wl])
                assert np.array_equal(dec(o["leaf_hash"][k]), ws)
                assert np.array_equal(dec(o["states"][k, -1, -1, 1]), st.root())          # the fold ends in the current root
        elif op == "checkpoint":
            path = str(tmp_path / f"t{step}.imt")
            tree.save(path)
            tree.close()
            tree = e.load_tree(path)
        elif op == "rebuild":
            tree.rebuild_from_leaves(enc(st.pre))
        assert np.array_equal(dec(tree.root()), st.root()), (step, op)
        assert tree.occupied == occupied
    assert np.array_equal(dec(tree.preimages(n)), st.pre)
    want_levels = O.levels(st.tree, n)
    for lvl in range(depth + 1):
        assert np.array_equal(dec(tree.level(lvl, n >> lvl)), want_levels[lvl]), lvl

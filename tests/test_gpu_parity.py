"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle on the same seeded inputs —
bit-exact (everything here is integer arithmetic). Run on the B200 box with `-m gpu`."""
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


def test_known_answer_of_the_reference(eng):
    # /root/reference/src/indexed_merkle_tree.rs:247-251
    out = eng.hash3(np.zeros((1, 3, 4), np.uint64))
    assert imt_b200.fe_to_int(out[0]) == 1960587138944869480785025106734196872454309951825657414575195034687326603497
    assert imt_b200.fe_to_int(eng.hash2(imt_b200.fes_from_ints([1, 2]))[0]) == int(GOLD["h2_1_2"])
    assert imt_b200.fe_to_int(eng.hash3(imt_b200.fes_from_ints([1, 2, 3]))[0]) == int(GOLD["h3_1_2_3"])


# 1 ... 1776: the lead / helper latency kernel (12 hashes per block, one block per SM: 12, 13 and 1776, 1777 are its edges);
# up to 8192: 3 lanes per hash; above: one thread per hash
@pytest.mark.parametrize("n", [1, 2, 12, 13, 127, 128, 129, 1776, 1777, 4099, 8193])
def test_batched_hashes_match_oracle(eng, n):
    x = synth.field_elements(3 * n, seed=n)
    assert np.array_equal(eng.hash3(x), O.hash3(x, 8))
    y = x[: 2 * n]
    assert np.array_equal(eng.hash2(y), O.hash2(y, 8))


def test_latency_kernels_agree_under_concurrent_load(eng):
    """The lead / helper latency kernel (csrc/poseidon_lh.cuh) hands operands between warps under named barriers; its protocol must
    hold under ANY delay of a warp, not only when the block has its SM to itself. A second context keeps every SM busy with
    thread-per-hash hashing (seven warps per scheduler beside the latency kernel's one) while small batches — one block, a partly
    filled block, one block per SM — are hashed over and over: every digest must equal the oracle's, every time."""
    import threading
    import torch
    other = imt_b200.Engine(0, "canonical")
    nbig = 1 << 21                                                  # ~32 ms of thread-per-hash hashing per call, inputs resident
    d_big = synth.field_elements_torch(3 * nbig, seed=99, device="cuda")
    d_out = torch.empty((nbig, 4), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    sizes = [1, 12, 13, 100, 1776]
    x = synth.field_elements(2 * max(sizes), seed=5).reshape(-1, 2, 4)
    want = O.hash2(x, 8)
    stop = threading.Event()
    loads = [0]

    def load():
        while not stop.is_set():
            other.hash3_dev(d_big, nbig, d_out)
            loads[0] += 1

    t = threading.Thread(target=load)
    t.start()
    try:
        for it in range(60):
            n = sizes[it % len(sizes)]
            assert np.array_equal(eng.hash2(x[:n]), want[:n]), (it, n)
    finally:
        stop.set()
        t.join()
        other.close()
    assert loads[0] >= 2    # the load really ran beside the small batches


def test_edge_values(eng):
    vals = [0, 1, 2, P - 1, P - 2, (1 << 64), (1 << 128) - 1, (1 << 253), P // 2, (1 << 254) % P]
    trip = [[a, b, c] for a in vals for b in vals[:4] for c in vals[-3:]]
    x = O.fes([v for t in trip for v in t])
    assert np.array_equal(eng.hash3(x), O.hash3(x, 8))
    pairs = O.fes([v for a in vals for b in vals for v in (a, b)])
    assert np.array_equal(eng.hash2(pairs), O.hash2(pairs, 8))


def test_trim_returns_cached_scratch(eng):
    import torch
    eng.hash3(synth.field_elements(3 * 200000, seed=1).reshape(-1, 3, 4))      # leaves ~25 MB of scratch in the pool
    eng.trim()
    assert np.array_equal(eng.hash2(O.fes([1, 2]).reshape(1, 2, 4)), O.fes([int(GOLD["h2_1_2"])]))   # and the engine still works


def test_empty_batch(eng):
    assert eng.hash2(np.zeros((0, 2, 4), np.uint64)).shape == (0, 4)
    assert eng.hash3(np.zeros((0, 3, 4), np.uint64)).shape == (0, 4)


def test_montgomery_format_is_halo2curves_memory_layout(eng_mont):
    x = synth.field_elements(300, seed=77)
    got = eng_mont.hash3(to_mont(x))
    assert np.array_equal(from_mont(got), O.hash3(x, 8))
    # R mod p, the Montgomery form of 1
    one = to_mont(O.fes([1]))
    assert [hex(int(v)) for v in one[0]] == ["0xac96341c4ffffffb", "0x36fc76959f60cd29", "0x666ea36f7879462e", "0xe0a77c19a07df2f"]


def test_non_canonical_input_is_rejected(eng, eng_mont):
    bad = O.fes([0, 0, 0]).copy()
    bad[1] = np.array([(P >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)  # == p
    for e in (eng, eng_mont):
        with pytest.raises(imt_b200.ImtError) as ei:
            e.hash3(bad)
        assert ei.value.status == _ffi.ERR_NON_CANONICAL
    all_ones = np.full((2, 4), 0xFFFFFFFFFFFFFFFF, np.uint64)
    with pytest.raises(imt_b200.ImtError):
        eng.hash2(all_ones)


def test_large_and_small_batches_take_different_kernels_and_agree(eng):
    """<= 8192 hashes run on the latency kernels (lead / helper up to 12 hashes per SM, 3 lanes per hash above), larger batches on
    the thread-per-hash kernel: same digests from all, and all reject a non-canonical element."""
    n = 9000
    x = synth.field_elements(3 * n, seed=4242).reshape(n, 3, 4)
    big = eng.hash3(x)
    assert np.array_equal(big, O.hash3(x, 8))
    assert np.array_equal(eng.hash3(x[:8192]), big[:8192]) and np.array_equal(eng.hash3(x[8192:]), big[8192:])
    y = x.reshape(-1, 2, 4)[:8500]
    assert np.array_equal(eng.hash2(y), O.hash2(y, 8))
    bad = x.copy()
    bad[8999, 2] = np.array([(P >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
    with pytest.raises(imt_b200.ImtError) as e:
        eng.hash3(bad)
    assert e.value.status == _ffi.ERR_NON_CANONICAL


@pytest.mark.parametrize("depth", [1, 3, 7, 12])
def test_tree_build_every_level(eng, depth):
    n = 1 << depth
    pre = synth.indexed_preimages(n, max(1, n - 3), seed=depth)
    leaves = O.hash3(pre, 8)
    want = O.levels(O.tree_build(leaves, 8), n)
    t = eng.build_from_leaves(pre)
    t2 = eng.build_from_hashes(leaves)
    assert t.depth == depth and t.num_leaves == n
    for lvl in range(depth + 1):
        assert np.array_equal(t.level(lvl, n >> lvl), want[lvl])
        assert np.array_equal(t2.level(lvl, n >> lvl), want[lvl])
    assert np.array_equal(t.root(), want[-1][0])
    assert np.array_equal(t.preimages(n), pre)


def test_tree_build_errors_match_the_reference(eng):
    with pytest.raises(imt_b200.ImtError, match="Cannot create Merkle Tree with no leaves") as e:
        eng.build_from_hashes(np.zeros((0, 4), np.uint64))
    assert e.value.status == _ffi.ERR_EMPTY
    with pytest.raises(imt_b200.ImtError, match="Leaves must be even") as e:
        eng.build_from_hashes(O.fes([1, 2, 3]))
    assert e.value.status == _ffi.ERR_ODD
    with pytest.raises(imt_b200.ImtError) as e:
        eng.build_from_hashes(O.fes([1, 2, 3, 4, 5, 6]))
    assert e.value.status == _ffi.ERR_NOT_POW2
    one = eng.build_from_hashes(O.fes([7]))  # utils.rs:27-33: tree = [leaf], root = leaf
    assert imt_b200.fe_to_int(one.root()) == 7 and one.depth == 0


def test_get_proofs_and_verify(eng):
    depth, n = 9, 512
    leaves = synth.field_elements(n, seed=42)
    tree_o = O.tree_build(leaves, 8)
    t = eng.build_from_hashes(leaves)
    idx = np.array([0, 1, 2, 3, n - 1, n - 2, 255, 256] + random.Random(1).sample(range(n), 40), np.uint64)
    sib, hel = t.get_proofs(idx)
    sib_fe, hel_fe = t.get_proofs(idx, helpers_as_fe=True)
    for k, i in enumerate(idx):
        s, h = O.get_proof(tree_o, n, int(i))
        assert np.array_equal(sib[k], s) and np.array_equal(hel[k], h)
        assert np.array_equal(sib_fe[k], s)
        assert O.to_ints(hel_fe[k]) == [int(v) for v in h]  # F::from(1) / F::from(0), utils.rs:79
    root = t.root()
    ok = eng.verify_proofs(leaves[idx.astype(np.int64)], idx, root, sib)
    assert ok.all()
    wrong = leaves[idx.astype(np.int64)].copy()
    wrong[3, 0] ^= np.uint64(1)
    ok = eng.verify_proofs(wrong, idx, root, sib)
    assert not ok[3] and ok.sum() == len(idx) - 1
    for k, i in enumerate(idx[:5]):  # the oracle accepts the GPU's paths too
        assert O.verify_proof(leaves[int(i)], int(i), root, sib[k])
    with pytest.raises(imt_b200.ImtError) as e:
        t.get_proofs(np.array([n], np.uint64))
    assert e.value.status == _ffi.ERR_INDEX_OOB


def test_hash_traces(eng, eng_mont):
    rng = random.Random(3)
    for arity in (2, 3):
        ins = [[0] * arity, [1, 2, 3][:arity], [P - 1] * arity] + [[rng.randrange(P) for _ in range(arity)] for _ in range(130)]
        x = O.fes([v for row in ins for v in row]).reshape(-1, arity, 4)
        dg, st = eng.trace_hashes(x, arity)
        dgm, stm = eng_mont.trace_hashes(to_mont(x), arity)
        for k in (0, 1, 2, 3, 64, 131, 132):
            wd, ws = O.hash_trace(x[k])
            assert np.array_equal(dg[k], wd) and np.array_equal(st[k], ws)
            assert np.array_equal(from_mont(dgm[k]), wd) and np.array_equal(from_mont(stm[k]), ws)
        want = O.hash3(x, 8) if arity == 3 else O.hash2(x, 8)
        assert np.array_equal(dg, want)
        # every trace's last state carries the digest in lane 1
        assert np.array_equal(st[:, -1, 1, :], dg)
    xor = 0
    for v in O.to_ints(eng.trace_hashes(np.zeros((1, 3, 4), np.uint64), 3)[1].reshape(-1, 4)):
        xor ^= v
    assert xor == int(GOLD["trace_h3_zero_xor"])


def test_merkle_proof_traces(eng):
    depth, n = 6, 64
    leaves = synth.field_elements(n, seed=5)
    t = eng.build_from_hashes(leaves)
    idx = np.array([0, 1, 37, 63], np.uint64)
    sib, hel = t.get_proofs(idx)
    roots, states = eng.trace_merkle_proofs(leaves[idx.astype(np.int64)], idx, sib)
    assert states.shape == (4, depth, 132, 3, 4)
    for k, i in enumerate(idx):
        h, index = leaves[int(i)], int(i)
        for lvl in range(depth):
            pair = np.stack([h, sib[k, lvl]]) if index % 2 == 0 else np.stack([sib[k, lvl], h])
            h, ws = O.hash_trace(pair)
            assert np.array_equal(states[k, lvl], ws)
            index //= 2
        assert np.array_equal(roots[k], h) and np.array_equal(h, t.root())


def test_tree_trace_proofs_equal_get_proof_plus_fold_trace(eng, eng_mont):
    """imt_tree_trace_proofs (one independent traced hash per (query, level), operands read from the stored levels) must
    give the bytes of get_proof + the serial fold trace — both formats, a sharded tree with a cap, an any-width context"""
    depth, n = 7, 128
    pre = synth.indexed_preimages(n, 90, seed=31)
    idx = np.array([0, 1, 2, 77, 126, 127, 77], np.uint64)
    for e, enc in ((eng, lambda a: a), (eng_mont, to_mont)):
        t = e.build_from_leaves(enc(pre))
        sib, _ = t.get_proofs(idx)
        leaves = t.level(0, n)[idx.astype(np.int64)]
        roots, want = e.trace_merkle_proofs(leaves, idx, sib)
        got = t.trace_proofs(idx)
        assert got.shape == (len(idx), depth, 132, 3, 4) and np.array_equal(got, want)
        assert np.array_equal(got[:, -1, -1, 1], roots)                 # the last traced state of every path holds the root
        with pytest.raises(imt_b200.ImtError) as ex:
            t.trace_proofs(np.array([n], np.uint64))
        assert ex.value.status == _ffi.ERR_INDEX_OOB
    # sharded: the cap levels of the path come from the replicated cap
    world, per = 4, n // 4
    whole = eng.build_from_leaves(pre)
    shards = [eng.build_from_leaves(pre[r * per:(r + 1) * per]) for r in range(world)]
    sub_roots = np.stack([s.root() for s in shards])
    for r, s in enumerate(shards):
        s.attach_cap(r, world, sub_roots)
        own = np.array([r * per, r * per + 5, (r + 1) * per - 1], np.uint64)
        assert np.array_equal(s.trace_proofs(own), whole.trace_proofs(own))
    # any-width context
    g = imt_b200.Engine(0, "canonical", t=4, rate=3, r_f=8, r_p=56)
    tg = g.build_from_leaves(pre)
    sib, _ = tg.get_proofs(idx)
    _, want = g.trace_merkle_proofs(tg.level(0, n)[idx.astype(np.int64)], idx, sib)
    assert np.array_equal(tg.trace_proofs(idx), want) and want.shape == (len(idx), depth, 65, 4, 4)
    g.close()


def test_tree_trace_proofs_chunked_pipeline(eng_mont):
    """more queries than one pipeline chunk (8192): the double-buffered host path must deliver every chunk"""
    depth, n, q = 4, 16, 8192 * 2 + 100
    t = eng_mont.build_from_hashes(to_mont(synth.field_elements(n, seed=8)))
    idx = (np.arange(q, dtype=np.uint64) * np.uint64(7)) % np.uint64(n)
    got = t.trace_proofs(idx)
    ref = t.trace_proofs(np.arange(n, dtype=np.uint64))
    assert np.array_equal(got, ref[idx.astype(np.int64)])


def test_reference_api_mirror_runs_the_reference_test_scenario():
    """The native half of test_insert_leaf_multiple_round (indexed_merkle_tree.rs:679-741), written against the same
    names the reference uses, with its O(n) re-hash + rebuild per round — all hashing on the GPU."""
    from imt_b200 import Poseidon, IndexedMerkleTree, IndexedMerkleTreeLeaf as IMTLeaf, hash_nullifier_pre_images
    native_hasher = Poseidon(8, 57)
    native_hasher.update([0, 0, 0])
    assert native_hasher.squeeze_and_reset() == int(GOLD["kat_h3_zero"])  # test_hash_zero, IMT:805-810
    pre = [IMTLeaf(0, 0, 0) for _ in range(8)]
    leaves = hash_nullifier_pre_images(pre)
    tree = IndexedMerkleTree.new(native_hasher, leaves)
    assert tree.get_root() == int(GOLD["empty_depth3_root"])
    for rnd, new_val in enumerate(GOLD["scenario_inserts"]):
        old_root = tree.get_root()
        arr = O.fes([v for l in pre for v in l.as_tuple()]).reshape(8, 3, 4)
        new_arr, low_idx, _ = O.update_idx_leaf(arr, O.fe(new_val), rnd + 1)   # host-side list surgery (IMT:632-660)
        assert low_idx == GOLD["scenario_low_idx"][rnd]
        low_proof, low_helper = tree.get_proof(low_idx)
        assert tree.verify_proof(leaves[low_idx], low_idx, old_root, low_proof)
        pre = [IMTLeaf(*O.to_ints(l)) for l in new_arr]
        leaves = hash_nullifier_pre_images(pre)
        tree = IndexedMerkleTree.new(native_hasher, leaves)
        new_proof, new_helper = tree.get_proof(rnd + 1)
        assert tree.get_root() == int(GOLD["scenario_roots"][rnd])
        assert tree.verify_proof(leaves[rnd + 1], rnd + 1, tree.get_root(), new_proof)
        assert new_helper == [1 if ((rnd + 1) >> l) % 2 == 0 else 0 for l in range(3)]
    with pytest.raises(ValueError, match="Leaves must be even"):
        IndexedMerkleTree.new(native_hasher, [1, 2, 3])
    with pytest.raises(ValueError, match="Cannot create Merkle Tree with no leaves"):
        IndexedMerkleTree.new(native_hasher, [])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_subtree_sharding_reassembles_the_single_tree(eng, world):
    depth, n = 10, 1024
    pre = synth.indexed_preimages(n, seed=11)
    whole = eng.build_from_leaves(pre)
    per = n // world
    shards = [eng.build_from_leaves(pre[r * per:(r + 1) * per]) for r in range(world)]
    roots = np.stack([s.root() for s in shards])
    for r, s in enumerate(shards):
        s.attach_cap(r, world, roots)
        assert s.depth == depth and s.num_leaves == n
        assert np.array_equal(s.root(), whole.root())
        idx = np.array([r * per, r * per + 1, (r + 1) * per - 1, r * per + per // 3], np.uint64)
        sib, hel = s.get_proofs(idx)
        wsib, whel = whole.get_proofs(idx)
        assert np.array_equal(sib, wsib) and np.array_equal(hel, whel)
        with pytest.raises(imt_b200.ImtError):
            s.get_proofs(np.array([((r + 1) % world) * per], np.uint64))
    assert np.array_equal(whole.root(), O.build_from_preimages(pre, 4))     # (seed 11: not the golden file's default-seed tree)


def test_depth16_roots_match_golden(eng):
    n = 1 << 16
    assert imt_b200.fe_to_int(eng.build_from_leaves(synth.random_preimages(n)).root()) == int(GOLD["build_roots"]["16"]["random"])
    assert imt_b200.fe_to_int(eng.build_from_leaves(synth.indexed_preimages(n)).root()) == int(GOLD["build_roots"]["16"]["indexed"])


def test_depth20_root_matches_golden_and_round_trips(eng):
    """BASELINE config[1]: 1M synthetic leaves. Root vs the committed oracle fixture; then size-independent
    properties: every sampled path folds back to the root (GPU and oracle verifiers), rebuild is idempotent."""
    n = 1 << 20
    pre = synth.indexed_preimages(n)
    t = eng.build_from_leaves(pre)
    root = t.root()
    assert imt_b200.fe_to_int(root) == int(GOLD["build_roots"]["20"]["indexed"])
    idx = np.array(random.Random(20).sample(range(n), 2048) + [0, n - 1], np.uint64)
    sib, hel = t.get_proofs(idx)
    leaf_hashes = O.hash3(pre[idx.astype(np.int64)], 8)
    assert eng.verify_proofs(leaf_hashes, idx, root, sib).all()
    for k in range(4):
        assert O.verify_proof(leaf_hashes[k], int(idx[k]), root, sib[k])
    assert np.array_equal(hel, ((idx[:, None] >> np.arange(20, dtype=np.uint64)[None, :]) & np.uint64(1)) == 0)
    t.rebuild_from_leaves(pre)
    assert np.array_equal(t.root(), root)


def test_depth24_roots_match_golden(eng):
    """The headline size (BASELINE metric: depth-24 build, 33 554 431 hashes): roots of the random and the indexed
    synthetic trees against the oracle's (tests/golden/make_golden.py --depth24, ~6 min per tree on 8 host cores).
    The leaves are synthesised on the device (the torch generators are checked against the numpy ones on the CPU)."""
    import torch
    n = 1 << 24
    d_pre = synth.field_elements_torch(3 * n, device="cuda").view(n, 3, 4)
    t = eng.build_from_leaves_dev(d_pre, n)
    assert imt_b200.fe_to_int(t.root()) == int(GOLD["build_roots"]["24"]["random"])
    d_pre = synth.indexed_preimages_torch(n, device="cuda")
    torch.cuda.synchronize()
    t.rebuild_from_leaves_dev(d_pre)
    assert imt_b200.fe_to_int(t.root()) == int(GOLD["build_roots"]["24"]["indexed"])
    assert t.occupied == n


def test_depth26_build_composes_from_depth24_subtrees(eng):
    """Beyond the headline size (4x the BASELINE depth-24 tree: 2^26 leaves, 134 217 727 hashes, 6 GiB of preimages): no oracle
    root exists at this size, so the check is the size-independent composition property of utils.rs:41-51 — the root of the
    whole tree is H2(H2(r0, r1), H2(r2, r3)) of the roots of its four contiguous depth-24 quarters (each built on its own,
    the two cap levels recomputed by the ORACLE on the CPU) — plus paths of the first / last / random leaves (indices beyond
    2^24 and the top helper bits) verified on the GPU and, for two of them, folded by the oracle."""
    import torch
    depth = 26
    n, q = 1 << depth, 1 << 24
    d_pre = synth.field_elements_torch(3 * n, seed=26, device="cuda").view(n, 3, 4)
    torch.cuda.synchronize()
    whole = eng.build_from_leaves_dev(d_pre, n)
    assert whole.depth == depth and whole.num_leaves == n
    root = whole.root()
    quarter = eng.build_from_leaves_dev(d_pre[:q], q)
    roots = [quarter.root()]
    for k in range(1, 4):
        quarter.rebuild_from_leaves_dev(d_pre[k * q:(k + 1) * q])
        roots.append(quarter.root())
    r = np.stack(roots)
    mid = O.hash2(r.reshape(4, 4), 1)            # H2(r0, r1), H2(r2, r3) on the CPU
    assert np.array_equal(O.hash2(mid.reshape(2, 4), 1)[0], root)
    assert np.array_equal(whole.level(24, 4), r)
    rng = np.random.default_rng(26)
    idx = np.concatenate([[0, n - 1, q, q - 1, 3 * q + 12345], rng.integers(0, n, 59)]).astype(np.uint64)
    sib, hel = whole.get_proofs(idx)
    pre = d_pre[torch.from_numpy(idx.astype(np.int64)).cuda()].cpu().numpy().view(np.uint64)
    leaf_hashes = O.hash3(pre, 8)
    assert eng.verify_proofs(leaf_hashes, idx, root, sib).all()
    assert np.array_equal(hel, ((idx[:, None] >> np.arange(depth, dtype=np.uint64)[None, :]) & np.uint64(1)) == 0)
    for k in (1, 4):
        assert O.verify_proof(leaf_hashes[k], int(idx[k]), root, sib[k])
    bad = sib.copy()
    bad[1, depth - 1, 0] ^= np.uint64(1)
    assert not eng.verify_proofs(leaf_hashes[1:2], idx[1:2], root, bad[1:2])[0]


def test_verify_rejects_a_root_encoded_as_root_plus_p(eng):
    """a root given as root + p (still < 2^256) must not fold onto its canonical twin: every verify path — the
    3-lanes-per-path kernel (q <= 8192), the thread-per-path kernel (q > 8192) and the any-width kernels — reports
    IMT_ERR_NON_CANONICAL, as the header promises for every input >= p."""
    from imt_b200 import _ffi
    n, depth = 64, 6
    pre = synth.indexed_preimages(n, 40, seed=3)
    tree = eng.build_from_leaves(pre)
    root = tree.root()
    bad_root = imt_b200.fe_from_int(imt_b200.fe_to_int(root) + imt_b200.P)
    assert imt_b200.fe_to_int(bad_root) < (1 << 256)
    leaves = eng.hash3(pre)
    for q in (1, 9000):
        idx = (np.arange(q) % n).astype(np.uint64)
        sib, _ = tree.get_proofs(idx)
        assert eng.verify_proofs(leaves[idx], idx, root, sib).all()
        with pytest.raises(imt_b200.ImtError) as ei:
            eng.verify_proofs(leaves[idx], idx, bad_root, sib)
        assert ei.value.status == _ffi.ERR_NON_CANONICAL
    g = imt_b200.Engine(0, "canonical", generic=True)
    gt = g.build_from_leaves(pre)
    idx = np.array([5], np.uint64)
    sib, _ = gt.get_proofs(idx)
    assert g.verify_proofs(leaves[idx], idx, root, sib).all()
    with pytest.raises(imt_b200.ImtError) as ei:
        g.verify_proofs(leaves[idx], idx, bad_root, sib)
    assert ei.value.status == _ffi.ERR_NON_CANONICAL
    g.close()


def test_trace_2p16_paths_of_the_depth20_tree_sampled_against_the_oracle(eng):
    """BASELINE config 4 at its depth-20 size: the verify_merkle_proof witness traces of 2^16 uniform random paths of the
    1M-leaf tree (16.6 GB of states, kept on the device), 64 sampled paths checked state by state against the oracle, which
    builds the whole tree itself (leaf hashes, levels, paths) and traces every hash of the fold."""
    import torch
    depth, q = 20, 1 << 16
    n = 1 << depth
    pre = synth.indexed_preimages(n, n - 1000, seed=20)
    th = O.max_threads()
    levels = O.tree_build(O.hash3(pre, th), th)
    tree = eng.build_from_leaves(pre)
    assert np.array_equal(tree.root(), levels[-1])
    rng = np.random.default_rng(16)
    idx = rng.integers(0, n, q).astype(np.uint64)
    dev = torch.device("cuda", 0)
    d_idx = torch.from_numpy(idx.view(np.int64)).to(dev)
    d_states = torch.empty((q, depth, 132, 3, 4), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()                                                  # the engine runs on its own stream: inputs must be complete
    tree.trace_proofs_dev(d_idx, q, d_states)
    d_root = torch.from_numpy(levels[-1].view(np.int64).copy()).to(dev)
    assert bool((d_states[:, -1, -1, 1] == d_root).all())                     # every one of the 65 536 folds ends in the root
    for k in [0, q - 1] + [int(x) for x in rng.integers(0, q, 62)]:
        i = int(idx[k])
        sib, _ = O.get_proof(levels, n, i)
        h = O.hash3(pre[i:i + 1], 1)[0]
        got = d_states[k].cpu().numpy().view(np.uint64)
        for lvl in range(depth):
            pair = np.stack([h, sib[lvl]]) if i % 2 == 0 else np.stack([sib[lvl], h])
            h, ws = O.hash_trace(pair)
            assert np.array_equal(got[lvl], ws), (k, lvl)
            i //= 2
    tree.close()

"""GPU parity tests of the any-width Poseidon instances (utils.rs:6, 19 and indexed_merkle_tree.rs:65, 127,
231 are generic over T and RATE): the any-width kernels through the C-ABI against the Python oracle and against the
published permutation vectors, and — for the reference's own instance — against the tuned kernels. Bit-exact."""
import numpy as np
import pytest

import imt_b200
from imt_b200 import synth
import poseidon_ref as R
from test_spec_params import KAT_T3, KAT_T5, GOLD

pytestmark = pytest.mark.gpu
P = imt_b200.P
INSTANCES = [(2, 8, 56), (3, 8, 57), (4, 8, 56), (5, 8, 60), (3, 6, 10)]
_specs = {}


def spec(t, r_f, r_p):
    if (t, r_f, r_p) not in _specs:
        _specs[(t, r_f, r_p)] = R.Spec(r_f, r_p, t)
    return _specs[(t, r_f, r_p)]


def engine(t, r_f, r_p, fmt="canonical"):
    return imt_b200.Engine(0, fmt, t=t, rate=t - 1, r_f=r_f, r_p=r_p, generic=True)


def ints(a):
    return [imt_b200.fe_to_int(r) for r in np.asarray(a).reshape(-1, 4)]


def test_published_permutation_vectors_on_the_gpu():
    e3, e5 = engine(3, 8, 57), engine(5, 8, 60)
    assert ints(e3.permute(imt_b200.fes_from_ints([0, 1, 2]))) == KAT_T3
    assert ints(e5.permute(imt_b200.fes_from_ints([0, 1, 2, 3, 4]))) == KAT_T5
    # the tuned context reaches the same permutation through its lazily derived any-width parameters
    d = imt_b200.Engine(0, "canonical")
    assert ints(d.permute(imt_b200.fes_from_ints([0, 1, 2]))) == KAT_T3


def test_golden_digests_of_every_instance():
    """committed vectors (tests/golden/golden.json, written by make_golden.py --spec-only): H(1, ..., k) for k = 0..6"""
    for key, digests in GOLD["hashes"].items():
        t, r_f, r_p = (int(x) for x in key.split(","))
        e = engine(t, r_f, r_p)
        for k, d in digests.items():
            k = int(k)
            got = e.hash(imt_b200.fes_from_ints(list(range(1, k + 1))).reshape(1, k, 4), k, n=1)
            assert imt_b200.fe_to_int(got[0]) == int(d), (key, k)
        e.close()


@pytest.mark.parametrize("t,r_f,r_p", INSTANCES)
def test_hash_of_every_input_length_matches_the_oracle(t, r_f, r_p):
    e, sp = engine(t, r_f, r_p), spec(t, r_f, r_p)
    for arity in range(0, 2 * t + 2):
        n = 5
        x = synth.field_elements(max(arity, 1) * n, seed=100 * t + arity)[: arity * n].reshape(n, arity, 4)
        got = ints(e.hash(x, arity, n=n))
        want = [R.hash_n(ints(x[i]) if arity else [], sp) for i in range(n)]
        assert got == want, (t, arity)


def test_montgomery_format_and_batch_sizes():
    t, r_f, r_p = 4, 8, 56
    e, em, sp = engine(t, r_f, r_p), engine(t, r_f, r_p, "montgomery"), spec(t, r_f, r_p)
    for n in (1, 127, 129, 1000):
        x = synth.field_elements(3 * n, seed=n).reshape(n, 3, 4)
        got = e.hash(x, 3)
        xm = imt_b200.fes_from_ints([v * (1 << 256) % P for v in ints(x)]).reshape(n, 3, 4)
        gm = em.hash(xm, 3)
        assert [v * (1 << 256) % P for v in ints(got)] == ints(gm)
        for i in (0, n // 2, n - 1):
            assert imt_b200.fe_to_int(got[i]) == R.hash_n(ints(x[i]), sp)


def test_generic_kernels_equal_the_tuned_ones_on_the_reference_instance():
    tuned, gen = imt_b200.Engine(0, "canonical"), engine(3, 8, 57)
    n = 300
    x = synth.field_elements(3 * n, seed=7)
    assert np.array_equal(tuned.hash3(x), gen.hash3(x))            # imt_poseidon_hash3 on a generic context
    assert np.array_equal(tuned.hash2(x[: 2 * n]), gen.hash2(x[: 2 * n]))
    assert np.array_equal(tuned.hash(x.reshape(n, 3, 4), 3), gen.hash(x.reshape(n, 3, 4), 3))
    assert imt_b200.fe_to_int(gen.hash3(np.zeros((1, 3, 4), np.uint64))[0]) == R.KAT_H3_ZERO   # indexed_merkle_tree.rs:247-251
    dt, st = tuned.trace_hashes(x[:6].reshape(2, 3, 4), 3)
    dg, sg = gen.trace_hashes(x[:6].reshape(2, 3, 4), 3)
    assert np.array_equal(dt, dg) and np.array_equal(st, sg)
    # input lengths the tuned kernels do not cover (1, 4, 5, 0) run on the any-width kernels of the same context
    for arity in (0, 1, 4, 5):
        y = synth.field_elements(max(arity, 1) * 3, seed=arity)[: arity * 3].reshape(3, arity, 4)
        assert ints(tuned.hash(y, arity, n=3)) == [R.hash_n(ints(y[i]) if arity else []) for i in range(3)]
    d1, s1 = tuned.trace_hashes(x[:5].reshape(1, 5, 4), 5)
    wd, ws = R.hash_trace_n(ints(x[:5]))
    assert ints(d1) == [wd] and ints(s1) == [v for stt in ws for v in stt]


@pytest.mark.parametrize("t,r_f,r_p", [(2, 8, 56), (4, 8, 56), (5, 8, 60)])
def test_witness_trace_matches_the_oracle(t, r_f, r_p):
    e, sp = engine(t, r_f, r_p), spec(t, r_f, r_p)
    for arity in (2, 3):
        x = synth.field_elements(arity * 2, seed=t + arity).reshape(2, arity, 4)
        dg, st = e.trace_hashes(x, arity)
        assert st.shape == (2, (arity // (t - 1) + 1) * (1 + r_f + r_p), t, 4)
        for i in range(2):
            wd, ws = R.hash_trace_n(ints(x[i]), sp)
            assert imt_b200.fe_to_int(dg[i]) == wd
            assert ints(st[i]) == [v for stt in ws for v in stt]


def _flat_sbox(cells):
    return [v for c in cells for v in c]


def test_extended_sbox_trace_matches_the_oracle():
    """SURVEY 8a row 9, optional part: (x^2, x^4, x^5 + c) of every S-box, tuned kernels (arity 2 / 3, both formats), the
    any-width kernels (other arities, other instances) and the tree-path trace"""
    tuned = imt_b200.Engine(0, "canonical")
    mont = imt_b200.Engine(0, "montgomery")
    R256 = 1 << 256
    for arity in (2, 3, 5):
        x = synth.field_elements(arity * 3, seed=40 + arity).reshape(3, arity, 4)
        dg, st, sb = tuned.trace_hashes_ext(x, arity)
        assert sb.shape == (3, (arity // 2 + 1) * 81, 3, 4)
        xm = imt_b200.fes_from_ints([v * R256 % P for v in ints(x)]).reshape(3, arity, 4)
        dgm, stm, sbm = mont.trace_hashes_ext(xm, arity)
        for i in range(3):
            cells = []
            wd, ws = R.hash_trace_n(ints(x[i]), None, cells)
            assert imt_b200.fe_to_int(dg[i]) == wd and ints(st[i]) == [v for s_ in ws for v in s_]
            assert ints(sb[i]) == _flat_sbox(cells), arity
            assert ints(sbm[i]) == [v * R256 % P for v in _flat_sbox(cells)]
        d0, s0 = tuned.trace_hashes(x, arity)                      # the plain trace is unchanged by the extension
        assert np.array_equal(d0, dg) and np.array_equal(s0, st)
    g = engine(4, 8, 56)
    sp = spec(4, 8, 56)
    x = synth.field_elements(6, seed=3).reshape(2, 3, 4)
    dg, st, sb = g.trace_hashes_ext(x, 3)
    assert sb.shape == (2, 2 * (8 * 4 + 56), 3, 4)              # 3 inputs fill one RATE-chunk; the padding takes a second permutation
    for i in range(2):
        cells = []
        wd, ws = R.hash_trace_n(ints(x[i]), sp, cells)
        assert imt_b200.fe_to_int(dg[i]) == wd and ints(sb[i]) == _flat_sbox(cells)
    # tree paths: every (query, level) hash with its S-box cells
    n = 16
    leaves = synth.field_elements(n, seed=9)
    for e, sp_ in ((tuned, None), (g, sp)):
        t = e.build_from_hashes(leaves)
        idx = np.array([0, 5, 15], np.uint64)
        st, sb = t.trace_proofs_ext(idx)
        assert np.array_equal(st, t.trace_proofs(idx))
        for k, i in enumerate(idx):
            for lvl in range(4):
                pair = ints(t.level(lvl, n >> lvl)[(int(i) >> lvl) & ~1:][:2])
                cells = []
                R.hash_trace_n(pair, sp_, cells)
                assert ints(sb[k, lvl]) == _flat_sbox(cells), (k, lvl)


@pytest.mark.parametrize("t,r_f,r_p", [(2, 8, 56), (4, 8, 56), (5, 8, 60)])
def test_tree_paths_folds_and_inserts_with_another_instance(t, r_f, r_p):
    """the whole tree API on an any-width context: the mirror of the reference's test_insert_leaf_multiple_round
    (indexed_merkle_tree.rs:679-803) with Poseidon::<Fr, T, T-1>"""
    e, sp = engine(t, r_f, r_p), spec(t, r_f, r_p)
    h2 = lambda a, b: R.hash_n([a, b], sp)          # noqa: E731
    h3 = lambda a, b, c: R.hash_n([a, b, c], sp)    # noqa: E731
    depth = 3
    n = 1 << depth
    pre = [[0, 0, 0] for _ in range(n)]
    tree = e.build_from_leaves(imt_b200.fes_from_ints([v for p in pre for v in p]).reshape(n, 3, 4))
    want = R.IndexedMerkleTree([h3(*p) for p in pre], h2)
    assert imt_b200.fe_to_int(tree.root()) == want.root
    vals = [30, 10, 20, 5, 50, 35]                  # indexed_merkle_tree.rs:683-690
    w = tree.insert_batch(imt_b200.fes_from_ints(vals))
    for r, v in enumerate(vals):
        new_pre, low = R.update_idx_leaf(pre, v, r + 1)
        assert int(w["low_idx"][r]) == low
        assert ints(w["low_siblings"][r]) == want.get_proof(low)[0]
        assert imt_b200.fe_to_int(w["old_roots"][r]) == want.root
        want = R.IndexedMerkleTree([h3(*p) for p in new_pre], h2)
        assert imt_b200.fe_to_int(w["new_roots"][r]) == want.root
        assert ints(w["new_siblings"][r]) == want.get_proof(r + 1)[0]
        pre = new_pre
    assert imt_b200.fe_to_int(tree.root()) == want.root
    for lvl in range(depth + 1):
        assert ints(tree.level(lvl, n >> lvl)) == want.tree[lvl]
    idx = np.arange(n, dtype=np.uint64)
    sib, hel = tree.get_proofs(idx)
    leaves = tree.level(0, n)
    roots = np.broadcast_to(tree.root(), (n, 4)).copy()
    assert e.verify_proofs(leaves, idx, roots, sib).all()
    bad = sib.copy()
    bad[3, 1, 0] ^= 1
    ok = e.verify_proofs(leaves, idx, roots, bad)
    assert not ok[3] and ok.sum() == n - 1
    rts, st = e.trace_merkle_proofs(leaves, idx, sib)
    assert np.array_equal(rts, roots)
    hh, ix = want.tree[0][5], 5
    for lvl in range(depth):
        s = want.get_proof(5)[0][lvl]
        a, b = (hh, s) if ix % 2 == 0 else (s, hh)
        hh, ws = R.hash_trace_n([a, b], sp)
        assert ints(st[5, lvl]) == [v for stt in ws for v in stt]
        ix //= 2
    # low-leaf lookups are hash-free but go through the same tree object
    low, matched = tree.low_leaf_lookup(imt_b200.fes_from_ints([7, 31, 60]))
    assert matched.all() and [int(x) for x in low] == [4, 1, 5]


def test_reference_mirror_with_another_instance():
    t, r_f, r_p = 4, 8, 56
    sp = spec(t, r_f, r_p)
    h = imt_b200.Poseidon(r_f, r_p, t=t, rate=t - 1)
    h.update([1, 2])
    h.update([3, 4, 5])
    assert h.squeeze_and_reset() == R.hash_n([1, 2, 3, 4, 5], sp)
    assert h.squeeze_and_reset() == R.hash_n([], sp)                      # reset: an empty sponge again
    leaves = [R.hash_n([i, 0, 0], sp) for i in range(8)]
    tree = imt_b200.IndexedMerkleTree.new(h, leaves)
    want = R.IndexedMerkleTree(leaves, lambda a, b: R.hash_n([a, b], sp))
    assert tree.get_root() == want.get_root()
    proof, helper = tree.get_proof(6)
    assert (proof, helper) == want.get_proof(6)
    assert tree.verify_proof(leaves[6], 6, tree.get_root(), proof)
    # default instance: any input length through the mirror
    d = imt_b200.Poseidon(8, 57)
    d.update([1, 2, 3, 4])
    assert d.squeeze_and_reset() == R.hash_n([1, 2, 3, 4])


def test_unsupported_instance_raises():
    with pytest.raises(imt_b200.ImtError):
        imt_b200.Engine(0, "canonical", t=6, rate=5)
    with pytest.raises(imt_b200.ImtError):
        imt_b200.Engine(0, "canonical", t=3, rate=1)

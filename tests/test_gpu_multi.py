"""Multi-GPU behind the C-ABI (csrc/imt_comm.cu): `IndexedMerkleTree::new` (utils.rs:20-57) sharded by subtree with the root
exchange INSIDE the library. On a one-GPU box the group lists device 0 several times and exchanges by stream-ordered copies
(NCCL refuses duplicate devices) — every other line of the code path is the one the 8-GPU run takes; with >= 2 GPUs the same
tests run over NCCL (single process: ncclCommInitAll; process per GPU: ncclCommInitRank from plain C, no Python)."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import imt_b200
from imt_b200 import _ffi, synth
import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
pytestmark = pytest.mark.gpu


def _devices(world):
    import torch
    n = torch.cuda.device_count()
    return [i % n for i in range(world)]


@pytest.fixture(scope="module")
def eng():
    e = imt_b200.Engine(0, "canonical")
    yield e
    e.close()


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_multi_build_equals_single_gpu_and_oracle(eng, world):
    depth, n = 12, 1 << 12
    pre = synth.indexed_preimages(n, n - 100, seed=5)
    m = imt_b200.Multi(_devices(world))
    assert m.size == world
    t = m.build_from_leaves(pre)
    assert t.depth == depth and t.num_leaves == n
    want = O.tree_build(O.hash3(pre, 4), 4)
    assert np.array_equal(t.root(), want[-1])
    whole = eng.build_from_leaves(pre)
    idx = np.array([0, 1, n // world - 1, n // world % n, n // 2 + 3, n - 1, 17, 17], np.uint64)
    sib, hel = t.get_proofs(idx)
    wsib, whel = whole.get_proofs(idx)
    assert np.array_equal(sib, wsib) and np.array_equal(hel, whel)
    lv, lg = t.leaves(idx)
    wl, wg = whole.leaves(idx)
    assert np.array_equal(lv, wl) and np.array_equal(lg, wg)
    with pytest.raises(imt_b200.ImtError) as ei:
        t.get_proofs(np.array([n], np.uint64))
    assert ei.value.status == _ffi.ERR_INDEX_OOB
    # rebuild with other leaves: the cap follows
    pre2 = synth.indexed_preimages(n, n // 2, seed=6)
    t.rebuild_from_leaves(pre2)
    assert np.array_equal(t.root(), eng.build_from_leaves(pre2).root())
    m.close()


@pytest.mark.parametrize("world", [2, 4])
def test_multi_lookups_and_inserts_equal_single_gpu(eng, world):
    n, occ = 1 << 10, 600
    pre = synth.indexed_preimages(n, occ, seed=9)
    m = imt_b200.Multi(_devices(world))
    t = m.build_from_leaves(pre)
    whole = eng.build_from_leaves(pre)
    assert t.occupied == occ == whole.occupied
    vals = np.concatenate([synth.field_elements(300, seed=77), pre[1:40, 0], np.zeros((1, 4), np.uint64)])   # absent, present, zero
    low, matched = t.low_leaf_lookup(vals)
    wlow, wmatched = whole.low_leaf_lookup(vals)
    assert np.array_equal(low, wlow) and np.array_equal(matched, wmatched)
    for k in (0, 5, 299, 310, 339):                               # the oracle's literal linear scan (IMT:632-660)
        assert (int(low[k]), bool(matched[k])) == O.low_leaf(pre, vals[k])
    o, wo = t.non_inclusion_paths(vals[:300]), whole.non_inclusion_paths(vals[:300])
    for k in ("low_idx", "matched", "low_leaves", "siblings", "helpers", "is_largest"):
        assert np.array_equal(o[k], wo[k]), k
    new = synth.field_elements(200, seed=1234)
    w, ww = t.insert_batch(new), whole.insert_batch(new)
    assert set(ww) - set(w) == {"fold_nodes"}                      # the chain values of the folds come from single-GPU batches only
    for k in w:
        assert np.array_equal(w[k], ww[k]), k
    assert np.array_equal(t.root(), whole.root()) and t.occupied == occ + 200
    low, _ = t.low_leaf_lookup(vals)                               # the per-shard indices were merged, not rebuilt
    wlow, _ = whole.low_leaf_lookup(vals)
    assert np.array_equal(low, wlow)
    with pytest.raises(imt_b200.ImtError) as ei:                  # a repeated value: nothing is modified
        t.insert_batch(np.concatenate([new[:1], synth.field_elements(3, seed=4321)]))
    assert ei.value.status == _ffi.ERR_INVALID_ARG
    assert np.array_equal(t.root(), whole.root())
    with pytest.raises(imt_b200.ImtError) as ei:
        t.insert_batch(synth.field_elements(n, seed=99))
    assert ei.value.status == _ffi.ERR_TREE_FULL
    m.close()


def test_multi_witness_traces_equal_single_gpu_and_oracle(eng):
    n = 1 << 8
    pre = synth.indexed_preimages(n, n - 3, seed=21)
    m = imt_b200.Multi(_devices(4))
    t = m.build_from_leaves(pre)
    whole = eng.build_from_leaves(pre)
    idx = np.array([5, 250, 64, 65, 128, 0, 255, 5], np.uint64)
    st = t.trace_proofs(idx)
    assert np.array_equal(st, whole.trace_proofs(idx))
    levels = O.tree_build(O.hash3(pre, 4), 4)
    sib, _ = O.get_proof(levels, n, 250)
    h = O.hash3(pre[250:251], 1)[0]
    ix = 250
    for lvl in range(8):
        pair = np.stack([h, sib[lvl]]) if ix % 2 == 0 else np.stack([sib[lvl], h])
        h, states = O.hash_trace(pair)
        assert np.array_equal(st[1, lvl], states)
        ix //= 2
    m.close()


def _build_driver():
    _ffi.load()
    lib = _ffi.library_path()
    out = os.path.join(ROOT, "tests", "_build", "cabi_multi_driver")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cabi_multi_driver.c"), "-o", out, "-L", os.path.dirname(lib), "-limt_b200",
                    f"-Wl,-rpath,{os.path.dirname(lib)}"], check=True, env={k: v for k, v in os.environ.items() if k not in ("CC", "CXX")})
    return out


@pytest.mark.parametrize("world", [2, 4])
def test_c_client_builds_the_sharded_depth16_tree_in_one_process(world):
    exe = _build_driver()
    r = subprocess.run([exe, "multi", str(world), "16"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    lines = r.stdout.strip().splitlines()
    root = [l for l in lines if l.startswith("root ")][0]
    assert int(root.split()[1], 16) == int(GOLD["build_roots"]["16"]["random"])
    assert lines[-1] == "sharded == single-GPU: root 1 paths 1 lookups 1 inserts 1"


def test_c_client_one_process_per_gpu_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs: NCCL refuses two ranks on one device (run by tools/multi_gpu_check.sh under gpurun --gpus 2)")
    exe = _build_driver()
    with tempfile.TemporaryDirectory() as d:
        idfile = os.path.join(d, "nccl.id")
        procs = [subprocess.Popen([exe, "rank", str(r), "2", idfile, "16"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
        outs = [p.communicate(timeout=300) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, (so, se)
    root = [l for l in outs[0][0].splitlines() if l.startswith("root ")][0]
    assert int(root.split()[1], 16) == int(GOLD["build_roots"]["16"]["random"])
    assert outs[0][0].strip().splitlines()[-1] == "sharded == single-GPU: root 1 paths 1 lookups 1 inserts 1"
    assert [l for l in outs[1][0].splitlines() if l.startswith("root ")][0] == root

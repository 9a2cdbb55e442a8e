"""world_size-2 (and 4) `gloo` tests of the N>1 path on CPU: sharding.ShardedTree's host logic — root all-gather + cap,
owner-routed path / preimage queries, the candidate all-gather of sharded low-leaf lookups — driven with an
oracle-backed engine double (tests/_oracle_double.py) and compared with the unsharded oracle tree."""
import os
import random
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, depth, occupied, use_gpu, errors):
    try:
        sys.path[:0] = [os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "oracle"), HERE]
        import torch.distributed as dist
        import oracle as O
        import imt_b200
        from imt_b200 import synth, ShardedTree
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        if use_gpu:
            eng = imt_b200.Engine(0, "canonical")
        else:
            from _oracle_double import OracleEngine
            eng = OracleEngine()
        n = 1 << depth
        per = n // world
        pre = synth.indexed_preimages(n, occupied, seed=depth * 31 + world)
        whole = O.tree_build(O.hash3(pre, 4), 4)
        st = ShardedTree(eng, pre[rank * per:(rank + 1) * per])
        assert st.depth == depth and st.num_leaves == n
        assert np.array_equal(st.root(), whole[-1]), "sharded root differs from the single tree"
        rng = random.Random(5)
        idx = np.array([0, 1, per - 1, per, n - 1] + [rng.randrange(n) for _ in range(40)], np.uint64)
        sib, hel = st.get_proofs(idx)
        for k, i in enumerate(idx):
            s, h = O.get_proof(whole, n, int(i))
            assert np.array_equal(sib[k], s) and np.array_equal(hel[k], h), f"path {int(i)}"
        lv, lg = st.leaves(idx)
        assert np.array_equal(lv, pre[idx.astype(np.int64)])
        vals = [O.to_int(pre[i, 0]) for i in range(1, occupied)]
        qs = [0, 1, imt_b200.P - 1] + [rng.randrange(imt_b200.P) for _ in range(60)] + (rng.sample(vals, min(5, len(vals))) if vals else [])
        o = st.non_inclusion_paths(O.fes(qs))
        for k, v in enumerate(qs):
            wl, wm = O.low_leaf(pre, O.fe(v))
            assert (int(o["low_idx"][k]), bool(o["matched"][k])) == (wl, wm), f"low leaf of {v}"
            s, h = O.get_proof(whole, n, wl)
            assert np.array_equal(o["siblings"][k], s) and np.array_equal(o["helpers"][k], h)
            assert np.array_equal(o["low_leaves"][k], pre[wl]) and o["is_largest"][k] == (0 if pre[wl, 1].any() else 1)
        # by-query sharding of folds: every rank folds its slice, the roots all equal the tree root
        leaf_hashes = O.hash3(pre[idx.astype(np.int64)], 4)
        sl, roots, _ = st.trace_merkle_proofs(leaf_hashes, idx, sib, want_states=False)
        assert (roots == whole[-1]).all() and sl == st.query_slice(len(idx))
        # by-owner sharding of the path traces: each rank traces the leaves it owns from its stored levels + the cap
        mine, states = st.trace_proofs(idx)
        assert sorted(mine.tolist()) == [k for k, i in enumerate(idx) if int(i) // per == rank]
        for k, tr in zip(mine[:6], states[:6]):
            h, ix = O.hash3(pre[int(idx[k])][None], 1)[0], int(idx[k])
            for lvl in range(depth):
                pair = np.stack([h, sib[k, lvl]]) if ix % 2 == 0 else np.stack([sib[k, lvl], h])
                h, ws = O.hash_trace(pair)
                assert np.array_equal(tr[lvl], ws), f"trace of leaf {int(idx[k])} level {lvl}"
                ix //= 2
        if True:
            # sharded insert batch: against the oracle advancing the WHOLE tree one insert at a time
            free = n - occupied
            b = min(free - 2, 70 if use_gpu else 12)
            if b > 0:
                ins_vals = synth.field_elements(b, seed=depth * 7 + world)
                got = st.insert_batch(ins_vals, chunk=32 if use_gpu else 5)                  # several chunks, new slots crossing rank boundaries
                ost = O.InsertState(pre, threads=2)
                names = [("old_roots", "old_root"), ("low_idx", "low_idx"), ("low_leaves", "low_leaf"), ("low_siblings", "low_proof"),
                         ("low_helpers", "low_helper"), ("new_roots", "new_root"), ("new_leaves", "new_leaf"), ("new_siblings", "new_proof"),
                         ("new_helpers", "new_helper"), ("is_largest", "is_largest")]
                for k in range(b):
                    want = ost.insert(ins_vals[k], occupied + k, incremental=True)
                    for g, w in names:
                        assert np.array_equal(np.asarray(got[g][k]), np.asarray(want[w])), f"sharded insert {k}: {g} got {np.asarray(got[g][k]).tolist()} want {np.asarray(want[w]).tolist()}"
                assert np.array_equal(st.root(), ost.root())
                got_pre = st.tree.preimages(per) if use_gpu else st.tree.pre
                assert np.array_equal(got_pre, ost.pre[rank * per:(rank + 1) * per])
                qs2 = O.fes([rng.randrange(imt_b200.P) for _ in range(30)])
                lo2, ma2 = st.low_leaf_lookup(qs2)                        # the per-rank indexes absorbed the new keys
                for k in range(30):
                    assert (int(lo2[k]), bool(ma2[k])) == O.low_leaf(ost.pre, qs2[k])
                sib2, hel2 = st.get_proofs(idx)
                for k, i in enumerate(idx):
                    s2, h2 = O.get_proof(ost.tree, n, int(i))
                    assert np.array_equal(sib2[k], s2)
                with pytest.raises(imt_b200.ImtError if use_gpu else ValueError):
                    st.insert_batch(ins_vals[:1])                          # already present
        # a rebuild makes the cap stale until the roots are exchanged again
        pre2 = synth.indexed_preimages(n, max(1, occupied // 2), seed=99)
        st.rebuild(pre2[rank * per:(rank + 1) * per])
        assert np.array_equal(st.root(), O.tree_build(O.hash3(pre2, 4), 4)[-1])
        dist.barrier()
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        import traceback
        errors.put((rank, traceback.format_exc()))
        raise


def _run(world, depth, occupied, use_gpu=False):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    errors = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, depth, occupied, use_gpu, errors)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    msgs = []
    while not errors.empty():
        msgs.append(errors.get())
    for p in procs:
        if p.is_alive():
            p.kill()
            msgs.append((-1, "timeout"))
    assert not msgs and all(p.exitcode == 0 for p in procs), "\n".join(f"rank {r}: {m}" for r, m in msgs)


@pytest.mark.parametrize("world,depth,occupied", [(2, 6, 40), (2, 5, 1), (4, 6, 64), (4, 6, 9)])
def test_sharded_tree_host_logic_gloo(world, depth, occupied):
    _run(world, depth, occupied)


@pytest.mark.gpu
@pytest.mark.parametrize("world,depth,occupied", [(2, 10, 700), (4, 8, 40), (4, 7, 60), (2, 6, 1)])
def test_sharded_tree_real_engine_gloo(world, depth, occupied):
    """the same scenario with the real C-ABI library on cuda:0 in every process (ranks share the one GPU of the box)"""
    _run(world, depth, occupied, use_gpu=True)

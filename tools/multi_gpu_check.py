#!/usr/bin/env python
"""NCCL check of the sharded path, run under torchrun on an N-GPU box:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py [depth]
Every rank builds its subtree on its own GPU; every exchange (the root all-gather, the lookup candidates, the insert rounds) is
issued by libimt_b200.so's own NCCL communicator (csrc/imt_comm.cu) — torch.distributed only launches the ranks and carries the
128-byte NCCL id. The sharded tree is compared with (a) the single-GPU tree of the same leaves built on this rank's GPU and
(b) the CPU oracle."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import imt_b200  # noqa: E402
from imt_b200 import synth, ShardedTree  # noqa: E402
import oracle as O  # noqa: E402


def main():
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = imt_b200.Engine(local, "canonical")
    n = 1 << depth
    per = n // world
    occupied = n - n // 8
    pre = synth.indexed_preimages(n, occupied, seed=depth)
    st = ShardedTree(eng, pre[rank * per:(rank + 1) * per])
    whole = eng.build_from_leaves(pre)                      # the same tree, unsharded, on this rank's GPU
    assert np.array_equal(st.root(), whole.root()), "sharded root != single-GPU root"
    assert np.array_equal(st.root(), O.build_from_preimages(pre, O.max_threads())), "root != CPU oracle"
    rng = random.Random(7)
    idx = np.array([0, per - 1, per % n, n - 1] + [rng.randrange(n) for _ in range(2000)], np.uint64)
    sib, hel = st.get_proofs(idx)
    wsib, whel = whole.get_proofs(idx)
    assert np.array_equal(sib, wsib) and np.array_equal(hel, whel), "sharded paths differ"
    qs = synth.field_elements(5000, seed=77)
    o = st.non_inclusion_paths(qs)
    w = whole.non_inclusion_paths(qs)
    for k in ("low_idx", "low_leaves", "siblings", "helpers", "is_largest"):
        assert np.array_equal(o[k], w[k]), k
    assert np.array_equal(o["matched"], w["matched"].astype(bool))
    leaf_hashes = eng.hash3(pre[idx.astype(np.int64)])
    sl, roots, states = st.trace_merkle_proofs(leaf_hashes, idx, sib)
    assert (roots == st.root()).all() and states.shape[1:] == (depth, 132, 3, 4)
    # traces sharded by leaf owner (operands read from the stored levels + the cap) == the serial fold traces of those queries
    mine, own_states = st.trace_proofs(idx)
    _, want_states = eng.trace_merkle_proofs(leaf_hashes[mine], idx[mine], sib[mine])
    assert np.array_equal(own_states, want_states), "owner-sharded traces differ from the fold traces"
    # sharded insert batch == the single-GPU insert batch of the same tree, field by field
    vals = synth.field_elements(min(3000, n - occupied), seed=99)
    got = st.insert_batch(vals)
    want = whole.insert_batch(vals)
    for k in want:
        if k == "fold_nodes":   # the chain values for the one-launch witness trace come from single-GPU batches only
            continue
        assert np.array_equal(np.asarray(got[k]), np.asarray(want[k])), f"sharded insert: {k}"
    assert np.array_equal(st.root(), whole.root())
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU check ok: world={world} depth={depth} root={imt_b200.fe_to_int(st.root()):#x} "
              f"(library communicator: rank/world/NCCL {eng.comm_info()})", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

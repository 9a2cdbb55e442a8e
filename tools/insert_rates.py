#!/usr/bin/env python
"""Insert throughput of imt_insert_batch at depth 24 for one batch size (env B, default 131072) and one internal chunk size
(env IMT_INSERT_CHUNK, default = the library's): per-batch wall time through the host API into reused page-locked witness
buffers, inserts/s, and the low digits of the final root (must not depend on the chunking). Run on the GPU box:
   for c in 4096 32768 65536; do IMT_INSERT_CHUNK=$c python tools/insert_rates.py; done"""
import os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, torch
import imt_b200
from imt_b200 import synth
depth = 24; n = 1 << depth; b = int(os.environ.get("B", "131072"))
eng = imt_b200.Engine(0, "canonical")
dev = torch.device("cuda", 0)
d_pre = synth.indexed_preimages_torch(n, n - 8 * 131072, device=dev)
tree = eng.build_from_leaves_dev(d_pre, n)
out = tree.insert_buffers(b, depth, pinned=True)
slot = tree.occupied
ts = []
for s in range(4):
    vals = synth.field_elements(b, seed=7000 + s)
    t0 = time.perf_counter(); tree.insert_batch(vals, first_idx=slot, out=out); ts.append(time.perf_counter() - t0); slot += b
print("chunk", os.environ.get("IMT_INSERT_CHUNK", "default"), "batch", b, "ms", [round(t * 1e3, 2) for t in ts], "inserts/s", round(b / min(ts[1:])), "root", imt_b200.fe_to_int(tree.root()) % 10**12, flush=True)
# TRACE=<inserts>: the insert_leaf witness trace (imt_insert_witness_trace_dev) of the first <inserts> inserts of the last batch, device
# resident, with the chain values of the folds (one launch) and without (1 + depth dependent launches)
tb = min(b, int(os.environ.get("TRACE", "0")))
if tb:
    w = out
    keys = ("low_idx", "low_leaves", "low_siblings", "new_leaves", "new_siblings", "fold_nodes")
    dw = {k: torch.from_numpy(np.ascontiguousarray(w[k][:tb]).view(np.int64)).to(dev) for k in keys}
    S = 3 + 4 * depth
    d_states = torch.empty((tb, S, 132, 3, 4), dtype=torch.int64, device=dev)
    d_roots = torch.empty((tb, 4, 4), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    for name, d in (("one launch", dw), ("level loop", {k: v for k, v in dw.items() if k != "fold_nodes"})):
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); eng.trace_insert_witness_dev(d, tb, depth, slot - b, d_states, d_roots); best = min(best, time.perf_counter() - t0)
        assert np.array_equal(d_roots[:, 3].cpu().numpy().view(np.uint64), w["new_roots"][:tb])
        print(f"witness trace of {tb} inserts ({tb * S} traced hashes), {name}: {best * 1e3:.2f} ms = {tb * S / best / 1e6:.1f} M traced hashes/s", flush=True)

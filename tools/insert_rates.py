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

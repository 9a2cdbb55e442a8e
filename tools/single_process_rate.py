#!/usr/bin/env python
"""The shape a Rust host uses: ONE process, ONE call, N GPUs (imt_multi_create -> ncclCommInitAll inside the library).
   python tools/single_process_rate.py [N] [depth]        on an N-GPU box
Builds the depth-24 tree of bench.py's synthetic stream from HOST leaves (page-locked, then plain pageable memory) through
imt_multi_build_from_leaves / imt_mtree_rebuild_from_leaves, wall clock around the call (it returns with the root ready),
and checks the root against the golden one (tests/golden/golden.json bench_roots). One JSON line."""
import json, os, sys, time
ROOT = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import imt_b200
from imt_b200 import synth

world = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 24
n = 1 << depth
per = n // world
h_pre = torch.empty((n, 3, 4), dtype=torch.int64, pin_memory=True)
for r in range(world):                                                  # generated on the devices, as bench.py does per rank
    dev = torch.device("cuda", r)
    h_pre[r * per:(r + 1) * per].copy_(synth.field_elements_torch(3 * per, synth.DEFAULT_SEED, first=3 * per * r, device=dev).view(per, 3, 4))
    torch.cuda.synchronize(dev)
    torch.cuda.empty_cache()
t0 = time.perf_counter()
m = imt_b200.Multi(list(range(world)), "montgomery")                   # contexts + ncclCommInitAll
t_create = time.perf_counter() - t0
t0 = time.perf_counter()
tree = m.build_from_leaves_ptr(h_pre.data_ptr(), n)                     # IndexedMerkleTree::new over N GPUs: allocation + build (+ one-time lazy init)
t_first = time.perf_counter() - t0
tree.close()
t0 = time.perf_counter()
tree = m.build_from_leaves_ptr(h_pre.data_ptr(), n)                     # the same again: buffers recycled through the stream-ordered pools
t_second = time.perf_counter() - t0
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json"))).get("bench_roots", {}).get(str(depth))
root = tree.root()
assert gold is None or imt_b200.fe_to_int(root) == int(gold, 16), "root differs from the golden root"


def best(fn, k=4):
    ts = []
    for _ in range(k):
        t0 = time.perf_counter(); fn(); tree.root(); ts.append(time.perf_counter() - t0)
    return min(ts[1:]), ts


t_pinned, _ = best(lambda: tree.rebuild_from_leaves_ptr(h_pre.data_ptr()))
pageable = np.empty((n, 3, 4), np.uint64)                               # a Rust Vec<F>
pageable[...] = h_pre.numpy().view(np.uint64)
t_page, _ = best(lambda: tree.rebuild_from_leaves_ptr(pageable.ctypes.data))
assert np.array_equal(tree.root(), root)
hashes = 2 * n - 1
print(json.dumps({"call": "imt_mtree_rebuild_from_leaves (one process, one call)", "n_gpus": world, "depth": depth, "nccl": m.nccl_version,
                  "multi_create_ms": t_create * 1e3, "first_build_with_alloc_ms": t_first * 1e3, "second_build_with_alloc_ms": t_second * 1e3, "pinned_host_ms": t_pinned * 1e3, "pinned_hashes_per_s": hashes / t_pinned,
                  "pageable_host_ms": t_page * 1e3, "pageable_hashes_per_s": hashes / t_page, "root": f"{imt_b200.fe_to_int(root):064x}",
                  "root_check": "golden" if gold else "none"}), flush=True)

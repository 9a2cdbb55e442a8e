"""A/B of the latency kernels on a B200: microseconds per batch of n two-input hashes (one tree level near the root) through
imt_poseidon_hash2_dev, and whole builds at depth 16 / 20, for the kernel the library picks. Run twice:
    python tools/latency_ab.py                     # lead / helper kernel up to one block per SM, 3 lanes per hash above
    IMT_LH_MAX_NODES=0 python tools/latency_ab.py  # 3 lanes per hash only (round-2 state before the lead / helper kernel)
Digests of the two runs must be identical (printed as a checksum)."""
import hashlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import imt_b200
from imt_b200 import synth


def main():
    eng = imt_b200.Engine(0, "montgomery")
    torch.cuda.synchronize()
    sizes = [1, 8, 12, 13, 64, 256, 512, 888, 1024, 1776, 1777, 2048, 4096, 8192]
    d_in = synth.field_elements_torch(2 * max(sizes), seed=7, device="cuda")
    d_out = torch.zeros((max(sizes), 4), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    digest = hashlib.sha256()
    print("kernel bound IMT_LH_MAX_NODES =", os.environ.get("IMT_LH_MAX_NODES", "(default: 12 per SM)"))
    for n in sizes:
        for _ in range(3):
            eng.hash2_dev(d_in, n, d_out)
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.hash2_dev(d_in, n, d_out)   # synchronous on return
        dt = (time.perf_counter() - t0) / reps
        digest.update(d_out[:n].cpu().numpy().tobytes())
        print(f"hash2 x {n:5d}: {dt * 1e6:8.1f} us per call (wall clock, call + kernel + sync)")
    for depth in (10, 16, 20):
        n = 1 << depth
        d_pre = synth.field_elements_torch(3 * n, seed=depth, device="cuda").view(n, 3, 4)
        torch.cuda.synchronize()
        t = eng.build_from_leaves_dev(d_pre, n)
        for _ in range(2):
            t.rebuild_from_leaves_dev(d_pre)
        reps = 10
        t0 = time.perf_counter()
        for _ in range(reps):
            t.rebuild_from_leaves_dev(d_pre)
        dt = (time.perf_counter() - t0) / reps
        digest.update(np.asarray(t.root()).tobytes())
        print(f"depth-{depth} build: {dt * 1e3:8.3f} ms")
        t.close()
    print("checksum of all digests and roots:", digest.hexdigest()[:16])


if __name__ == "__main__":
    main()

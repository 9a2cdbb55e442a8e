#!/usr/bin/env python
"""Throughput of the any-width Poseidon kernels (imt_spec.cu) next to the tuned <3, 2>(8, 57) kernels: hashes/s and
permutations/s for 2^20 hashes of 2 inputs, leaves resident in HBM, CUDA events on the launching stream. One JSON line per
instance. Run on the GPU box: python tools/spec_rates.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import imt_b200
    from imt_b200 import synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n = 1 << 20
    d_in = synth.field_elements_torch(2 * n, synth.DEFAULT_SEED, device=dev).view(n, 2, 4)
    d_out = torch.empty((n, 4), dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev)
    rows = []
    for name, kw in (("tuned <3,2>(8,57)", dict()), ("generic <3,2>(8,57)", dict(generic=True)), ("generic <2,1>(8,56)", dict(t=2, rate=1, r_p=56)),
                     ("generic <4,3>(8,56)", dict(t=4, rate=3, r_p=56)), ("generic <5,4>(8,60)", dict(t=5, rate=4, r_p=60))):
        eng = imt_b200.Engine(0, "montgomery", **kw)
        eng.set_stream(stream.cuda_stream)
        for _ in range(2):
            eng.hash_dev(d_in, 2, n, d_out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            eng.hash_dev(d_in, 2, n, d_out)
        e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1) / 3
        perms = 2 // eng.rate + 1
        rows.append({"instance": name, "hashes_per_s": n / (ms * 1e-3), "perms_per_hash": perms, "perms_per_s": n * perms / (ms * 1e-3), "ms": ms})
        print(json.dumps(rows[-1]), flush=True)
        eng.close()


if __name__ == "__main__":
    main()

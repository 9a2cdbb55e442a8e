"""Regenerates tests/golden/golden.json from the ORACLE (oracle/imt_oracle.c + oracle/poseidon_ref.py).

The reference itself cannot run here (Rust, un-vendored deps, no toolchain), so these vectors come from the oracle,
which is pinned to the reference's single numeric known-answer (indexed_merkle_tree.rs:247-251) and cross-checked
between two independent implementations (tests/test_oracle.py). Usage: python tests/golden/make_golden.py
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402
import oracle as O  # noqa: E402
import poseidon_ref as R  # noqa: E402
import imt_b200  # noqa: E402  (synthetic-input definitions only; no GPU is touched)
from imt_b200 import synth  # noqa: E402


def spec_vectors():
    """any-width instances: published permutation vectors + sponge digests from the Python oracle"""
    out = {"published_perm_x5_254_3_input_0_1_2": [hex(v) for v in R.permute([0, 1, 2])],
           "published_perm_x5_254_5_input_0_to_4": [hex(v) for v in R.permute([0, 1, 2, 3, 4], R.Spec(8, 60, 5))], "hashes": {}}
    for t, r_f, r_p in ((2, 8, 56), (3, 8, 57), (4, 8, 56), (5, 8, 60)):
        sp = R.Spec(r_f, r_p, t)
        out["hashes"][f"{t},{r_f},{r_p}"] = {str(k): str(R.hash_n(list(range(1, k + 1)), sp)) for k in range(0, 7)}
    return out


def bench_root(depth, th):
    """Root of bench.py's headline input: the synthetic stream fed to a MONTGOMERY-format context, i.e. the same bit patterns
    read as x * 2^256 mod p. Returned as bench.py prints it: the root's Montgomery-form bytes as one big-endian hex number."""
    pre = O.convert(synth.random_preimages(1 << depth), False, th)        # the canonical values those bit patterns stand for
    root = O.convert(O.build_from_preimages(pre, th).reshape(1, 4), True)  # back to the in-memory form
    return f"{O.to_int(root):064x}"


def main():
    if "--bench-roots" in sys.argv:  # refresh only bench.py's expected roots (depth 24: ~6 min on 8 cores), keep the rest
        path = os.path.join(HERE, "golden.json")
        g = json.load(open(path))
        th = O.max_threads()
        br = g.get("bench_roots", {})
        for depth in (10, 16, 20) + ((24,) if "--depth24" in sys.argv else ()):
            t0 = time.time()
            br[str(depth)] = bench_root(depth, th)
            print(f"bench root depth {depth}: {br[str(depth)]} ({time.time() - t0:.1f}s)", flush=True)
        g["bench_roots"] = br
        with open(path, "w") as f:
            json.dump(g, f, indent=1)
        print("updated golden.json: bench_roots")
        return
    if "--spec-only" in sys.argv:  # refresh only the any-width vectors, keep everything else byte for byte
        path = os.path.join(HERE, "golden.json")
        g = json.load(open(path))
        g["spec"] = spec_vectors()
        with open(path, "w") as f:
            json.dump(g, f, indent=1)
        print("updated golden.json: spec")
        return
    th = O.max_threads()
    g = {"seed": synth.DEFAULT_SEED, "kat_h3_zero": str(R.KAT_H3_ZERO)}
    g["h2_0_0"] = str(R.hash2(0, 0))
    g["h2_1_2"] = str(R.hash2(1, 2))
    g["h3_1_2_3"] = str(R.hash3(1, 2, 3))
    g["empty_depth3_root"] = str(R.IndexedMerkleTree(R.hash_preimages([[0, 0, 0]] * 8)).root)
    rounds, pre = R.insert_rounds(3, [30, 10, 20, 5, 50, 35])
    g["scenario_inserts"] = [30, 10, 20, 5, 50, 35]
    g["scenario_low_idx"] = [r["low_idx"] for r in rounds]
    g["scenario_roots"] = [str(r["new_root"]) for r in rounds]
    g["scenario_final_preimages"] = [[str(v) for v in leaf] for leaf in pre]
    # first and last state of the H3(0,0,0) trace and a checksum of all 132
    d, tr = R.hash_trace([0, 0, 0])
    g["trace_h3_zero_first"] = [str(v) for v in tr[0]]
    g["trace_h3_zero_last"] = [str(v) for v in tr[-1]]
    g["trace_h3_zero_xor"] = str(int(np.bitwise_xor.reduce([v for s in tr for v in s])) if False else __import__("functools").reduce(lambda a, b: a ^ b, [v for s in tr for v in s]))
    roots = {}
    # depth 24 (the headline size) takes ~6 min per tree on 8 cores: pass --depth24 to regenerate it, else the stored value is kept
    old = {}
    try:
        old = json.load(open(os.path.join(HERE, "golden.json"))).get("build_roots", {})
    except Exception:
        pass
    for depth in (3, 10, 16, 20) + ((24,) if "--depth24" in sys.argv else ()):
        n = 1 << depth
        t0 = time.time()
        r1 = O.build_from_preimages(synth.random_preimages(n), th)
        r2 = O.build_from_preimages(synth.indexed_preimages(n), th)
        roots[str(depth)] = {"random": str(O.to_int(r1)), "indexed": str(O.to_int(r2))}
        print(f"depth {depth}: {time.time() - t0:.1f}s", flush=True)
    if "24" not in roots and "24" in old:
        roots["24"] = old["24"]
    g["build_roots"] = roots
    g["spec"] = spec_vectors()
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote golden.json")


if __name__ == "__main__":
    main()

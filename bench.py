#!/usr/bin/env python
"""bench.py — Poseidon hashes/s building the depth-24 indexed Merkle tree (BASELINE.json's metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--depth D] [--impl reference]
  N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one full build (leaf hashing H3(val,next_val,next_idx) + every level) of the depth-D tree over synthetic
leaves. With N ranks the tree is sharded by subtree: each rank builds 2^D / N leaves, the N subtree roots cross
NVLink in one NCCL all-gather, and every rank builds the log2(N) cap levels (strong scaling: D is fixed).
Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for how each field is obtained.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MACS_PER_HASH = 163_200          # 2 perms x 600 Montgomery muls x 136 32x32->64 MACs
EXECUTED_MACS_PER_HASH = 123_792  # what the kernels issue: squarings are 36 products, 3-term dot products reduce once (ncu: profiles/)
LEAF_DRAM_BYTES_PER_HASH = 129.3  # dram__bytes_read+write of k_hash<3> per leaf: (1.643 + 0.527) GB / 2^24 leaves, ncu --set full (profiles/r02_summary.md)
TRACE_DRAM_BYTES_PER_HASH = 12568.0  # dram__bytes_read+write of k_trace_tree_paths per traced hash: 4.118 GB / 327 680 hashes (profiles/r01e_summary.md); algorithmic 12 672 + 64
METRIC = "poseidon_hashes_per_s_depth24_tree_build"


def workload_name(depth):
    return f"depth-{depth} indexed Merkle tree build (leaf H3 + all levels), BN254 Poseidon T=3 R_F=8 R_P=57"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--depth", type=int, default=24)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="build", choices=["build", "paths", "lookups"],
                    help="build = the headline metric (default); paths = 2^16 path extraction + witness traces (BASELINE config 4); "
                         "lookups = 1M low-leaf lookups + non-inclusion paths + 4096 inserts (config 5, one GPU). "
                         "paths with --gpus N > 1 (torchrun) = the traces sharded by leaf owner over the N-GPU sharded tree.")
    ap.add_argument("--queries", type=int, default=0, help="query count of the paths / lookups workloads (default 2^16 / 2^20)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the paths / lookups / insert legs of the default run")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md's clocks line). The sampler runs
    from before the warm-up; every row is stamped on arrival and mark_begin()/mark_end() bracket the timed region. When
    the region is shorter than the sampling period (8-GPU steps last ~80 ms) the rows taken under the identical load of
    the warm-up steps just before it are used, and `window` says so."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.monotonic()

    def mark_end(self):
        self.t1 = time.monotonic()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        t0, t1 = self.t0 or 0.0, self.t1 or float("inf")
        inside = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.05]
        window = "timed region"
        if not inside:  # region shorter than the sampling period: the warm-up steps right before it ran the same load
            inside = [r for ts, r in self.rows if t0 - 1.0 <= ts <= t1 + 0.25]
            window = "timed region shorter than the sampling period: samples from the warm-up steps within 1 s before it"
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            try:
                sm.append(float(r[0])), mx.append(float(r[1])), pw.append(float(r[2]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons), "window": window}


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_build_rate(depth_hint_seconds, threads=None):
    """Times the oracle's restatement of the reference CPU path — leaf hashing (IMT:662-671) + level-by-level build
    (utils.rs:41-51) — on a bounded sample. Returns (hashes/s, threads, sample description)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import imt_b200
    from imt_b200 import synth
    th = threads or O.max_threads()
    d = 10
    pre = synth.random_preimages(1 << d)
    t0 = time.perf_counter()
    O.build_from_preimages(pre, th)
    dt = time.perf_counter() - t0
    rate = (2 * (1 << d) - 1) / dt
    # pick the depth whose build fits the budget
    d = max(10, min(20, int((rate * depth_hint_seconds / 2)).bit_length() - 1))
    pre = synth.random_preimages(1 << d)
    t0 = time.perf_counter()
    O.build_from_preimages(pre, th)
    dt = time.perf_counter() - t0
    hashes = 2 * (1 << d) - 1
    return hashes / dt, th, f"depth-{d} build ({hashes} hashes) in {dt:.2f}s, {th} threads", d, dt


def cpu_single_thread_rate(seconds=2.0):
    """The same port on ONE thread — the reference itself is strictly sequential (its hasher is a `&mut` borrow,
    utils.rs:21, 43-47). Returns (hashes/s, sample description)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from imt_b200 import synth
    d = 10
    pre = synth.random_preimages(1 << d)
    t0 = time.perf_counter()
    O.build_from_preimages(pre, 1)
    rate = (2 * (1 << d) - 1) / (time.perf_counter() - t0)
    d = max(10, min(18, int(rate * seconds / 2).bit_length() - 1))
    pre = synth.random_preimages(1 << d)
    t0 = time.perf_counter()
    O.build_from_preimages(pre, 1)
    dt = time.perf_counter() - t0
    hashes = 2 * (1 << d) - 1
    return hashes / dt, f"depth-{d} build ({hashes} hashes) in {dt:.2f}s, 1 thread"


def cpu_trace_rate(seconds=2.0):
    """the oracle's traced hash (132 x 3 states recorded) on one thread: the CPU side of the witness-trace legs"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle as O
    from imt_b200 import synth
    pairs = synth.field_elements(64).reshape(32, 2, 4)
    t0, cnt = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        for p_ in pairs:
            O.hash_trace(p_)
        cnt += len(pairs)
    return {"value": cnt / (time.perf_counter() - t0), "unit": "hashes/s", "cores": 1, "kind": "port", "sample": f"{cnt} traced hashes, oracle, 1 thread"}


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path. The Rust crate cannot be built here
    (no cargo/rustc, un-vendored git deps), so this is the oracle PORT of it, with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import imt_b200
    from imt_b200 import synth
    th = O.max_threads()
    # a step = a bounded sample of the depth-D workload: one depth-S build, S sized for ~3 s per step. Poseidon's cost is data-
    # and size-independent (every hash is the same 2 permutations), so hashes/s of the sample IS hashes/s of the full build;
    # ms_per_step is reported for the DECLARED workload (hashes_per_step / value) and flagged as extrapolated, the measured
    # sample time sits beside it.
    rate, _, _, _, _ = cpu_build_rate(1.0, th)
    S = max(10, min(a.depth, int(rate * 3.0 / 2).bit_length() - 1))
    pre = O.convert(synth.random_preimages(1 << S), False, th)   # the GPU arm's input: the same stream read as Montgomery values
    hashes = 2 * (1 << S) - 1
    for _ in range(a.warmup):
        O.build_from_preimages(pre, th)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        O.build_from_preimages(pre, th)
    dt = time.perf_counter() - t0
    v = hashes * a.steps / dt
    full_hashes = 2 * (1 << a.depth) - 1
    sample = f"each step = depth-{S} build ({hashes} hashes) of the depth-{a.depth} workload, {th} host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "hashes/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * full_hashes / v, "ms_per_step_extrapolated": S != a.depth, "sample_ms_per_step": 1e3 * dt / a.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32x8-montgomery(cpu: u64x4)",
        "data": "synthetic", "config": {"workload": workload_name(a.depth), "depth": a.depth, "leaves": 1 << a.depth, "hashes_per_step": full_hashes,
                                        "sharding": f"host threads x{th}", "seed": synth.DEFAULT_SEED, "fe_format": "montgomery",
                                        "sample_depth": S, "sample_hashes_per_step": hashes,
                                        "sample_note": f"timed on depth-{S} sample builds (hash cost is size-independent); ms_per_step is value-consistent for depth {a.depth}"},
        "cpu_baseline": {"value": v, "unit": "hashes/s", "cores": th, "kind": "port", "sample": sample,
                         "single_thread": dict(zip(("value", "sample"), cpu_single_thread_rate()))},
        "e2e": {"value": v, "unit": "hashes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C port of the reference's Rust CPU path (oracle/imt_oracle.c): the crate itself is unbuildable here",
    }), flush=True)


# ------------------------------------------------------------------------------------------- secondary workloads
def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def _ev_time(torch, stream, fn, steps, warmup):
    """CUDA-event time of `steps` calls of fn on `stream`, ms per call"""
    for _ in range(warmup):
        fn()
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    stream.synchronize()
    return e0.elapsed_time(e1) / steps


def run_paths(a):
    """BASELINE config 4: 2^16 Merkle paths out of the depth-D tree + the per-round Poseidon witness trace of
    verify_merkle_proof for each (D hashes x 132 x 3 FE per query). One JSON line."""
    import numpy as np
    import torch
    import imt_b200
    from imt_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    depth, q = a.depth, a.queries or (1 << 16)
    n = 1 << depth
    eng = imt_b200.Engine(0, "montgomery")
    stream = torch.cuda.current_stream(dev)
    eng.set_stream(stream.cuda_stream)
    d_pre = synth.field_elements_torch(3 * n, synth.DEFAULT_SEED, device=dev).view(n, 3, 4)
    tree = eng.build_from_leaves_dev(d_pre, n)
    g = torch.Generator(device=dev)
    g.manual_seed(synth.DEFAULT_SEED)
    d_idx = torch.randint(0, n, (q,), generator=g, device=dev, dtype=torch.int64)
    d_sib = torch.empty((q, depth, 4), dtype=torch.int64, device=dev)
    d_hel = torch.empty((q, depth), dtype=torch.uint8, device=dev)
    d_leaf = torch.empty((q, 4), dtype=torch.int64, device=dev)
    d_roots = torch.empty((q, 4), dtype=torch.int64, device=dev)
    d_states = torch.empty((q, depth, 132, 3, 4), dtype=torch.int64, device=dev)   # 19.9 GB at depth 24
    eng.hash3_dev(d_pre[d_idx].contiguous(), q, d_leaf)                           # leaf hashes of the queried slots

    t_gather = _ev_time(torch, stream, lambda: tree.get_proofs_dev(d_idx, q, d_sib, d_hel), a.steps, a.warmup)
    l0 = eng.launches
    t_trace = _ev_time(torch, stream, lambda: eng.trace_merkle_proofs_dev(d_leaf, d_idx, d_sib, q, depth, d_states, d_roots), a.steps, a.warmup)
    launches = eng.launches - l0
    t_fold = _ev_time(torch, stream, lambda: eng.trace_merkle_proofs_dev(d_leaf, d_idx, d_sib, q, depth, None, d_roots), a.steps, a.warmup)
    # the same traces straight from the resident tree: q x depth independent hashes instead of q serial folds
    chk = d_states[:: max(1, q // 64)].clone()
    d_states.zero_()
    l1 = eng.launches
    t_tree = _ev_time(torch, stream, lambda: tree.trace_proofs_dev(d_idx, q, d_states), a.steps, a.warmup)
    launches_tree = eng.launches - l1
    stream.synchronize()
    assert bool((d_states[:: max(1, q // 64)] == chk).all()), "tree trace differs from the fold trace"
    root = torch.empty(4, dtype=torch.int64, device=dev)
    tree.root_dev(root)
    stream.synchronize()
    assert bool((d_roots == root).all()), "a traced path does not fold to the tree root"
    hashes = q * depth
    trace_bytes = hashes * 132 * 96
    gather_bytes = q * depth * (32 + 32 + 1)
    peaks = _peaks()
    hbm = peaks.get("hbm_gbs", 6650.0)
    imad_rate, imad_mhz = eng.calibrate_imad(150.0)
    # e2e through the host API: HOST indices / leaves in, paths + traces out into PINNED host memory (a quarter of the
    # batch: 5 GB of traces; the full 19.9 GB would only make the allocation longer, the rate is the same)
    qe = min(q, 1 << 14)
    h_idx = d_idx[:qe].cpu().numpy().astype(np.uint64)
    h_leaf = d_leaf[:qe].cpu().numpy().view(np.uint64)
    h_states = torch.empty((qe, depth, 132, 3, 4), dtype=torch.int64, pin_memory=True)
    st_view = h_states.numpy().view(np.uint64)
    e2e_times = []
    for _ in range(2):
        t0 = time.perf_counter()
        sib, _ = tree.get_proofs(h_idx)
        tree.trace_proofs(h_idx, out_states=st_view)
        e2e_times.append(time.perf_counter() - t0)
    dt = min(e2e_times)
    assert bool((h_states[-1, -1, -1, 1] == root.cpu()).all()), "last traced state of the last path is not the root"
    # CPU leg: the oracle tracing the same folds, one thread (the reference's hasher is a &mut borrow)
    rinv = pow(1 << 256, -1, imt_b200.P)
    cs = min(qe, 48)
    lv = O.fes([x * rinv % imt_b200.P for x in O.to_ints(h_leaf[:cs])])
    sb = O.fes([x * rinv % imt_b200.P for x in O.to_ints(sib[:cs].reshape(-1, 4))]).reshape(cs, depth, 4)
    t0 = time.perf_counter()
    for k in range(cs):
        h, ix = lv[k], int(h_idx[k])
        for lvl in range(depth):
            h, _ = O.hash_trace(np.stack([h, sb[k, lvl]]) if ix % 2 == 0 else np.stack([sb[k, lvl], h]))
            ix //= 2
    cpu = cs * depth / (time.perf_counter() - t0)
    out = {
        "metric": "merkle_path_witness_trace_hashes_per_s", "value": hashes / (t_tree * 1e-3), "unit": "hashes/s", "n_gpus": 1, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": t_tree, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32x8-montgomery",
        "data": "synthetic",
        "config": {"workload": f"2^{q.bit_length() - 1} uniform random paths of the depth-{depth} tree: batched get_proof + verify_merkle_proof witness trace "
                               f"(132 x 3 FE per hash)", "depth": depth, "queries": q, "hashes_per_step": hashes, "trace_bytes_per_step": trace_bytes,
                   "l2_policy": f"trace output {trace_bytes / 2**30:.1f} GiB per step, far larger than L2", "seed": synth.DEFAULT_SEED, "fe_format": "montgomery"},
        "e2e": {"value": qe * depth / dt, "unit": "hashes/s", "sample": f"{qe} queries through the host API, traces into pinned host memory ({qe * depth * 132 * 96 / 2**30:.1f} GiB, PCIe-bound)",
                "d2h_gbs": qe * depth * 132 * 96 / dt / 1e9,
                "h2d_bytes_per_step": qe * (8 + 32 + depth * 32), "d2h_bytes_per_step": qe * depth * (32 + 1 + 132 * 96)},
        "gpu_launches": launches_tree,
        "roofline": {"bound": "imad", "kernel": "k_trace_tree_paths", "achieved": hashes * MACS_PER_HASH / (t_tree * 1e-3) / 1e9, "peak": imad_rate / 1e9,
                     "unit": "GMAC/s", "frac": hashes * MACS_PER_HASH / (t_tree * 1e-3) / imad_rate, "traffic": TRACE_DRAM_BYTES_PER_HASH * hashes,
                     "traffic_note": "bytes per launch = ncu dram__bytes_read.sum + dram__bytes_write.sum of this kernel per traced hash (12 568 B, one --set full "
                                     "capture, profiles/r01e_summary.md) x hashes per launch; algorithmic 12 672 B written + 64 B read per hash",
                     "peak_source": f"imt_calibrate_imad in this run ({imad_mhz:.0f} MHz implied)",
                     "hbm": {"achieved_gbs": trace_bytes / (t_tree * 1e-3) / 1e9, "peak_gbs": hbm, "frac": trace_bytes / (t_tree * 1e-3) / 1e9 / hbm}},
        "parts": {"get_proofs_ms": t_gather, "serial_fold_trace_ms": t_trace, "serial_fold_trace_hashes_per_s": hashes / (t_trace * 1e-3),
                  "serial_fold_trace_launches": launches, "get_proofs_gbs": gather_bytes / (t_gather * 1e-3) / 1e9, "get_proofs_paths_per_s": q / (t_gather * 1e-3),
                  "fold_without_trace_ms": t_fold, "fold_hashes_per_s": hashes / (t_fold * 1e-3)},
        "cpu_baseline": {"value": cpu, "unit": "hashes/s", "cores": 1, "kind": "port", "sample": f"{cs} paths x {depth} traced hashes, oracle, 1 thread"},
    }
    print(json.dumps(out), flush=True)


def run_lookups(a):
    """BASELINE config 5: 1M low-leaf (predecessor) lookups + non-inclusion paths on the depth-D indexed tree, then a
    batched insert of 4096 leaves with per-insert roots. One JSON line (metric: lookups/s)."""
    import numpy as np
    import torch
    import imt_b200
    from imt_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    depth, q, b = a.depth, a.queries or (1 << 20), 4096
    n = 1 << depth
    m = n - b * (a.steps + a.warmup + 1)
    eng = imt_b200.Engine(0, "canonical")
    stream = torch.cuda.current_stream(dev)
    eng.set_stream(stream.cuda_stream)
    d_pre = synth.indexed_preimages_torch(n, m, device=dev)
    tree = eng.build_from_leaves_dev(d_pre, n)
    t0 = time.perf_counter()
    assert tree.occupied == m                                                      # builds the sorted index (one merge sort of m keys)
    t_index = time.perf_counter() - t0
    d_vals = synth.field_elements_torch(q, seed=555, device=dev)
    d_low = torch.empty(q, dtype=torch.int64, device=dev)
    d_match = torch.empty(q, dtype=torch.uint8, device=dev)
    d_sib = torch.empty((q, depth, 4), dtype=torch.int64, device=dev)
    d_hel = torch.empty((q, depth), dtype=torch.uint8, device=dev)
    l0 = eng.launches
    t_lookup = _ev_time(torch, stream, lambda: tree.low_leaf_lookup_dev(d_vals, q, d_low, d_match), a.steps, a.warmup)
    launches = eng.launches - l0
    t_paths = _ev_time(torch, stream, lambda: tree.get_proofs_dev(d_low, q, d_sib, d_hel), a.steps, a.warmup)
    assert bool(d_match.all())
    # spot check against the oracle's linear scan
    h_pre = d_pre.cpu().numpy().view(np.uint64)
    h_vals = d_vals[:4].cpu().numpy().view(np.uint64)
    t0 = time.perf_counter()
    for k in range(4):
        assert (int(d_low[k]), True) == O.low_leaf(h_pre, h_vals[k])
    cpu_lookup = 4 / (time.perf_counter() - t0)
    # e2e: host values in, low_idx + witnesses out
    h_q = d_vals.cpu().numpy().view(np.uint64)
    t0 = time.perf_counter()
    o = tree.non_inclusion_paths(h_q)                                            # fresh pageable numpy outputs (page faults included)
    t_e2e_fresh = time.perf_counter() - t0
    assert np.array_equal(o["low_idx"], d_low.cpu().numpy().astype(np.uint64))
    bufs = tree.non_inclusion_buffers(q, depth, pinned=True)                      # reusable page-locked outputs: the steady-state call
    h_qp = torch.empty((q, 4), dtype=torch.int64, pin_memory=True)
    h_qp.numpy().view(np.uint64)[:] = h_q
    e2e_times = []
    for _ in range(1 + a.steps):
        t0 = time.perf_counter()
        o = tree.non_inclusion_paths(h_qp.numpy().view(np.uint64), out=bufs)
        e2e_times.append(time.perf_counter() - t0)
    t_e2e = min(e2e_times[1:])
    assert np.array_equal(o["low_idx"], d_low.cpu().numpy().astype(np.uint64))
    # inserts: each step one batch of 4096 (host API: values in, full witness bundle out)
    ins = []
    ins_out = tree.insert_buffers(b, depth, pinned=True)                          # reused page-locked witness buffers
    next_slot = tree.occupied
    for s in range(a.warmup + a.steps):
        vals = synth.field_elements(b, seed=9000 + s)
        t0 = time.perf_counter()
        w = tree.insert_batch(vals, first_idx=next_slot, out=ins_out)
        ins.append(time.perf_counter() - t0)
        next_slot += b
    t_ins = sum(ins[a.warmup:]) / a.steps
    assert np.array_equal(w["new_roots"][-1], tree.root())
    peaks = _peaks()
    hbm = peaks.get("hbm_gbs", 6650.0)
    probes = max(1, m.bit_length())
    out = {
        "metric": "low_leaf_lookups_per_s", "value": q / (t_lookup * 1e-3), "unit": "lookups/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": t_lookup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u256 compare", "data": "synthetic",
        "config": {"workload": f"{q} uniform random low-leaf lookups on the depth-{depth} indexed tree ({m} occupied slots), non-inclusion paths, "
                               f"then {b} inserts per batch with per-insert roots", "depth": depth, "queries": q, "occupied": m,
                   "l2_policy": f"sorted index {m * 36 / 2**20:.0f} MiB, random probes", "seed": synth.DEFAULT_SEED, "fe_format": "canonical"},
        "e2e": {"value": q / t_e2e, "unit": "lookups/s", "ms_per_step": t_e2e * 1e3, "h2d_bytes_per_step": q * 32,
                "d2h_bytes_per_step": q * (8 + 1 + 96 + depth * 33 + 1), "note": "imt_non_inclusion_paths: lookup + low leaf + path + is_largest per value, into reused page-locked host buffers",
                "ms_per_step_fresh_pageable_outputs": t_e2e_fresh * 1e3},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "k_low_leaf_lookup", "achieved": q * probes * 32 / (t_lookup * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": q * probes * 32 / (t_lookup * 1e-3) / 1e9 / hbm, "traffic": None,
                     "note": f"algorithmic bytes = {probes} dependent 32-byte probes per lookup; the top of the search tree stays in L2, so this is a latency-bound gather"},
        "parts": {"index_build_s": t_index, "non_inclusion_path_gather_ms": t_paths,
                  "path_gather_gbs": q * depth * 65 / (t_paths * 1e-3) / 1e9, "insert_batch_ms": t_ins * 1e3, "inserts_per_s": b / t_ins,
                  "insert_hashes_per_s": 2 * b * (depth + 1) / t_ins},
        "cpu_baseline": {"value": cpu_lookup, "unit": "lookups/s", "cores": 1, "kind": "port",
                         "sample": "4 lookups by the reference's linear scan (update_idx_leaf, IMT:632-660) over the same preimages, oracle"},
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------- secondary legs of the default run
def secondary_paths(torch, dist, eng, tree, d_pre, depth, n_local, world, rank, dev, stream, imad_rate, steps):
    """BASELINE config 4 inside the default run, sharded BY LEAF OWNER with no exchange: 2^16 uniform random GLOBAL leaf indices
    (the same on every rank); every rank traces the paths whose leaves it owns from its own stored levels + the replicated cap
    (imt_tree_trace_proofs_dev) and drains its traces over its own PCIe link. Aggregate over ranks, max time over ranks."""
    import numpy as np
    import imt_b200
    from imt_b200 import synth
    q = 1 << 16
    n_total = n_local * world
    g = torch.Generator(device=dev)
    g.manual_seed(synth.DEFAULT_SEED)
    idx = torch.randint(0, n_total, (q,), generator=g, device=dev, dtype=torch.int64)
    mine = idx[(idx // n_local) == rank].contiguous()
    qm = int(mine.numel())
    d_states = torch.empty((max(qm, 1), depth, 132, 3, 4), dtype=torch.int64, device=dev)

    def maxr(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    l0 = eng.launches
    ms = maxr(_ev_time(torch, stream, lambda: tree.trace_proofs_dev(mine, qm, d_states), steps, 2))
    launches = (eng.launches - l0) // (steps + 2)
    root = torch.empty(4, dtype=torch.int64, device=dev)
    tree.root_dev(root)
    stream.synchronize()
    assert qm == 0 or bool((d_states[:, -1, -1, 1] == root).all()), "a traced path does not end in the root"
    # e2e: host indices in, traces out into pinned host memory, every rank over its own link
    qe = min(qm, max(1, (1 << 14) // world))
    h_idx = mine[:qe].cpu().numpy().astype(np.uint64)
    h_states = torch.empty((qe, depth, 132, 3, 4), dtype=torch.int64, pin_memory=True)
    view = h_states.numpy().view(np.uint64)
    tree.trace_proofs(h_idx, out_states=view)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    tree.trace_proofs(h_idx, out_states=view)
    dt = maxr(time.perf_counter() - t0)
    assert bool((h_states[:, -1, -1, 1] == root.cpu()).all())
    qe_total = torch.tensor([qe], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(qe_total)
    qe_total = int(qe_total.item())
    # the fold-based trace (another kernel: leaf hash + get_proofs siblings, folded serially) must write the same bytes
    ks = min(qm, 64)
    if ks:
        sel = mine[:ks].contiguous()
        sib = torch.empty((ks, depth, 4), dtype=torch.int64, device=dev)
        leaf = torch.empty((ks, 4), dtype=torch.int64, device=dev)
        st2 = torch.empty((ks, depth, 132, 3, 4), dtype=torch.int64, device=dev)
        r2 = torch.empty((ks, 4), dtype=torch.int64, device=dev)
        tree.get_proofs_dev(sel, ks, sib)
        eng.hash3_dev(d_pre[sel - rank * n_local].contiguous(), ks, leaf)
        eng.trace_merkle_proofs_dev(leaf, sel, sib, ks, depth, st2, r2)
        stream.synchronize()
        assert bool((st2 == d_states[:ks]).all()) and bool((r2 == root).all()), "tree trace differs from the fold trace"
    hashes = q * depth
    peaks = _peaks()
    hbm = peaks.get("hbm_gbs", 6650.0)
    return {
        "metric": "merkle_path_witness_trace_hashes_per_s", "value": hashes / (ms * 1e-3), "unit": "hashes/s", "ms_per_step": ms, "n_gpus": world,
        "config": {"workload": f"2^16 uniform random paths of the depth-{depth} tree, verify_merkle_proof witness trace (132 x 3 FE per hash), "
                               f"sharded by leaf owner x{world}", "queries": q, "hashes_per_step": hashes, "trace_bytes_per_step": hashes * 132 * 96},
        "e2e": {"value": qe_total * depth / dt, "unit": "hashes/s", "d2h_gbs": qe_total * depth * 132 * 96 / dt / 1e9,
                "sample": f"{qe_total} queries through imt_tree_trace_proofs (host indices in, traces into pinned host memory, one PCIe link per GPU)",
                "h2d_bytes_per_step": qe_total * 8, "d2h_bytes_per_step": qe_total * depth * 132 * 96},
        "gpu_launches": launches * world,
        "roofline": {"bound": "imad", "kernel": "k_trace_tree_paths", "achieved": hashes * MACS_PER_HASH / (ms * 1e-3) / 1e9,
                     "peak": imad_rate * world / 1e9, "unit": "GMAC/s", "frac": hashes * MACS_PER_HASH / (ms * 1e-3) / (imad_rate * world),
                     "traffic": TRACE_DRAM_BYTES_PER_HASH * hashes / world, "traffic_source": "profiles/r01e_summary.md (ncu, per traced hash) x hashes per launch per GPU",
                     "hbm": {"achieved_gbs": hashes * 132 * 96 / (ms * 1e-3) / 1e9, "peak_gbs": hbm * world}},
    }


def secondary_lookups(torch, eng, depth, dev, stream, steps):
    """BASELINE config 5 inside the default run (one GPU): 2^20 low-leaf lookups on the depth-D indexed tree, their non-inclusion
    witnesses through the host API, then batches of 4096 inserts with per-insert roots and both paths."""
    import numpy as np
    from imt_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    q, b = 1 << 20, 4096
    n = 1 << depth
    m = n - b * (steps + 3)
    # the whole leg runs in the engine's own format — Montgomery in the default run: what a halo2 host holds its field elements in
    # (zero-copy), and traces then need no from_mont per traced state element. The oracle's scan gets canonical host copies.
    mont = eng.fmt == 1
    ec = eng

    def in_fmt(d_canonical):
        if not mont:
            return d_canonical
        out_ = torch.empty_like(d_canonical)
        eng.convert_dev(d_canonical, d_canonical.numel() // 4, out_, to_montgomery=True)
        return out_

    d_pre_c = synth.indexed_preimages_torch(n, m, device=dev)
    h_pre = d_pre_c.cpu().numpy().view(np.uint64)
    d_pre = in_fmt(d_pre_c)
    del d_pre_c
    tree = ec.build_from_leaves_dev(d_pre, n)
    t0 = time.perf_counter()
    assert tree.occupied == m
    t_index = time.perf_counter() - t0
    d_vals_c = synth.field_elements_torch(q, seed=555, device=dev)
    h_vals = d_vals_c[:3].cpu().numpy().view(np.uint64)
    d_vals = in_fmt(d_vals_c)
    del d_vals_c
    d_low = torch.empty(q, dtype=torch.int64, device=dev)
    d_match = torch.empty(q, dtype=torch.uint8, device=dev)
    l0 = ec.launches
    t_lookup = _ev_time(torch, stream, lambda: tree.low_leaf_lookup_dev(d_vals, q, d_low, d_match), steps, 2)
    launches = (ec.launches - l0) // (steps + 2)
    assert bool(d_match.all())
    t0 = time.perf_counter()
    for k in range(3):                                                           # the reference's literal scan (IMT:632-660), the oracle
        assert (int(d_low[k]), True) == O.low_leaf(h_pre, h_vals[k])
    cpu_lookup = 3 / (time.perf_counter() - t0)
    del h_pre
    bufs = tree.non_inclusion_buffers(q, depth, pinned=True)
    h_q = torch.empty((q, 4), dtype=torch.int64, pin_memory=True)
    h_q.copy_(d_vals)
    hq = h_q.numpy().view(np.uint64)
    tree.non_inclusion_paths(hq, out=bufs)
    t0 = time.perf_counter()
    o = tree.non_inclusion_paths(hq, out=bufs)
    t_e2e = time.perf_counter() - t0
    assert np.array_equal(o["low_idx"], d_low.cpu().numpy().astype(np.uint64))
    # the whole verify_non_inclusion witness, traced, one call, device resident: lookup + low leaf + limbs + (1 + depth) traced hashes per value
    qt = 1 << 14
    d_lowt = torch.empty(qt, dtype=torch.int64, device=dev)
    d_leavest = torch.empty((qt, 3, 4), dtype=torch.int64, device=dev)
    d_limbst = torch.empty((qt, 6, 4), dtype=torch.int64, device=dev)
    d_statest = torch.empty((qt, 1 + depth, 132, 3, 4), dtype=torch.int64, device=dev)
    t_nit = _ev_time(torch, stream, lambda: tree.trace_non_inclusion_dev(d_vals[:qt], qt, d_lowt, d_leavest, d_statest, d_limbs=d_limbst), 2, 1)
    assert bool((d_lowt == d_low[:qt]).all())
    d_rootc = torch.empty(4, dtype=torch.int64, device=dev)
    tree.root_dev(d_rootc)
    stream.synchronize()
    assert bool((d_statest[:, -1, -1, 1] == d_rootc).all()), "a traced non-inclusion path does not end in the root"
    del d_statest
    ins_out = tree.insert_buffers(b, depth, pinned=True)
    next_slot, ins = m, []
    for s_ in range(steps + 2):
        vals = synth.field_elements(b, seed=9000 + s_)
        if mont:
            vals = ec.convert(vals, to_montgomery=True)
        t0 = time.perf_counter()
        w = tree.insert_batch(vals, first_idx=next_slot, out=ins_out)
        ins.append(time.perf_counter() - t0)
        next_slot += b
    t_ins = sum(ins[2:]) / steps
    assert np.array_equal(w["new_roots"][-1], tree.root())
    # the whole insert_leaf witness trace of the last batch, one call, device resident
    dw = {k: torch.from_numpy(np.ascontiguousarray(w[k]).view(np.int64) if w[k].dtype == np.uint64 else w[k]).to(dev)
          for k in ("low_idx", "low_leaves", "low_siblings", "new_leaves", "new_siblings", "fold_nodes")}
    S = 3 + 4 * depth
    d_states = torch.empty((b, S, 132, 3, 4), dtype=torch.int64, device=dev)
    d_roots = torch.empty((b, 4, 4), dtype=torch.int64, device=dev)
    et = ec
    # one launch: the chain values of the four folds came with the insert batch (imt_insert_witness::fold_nodes)
    t_wt = _ev_time(torch, stream, lambda: et.trace_insert_witness_dev(dw, b, depth, next_slot - b, d_states, d_roots), 2, 1)
    assert np.array_equal(d_roots[:, 3].cpu().numpy().view(np.uint64), w["new_roots"])
    sample_one = d_states[:: max(1, b // 64)].clone()
    # without them: 1 + depth dependent launches of 4b traced hashes (hash latency)
    dw_loop = {k: v for k, v in dw.items() if k != "fold_nodes"}
    t_wt_loop = _ev_time(torch, stream, lambda: et.trace_insert_witness_dev(dw_loop, b, depth, next_slot - b, d_states, d_roots), 2, 1)
    assert torch.equal(sample_one, d_states[:: max(1, b // 64)]), "one-launch trace != level-loop trace"
    assert np.array_equal(d_roots[:, 3].cpu().numpy().view(np.uint64), w["new_roots"])
    hbm = _peaks().get("hbm_gbs", 6650.0)
    probes = max(1, m.bit_length())
    out = {
        "metric": "low_leaf_lookups_per_s", "value": q / (t_lookup * 1e-3), "unit": "lookups/s", "ms_per_step": t_lookup, "n_gpus": 1,
        "config": {"workload": f"2^20 uniform random low-leaf lookups on the depth-{depth} indexed tree ({m} occupied), non-inclusion witnesses, "
                               f"then {b}-insert batches with per-insert roots and paths", "queries": q, "occupied": m},
        "e2e": {"value": q / t_e2e, "unit": "lookups/s", "ms_per_step": t_e2e * 1e3, "h2d_bytes_per_step": q * 32,
                "d2h_bytes_per_step": q * (8 + 1 + 96 + depth * 33 + 1), "call": "imt_non_inclusion_paths into reused page-locked buffers"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "k_low_leaf_lookup", "achieved": q * probes * 32 / (t_lookup * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": q * probes * 32 / (t_lookup * 1e-3) / 1e9 / hbm, "traffic": None,
                     "note": f"algorithmic bytes = {probes} dependent 32-byte probes per lookup (latency-bound gather)"},
        "cpu_baseline": {"value": cpu_lookup, "unit": "lookups/s", "cores": 1, "kind": "port", "sample": "3 lookups by the reference's linear scan over the same preimages (oracle)"},
        "insert": {"metric": "inserts_per_s", "value": b / t_ins, "unit": "inserts/s", "ms_per_batch": t_ins * 1e3, "batch": b,
                   "hashes_per_s": 2 * b * (depth + 1) / t_ins, "call": "imt_insert_batch (host values in, witness bundle out into page-locked buffers)",
                   "witness_trace": {"call": "imt_insert_witness_trace_dev with fold_nodes (one launch of independent traced hashes)", "hashes": b * S,
                                     "fe_format": "montgomery" if et.fmt == 1 else "canonical", "ms": t_wt, "hashes_per_s": b * S / (t_wt * 1e-3), "bytes": b * S * 132 * 96,
                                     "level_loop": {"ms": t_wt_loop, "hashes_per_s": b * S / (t_wt_loop * 1e-3),
                                                    "note": "the same call without fold_nodes: 1 + depth dependent launches"}}},
        "non_inclusion_trace": {"call": "imt_non_inclusion_witness_trace_dev (lookup + low leaf + limbs + the 1 + depth traced hashes of verify_non_inclusion)",
                                "queries": qt, "hashes": qt * (1 + depth), "ms": t_nit, "hashes_per_s": qt * (1 + depth) / (t_nit * 1e-3),
                                "fe_format": "montgomery" if mont else "canonical (one from_mont per traced state element; Montgomery contexts skip it)"},
        "index_build_s": t_index, "fe_format": "montgomery" if mont else "canonical",
    }
    tree.close()
    return out


def run_paths_sharded(a):
    """--workload paths --gpus N (torchrun): the depth-D tree sharded by subtree, 2^16 paths + witness traces sharded by leaf owner,
    every rank draining its traces over its own PCIe link; aggregate rates, max time over ranks. One JSON line (rank 0)."""
    import torch
    import torch.distributed as dist
    import imt_b200
    from imt_b200 import synth
    from imt_b200.sharding import attach_communicator
    world, rank, local_rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {a.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    eng = imt_b200.Engine(local_rank, "montgomery")
    stream = torch.cuda.current_stream(dev)
    eng.set_stream(stream.cuda_stream)
    attach_communicator(eng)
    n = (1 << a.depth) // world
    d_pre = synth.field_elements_torch(3 * n, synth.DEFAULT_SEED, first=3 * n * rank, device=dev).view(n, 3, 4)
    tree = eng.sharded_build_from_leaves_dev(d_pre, n)
    imad_rate, _ = eng.calibrate_imad(150.0)
    out = secondary_paths(torch, dist, eng, tree, d_pre, a.depth, n, world, rank, dev, stream, imad_rate, a.steps)
    out.update({"steps": a.steps, "warmup": a.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32x8-montgomery",
                "data": "synthetic"})
    if rank == 0:
        out["cpu_baseline"] = cpu_trace_rate()
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


# ------------------------------------------------------------------------------------------- GPU arm
def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    if a.workload == "paths" and a.gpus == 1:
        return run_paths(a)
    if a.workload == "paths":
        return run_paths_sharded(a)
    if a.workload == "lookups":
        return run_lookups(a)

    import torch
    import torch.distributed as dist
    import imt_b200
    from imt_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {a.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)  # plumbing only: barrier, max-over-ranks of the timing, the NCCL id

    depth = a.depth
    n_total = 1 << depth
    n = n_total // world
    eng = imt_b200.Engine(local_rank, "montgomery")   # Montgomery = halo2curves' in-memory form: zero-copy from Rust
    stream = torch.cuda.current_stream(dev)
    eng.set_stream(stream.cuda_stream)
    nccl_version = 0
    if world > 1:
        # the data-path collective lives INSIDE the C library (csrc/imt_comm.cu): the engine gets its own NCCL communicator
        # (imt_comm_create); torch.distributed only carries the 128-byte id from rank 0 to the other ranks
        from imt_b200.sharding import attach_communicator
        attach_communicator(eng)
        nccl_version = eng.comm_info()[2]

    # ---- synthetic leaves, generated on the device (setup, untimed): rank r owns leaves [r n, (r+1) n)
    d_pre = synth.field_elements_torch(3 * n, synth.DEFAULT_SEED, first=3 * n * rank, device=dev).view(n, 3, 4)
    h_pre = torch.empty((n, 3, 4), dtype=torch.int64, pin_memory=True)
    h_pre.copy_(d_pre)
    torch.cuda.synchronize(dev)
    # N > 1: imt_sharded_build_from_leaves_dev = local subtree build -> ncclAllGather of the N roots (sent straight out of the
    # tree's level buffer, received straight into the cap) -> log2(N) cap levels, all queued on one stream by one C call
    tree = eng.sharded_build_from_leaves_dev(d_pre, n) if world > 1 else eng.build_from_leaves_dev(d_pre, n)
    send = torch.zeros(4, dtype=torch.int64, device=dev)
    h_root = torch.zeros(4, dtype=torch.int64, pin_memory=True)

    def step_resident():
        if world > 1:
            tree.sharded_rebuild_from_leaves_dev(d_pre)                  # imt_sharded_rebuild_from_leaves_dev
        else:
            tree.rebuild_from_leaves_dev(d_pre)

    def step_e2e():
        if world > 1:
            tree.sharded_rebuild_from_leaves_ptr(h_pre.data_ptr())       # pinned host -> device inside the call, exchange included
        else:
            tree.rebuild_from_leaves_ptr(h_pre.data_ptr())
        tree.root_dev(send)
        h_root.copy_(send, non_blocking=True)                            # the step's result back on the host
        stream.synchronize()

    hashes_per_step = world * (2 * n - 1) + (world - 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launches
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), eng.launches - l0

    # ---- value: inputs resident in HBM. Per-kernel device times come from the library's own event pairs.
    eng.enable_timing(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(a.warmup, 3)):
        step_resident()
    eng.reset_timing()
    sampler.mark_begin()
    ms_total, launches = timed(step_resident, a.steps, 0)
    sampler.mark_end()
    clocks = sampler.stop()
    k3_ms, k3_launches, k3_hashes = eng.kernel_time(3)
    k2_ms, k2_launches, k2_hashes = eng.kernel_time(2)
    eng.enable_timing(False)
    value = hashes_per_step * a.steps / (ms_total * 1e-3)

    # ---- e2e: the reference-facing call with HOST buffers (H2D of the leaves + D2H of the root inside the timed region)
    ms_e2e, _ = timed(step_e2e, a.steps, a.warmup)
    e2e = hashes_per_step * a.steps / (ms_e2e * 1e-3)
    # the same call the reference's `IndexedMerkleTree::new` maps to: allocate, build, read the root, destroy — every step
    def step_e2e_alloc():
        t2 = eng.build_from_leaves_ptr(h_pre.data_ptr(), n)
        if world == 1:
            t2.root_dev(send)
            h_root.copy_(send, non_blocking=True)
            stream.synchronize()
        t2.close()
    ms_e2e_alloc = None
    if world == 1:
        step_e2e_alloc()                                                  # first call grows the pool
        t0 = time.perf_counter()
        for _ in range(2):
            step_e2e_alloc()
        ms_e2e_alloc = (time.perf_counter() - t0) / 2 * 1e3
    # the same call from PAGEABLE host memory (a Rust Vec<F> is plain malloc memory): the library copies straight from the
    # caller's pointer chunk by chunk, the driver stages each chunk while the previous chunk's kernel runs
    h_page = h_pre.numpy().copy()
    def step_e2e_pageable():
        if world > 1:
            tree.sharded_rebuild_from_leaves_ptr(h_page.ctypes.data)
        else:
            tree.rebuild_from_leaves_ptr(h_page.ctypes.data)
        tree.root_dev(send)
        h_root.copy_(send, non_blocking=True)
        stream.synchronize()
    ms_e2e_pageable, _ = timed(step_e2e_pageable, max(2, a.steps // 2), 1)
    ms_e2e_pageable /= max(2, a.steps // 2)
    del h_page
    # what a wrapper that creates a context per `IndexedMerkleTree::new` would add (INTEGRATION.md keeps ONE per process instead)
    t0 = time.perf_counter()
    tmp_eng = imt_b200.Engine(local_rank, "montgomery")
    tmp_eng.close()
    ctx_create_ms = (time.perf_counter() - t0) * 1e3
    root_hex = "".join(f"{int(x) & 0xFFFFFFFFFFFFFFFF:016x}" for x in reversed(h_root.tolist()))
    # the root of THIS input is pinned: tests/golden/golden.json "bench_roots" holds the CPU oracle's root of the same
    # Montgomery-interpreted synthetic stream (tests/golden/make_golden.py --bench-roots); every rank count must reproduce it
    try:
        gold_root = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json"))).get("bench_roots", {}).get(str(depth))
    except Exception:
        gold_root = None
    if gold_root is not None and root_hex != gold_root:
        raise SystemExit(f"bench.py: root {root_hex} differs from the oracle's golden root {gold_root} (depth {depth}, {world} GPUs)")
    root_check = "equals the CPU oracle's golden root (tests/golden/golden.json bench_roots)" if gold_root else "no golden root for this depth"

    # ---- roofline of the dominant kernel (leaf hashing: half of all hashes in one launch)
    imad_rate, imad_mhz = eng.calibrate_imad(150.0)
    k3_per_launch_ms = k3_ms / max(k3_launches, 1)
    achieved = (k3_hashes / max(k3_launches, 1)) * MACS_PER_HASH / (k3_per_launch_ms * 1e-3) if k3_ms else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    leaf_bytes = 128.0  # 96 B preimage read + 32 B hash written per leaf hash
    roofline = {
        "bound": "imad", "kernel": "k_hash<3> (leaf hashing)", "achieved": achieved / 1e9, "peak": imad_rate / 1e9, "unit": "GMAC/s",
        "frac": achieved / imad_rate if imad_rate else None,
        "executed_frac": achieved / imad_rate * EXECUTED_MACS_PER_HASH / MACS_PER_HASH if imad_rate else None,
        "traffic": LEAF_DRAM_BYTES_PER_HASH * (k3_hashes / max(k3_launches, 1)),
        "traffic_note": "bytes per launch = ncu dram__bytes_read.sum + dram__bytes_write.sum of this kernel per leaf (129.3 B, one --set full "
                        "capture, profiles/) x leaves per launch; algorithmic 128 B per leaf",
        "frac_note": "frac uses ALGORITHMIC MACs (163200 per hash) and exceeds 1 because squarings and one-reduction dot products execute "
                     "123792; executed_frac is the multiply-pipe utilisation (ncu sm__pipe_fmaheavy_cycles_active agrees)",
        "peak_source": f"calibrated in this run: IMAD.WIDE.U32.X carry chains saturating all SMs (imt_calibrate_imad); = 32 wide MACs/clk/SM at "
                       f"{imad_mhz:.0f} MHz. MEASURED_PEAKS.json has no integer peak",
        "algorithmic_macs_per_hash": MACS_PER_HASH, "kernel_ms_per_launch": k3_per_launch_ms,
        "kernel_share_of_step": (k3_ms / a.steps) / (ms_total / a.steps) if ms_total else None,
        "node_levels_ms_per_step": (ms_total - k3_ms) / a.steps, "node_kernel_launches_per_step": k2_launches / a.steps,
        "node_levels_note": "step time minus the leaf kernel: from depth 16 on the levels of the two half-trees run on two streams, so "
                            "per-launch event pairs of k_hash<2> overlap and are not summed",
        "hbm": {"achieved_gbs": (k3_hashes / max(k3_launches, 1)) * leaf_bytes / (k3_per_launch_ms * 1e-3) / 1e9 if k3_ms else None,
                "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"},
    }

    out = {
        "metric": METRIC, "value": value, "unit": "hashes/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32x8-montgomery", "data": "synthetic",
        "config": {"workload": workload_name(depth),
                   "depth": depth, "leaves": n_total, "leaves_per_gpu": n, "hashes_per_step": hashes_per_step,
                   "sharding": f"subtree x{world} + all-gather of {world} roots" if world > 1 else "single GPU",
                   "l2_policy": f"inputs larger than L2 ({n * 96 / 2**20:.0f} MiB of leaves per GPU per step)", "seed": synth.DEFAULT_SEED,
                   "fe_format": "montgomery"},
        "e2e": {"value": e2e, "unit": "hashes/s", "ms_per_step": ms_e2e / a.steps, "h2d_bytes_per_step": n_total * 96,
                "d2h_bytes_per_step": 32 * world, "call": ("imt_sharded_rebuild_from_leaves" if world > 1 else "imt_tree_rebuild_from_leaves") + " (host leaves -> existing tree) + root read back",
                "pageable": {"value": hashes_per_step / (ms_e2e_pageable * 1e-3), "ms_per_step": ms_e2e_pageable,
                             "note": "the same call with the leaves in plain malloc (pageable) memory, as a Rust Vec<F> is"},
                "ctx_create_ms": ctx_create_ms,
                "ms_per_step_with_alloc": ms_e2e_alloc,
                "with_alloc_note": "wall clock of imt_tree_build_from_leaves + root + imt_tree_destroy per step (2.5 GiB of tree buffers recycled through the stream-ordered pool), N=1 only"},
        "gpu_launches": launches * world, "clocks": clocks, "roofline": roofline, "root": root_hex, "root_check": root_check,
        "collective": (f"ncclAllGather of {world} x 32 B issued inside libimt_b200.so (imt_sharded_rebuild_from_leaves*, NCCL {nccl_version})"
                       if world > 1 else None),
    }
    # ---- BASELINE configs 4 and 5 as secondary legs of the same run (so that the driver's line carries them)
    if not a.no_secondary:
        sec = {}
        try:
            sec["paths"] = secondary_paths(torch, dist, eng, tree, d_pre, depth, n, world, rank, dev, stream, imad_rate, max(2, a.steps // 2))
        except Exception as e:  # a secondary leg must never cost the headline line
            sec["paths"] = {"error": repr(e)}
        tree.close()
        del d_pre, h_pre
        torch.cuda.empty_cache()
        eng.trim()
        if world == 1:
            try:
                sec["lookups"] = secondary_lookups(torch, eng, depth, dev, stream, max(2, a.steps // 2))
            except Exception as e:
                sec["lookups"] = {"error": repr(e)}
        else:
            sec["lookups"] = {"skipped": "one-GPU leg (python bench.py --workload lookups); the sharded lookups / inserts are covered by tests and tools/multi_gpu_check.py"}
        out["secondary"] = sec
    if rank == 0 and not a.no_cpu_baseline:
        if "paths" in out.get("secondary", {}) and "error" not in out["secondary"]["paths"]:
            out["secondary"]["paths"]["cpu_baseline"] = cpu_trace_rate()
        v, th, sample, _, _ = cpu_build_rate(a.cpu_seconds)
        out["cpu_baseline"] = {"value": v, "unit": "hashes/s", "cores": th, "kind": "port", "sample": sample,
                               "single_thread": dict(zip(("value", "sample"), cpu_single_thread_rate()))}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""CPU check of the PRODUCT's field / Poseidon source (csrc/fr.cuh, poseidon.cuh, poseidon_params.cpp) compiled in
host mode, where the PTX carry primitives are emulated with a thread-local flag. It proves the algorithm the
kernels run (even/odd wide accumulators, separate reduction, dedicated squaring, lazy [0,2p) ranges, the
carry-flag invariant) against Python big integers — without a GPU. The emulation is test-only."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

import poseidon_ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "indexed-merkle-tree-halo2_b200", "csrc")
P = R.P
RM = 1 << 256
RINV = pow(RM, -1, P)
u32p = ctypes.POINTER(ctypes.c_uint32)


@pytest.fixture(scope="module")
def shim():
    out = os.path.join(ROOT, "tests", "_build", "libhost_shim.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    srcs = [os.path.join(ROOT, "tests", "host_shim.cpp"), os.path.join(CSRC, "poseidon_params.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("fr.cuh", "poseidon.cuh", "poseidon_spec.cuh", "poseidon_params.h", "poseidon_lh_math.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", CSRC, "-x", "c++", *srcs, "-o", out], check=True)
    return ctypes.CDLL(out)


def w(x, n=8):
    return np.array([(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)], dtype=np.uint32)


def iv(a):
    return sum(int(v) << (32 * i) for i, v in enumerate(a))


def p_(a):
    return a.ctypes.data_as(u32p)


EDGE = [0, 1, P - 1, P, P + 1, 2 * P - 1, 2 * P - 2, 0xFFFFFFFF, (1 << 255) % (2 * P), int("ffffffff" * 8, 16) % (2 * P)]


def rnd(rng):
    r = rng.random()
    if r < 0.2:
        return rng.choice(EDGE)
    if r < 0.35:
        v = 0
        for i in range(8):
            v |= rng.choice([0, 0xFFFFFFFF, rng.getrandbits(32)]) << (32 * i)
        return v % (2 * P)
    return rng.randrange(2 * P)


def test_field_ops_semi_reduced(shim):
    rng = random.Random(7)
    r = np.zeros(8, np.uint32)
    o = np.zeros(16, np.uint32)
    for _ in range(6000):
        a, b = rnd(rng), rnd(rng)
        A, B = w(a), w(b)
        shim.shim_mont_mul(p_(r), p_(A), p_(B))
        v = iv(r)
        assert v % P == a * b * RINV % P and v < 2 * P
        shim.shim_mont_sqr(p_(r), p_(A))
        v = iv(r)
        assert v % P == a * a * RINV % P and v < 2 * P
        shim.shim_add_semi(p_(r), p_(A), p_(B))
        v = iv(r)
        assert v % P == (a + b) % P and v < 2 * P
        u = [rnd(rng) for _ in range(3)]
        m = [rng.choice([P - 1, rng.randrange(P)]) for _ in range(3)]
        U = np.concatenate([w(x) for x in u])
        M = np.concatenate([w(x) for x in m])
        shim.shim_dot3(p_(r), p_(U), p_(M))
        v = iv(r)
        assert v % P == sum(x * y for x, y in zip(u, m)) * RINV % P and v < 2 * P
        s = rnd(rng)
        shim.shim_mul_add(p_(r), p_(w(u[0])), p_(w(m[0])), p_(w(s)))
        v = iv(r)
        assert v % P == (u[0] * m[0] * RINV + s) % P and v < 2 * P


def test_wide_products_full_range(shim):
    rng = random.Random(11)
    o = np.zeros(16, np.uint32)
    full = (1 << 256) - 1
    cases = [(full, full), (full, 1), (0, full), (1 << 255, 1 << 255)]
    for _ in range(4000):
        cases.append((rng.getrandbits(256), rng.choice([rng.getrandbits(256), full, rng.getrandbits(32) << 224])))
    for a, b in cases:
        shim.shim_mul_wide_merged(p_(o), p_(w(a)), p_(w(b)))
        assert iv(o) == a * b
        shim.shim_sqr_wide(p_(o), p_(w(b)))
        assert iv(o) == b * b


def test_montgomery_conversions(shim):
    rng = random.Random(3)
    r = np.zeros(8, np.uint32)
    for x in [0, 1, P - 1] + [rng.randrange(P) for _ in range(500)]:
        shim.shim_to_mont(p_(r), p_(w(x)))
        m = iv(r)
        assert m % P == x * RM % P
        shim.shim_from_mont(p_(r), p_(w(m)))
        assert iv(r) == x


def test_params_match_reference_derivation(shim):
    sz = shim.shim_params_size()
    buf = np.zeros(sz // 4, np.uint32)
    shim.shim_params(p_(buf))
    sp = R.spec()
    exp = []
    for r_ in range(3):
        exp += sp.start[r_ + 1]
    exp += sp.start[4]
    for r_ in range(3):
        exp += sp.end[r_]
    exp += [0, 0, 0]
    exp += sp.start[0]
    for row in sp.mds:
        exp += row
    for row in sp.pre_sparse:
        exp += row
    for k in range(57):
        exp += [sp.partial[k]] + list(sp.sparse[k][0]) + list(sp.sparse[k][1])
    exp += [1 << 64, 1]
    assert sz == 32 * len(exp)
    got = [iv(buf[8 * i:8 * i + 8]) for i in range(len(exp))]
    assert all(g < P for g in got)  # canonical Montgomery
    assert [g * RINV % P for g in got] == exp


def test_hash_and_trace(shim):
    rng = random.Random(5)
    cases = [(3, [0, 0, 0]), (2, [0, 0]), (2, [1, 2]), (3, [1, 2, 3]), (2, [P - 1, P - 2]), (3, [P - 1] * 3)]
    cases += [(3, [rng.randrange(P) for _ in range(3)]) for _ in range(3)] + [(2, [rng.randrange(P) for _ in range(2)]) for _ in range(3)]
    for ar, x in cases:
        I = np.concatenate([w(v) for v in x])
        out = np.zeros(8, np.uint32)
        st = np.zeros(132 * 3 * 8, np.uint32)
        shim.shim_hash(p_(out), p_(I), ar, p_(st))
        d, tr = R.hash_trace(x)
        assert iv(out) == d
        assert [iv(st[8 * i:8 * i + 8]) for i in range(396)] == [v for s_ in tr for v in s_]
        shim.shim_hash(p_(out), p_(I), ar, None)
        assert iv(out) == d
        sb = np.zeros(162 * 3 * 8, np.uint32)                    # extended trace: (x^2, x^4, x^5 + c) of all 162 S-boxes
        shim.shim_hash_ext(p_(out), p_(I), ar, p_(st), p_(sb))
        cells = []
        R.hash_trace_n(x, None, cells)
        assert iv(out) == d and [iv(sb[8 * i:8 * i + 8]) for i in range(486)] == [v for c in cells for v in c]
    I = np.concatenate([w(0)] * 3)
    shim.shim_hash(p_(out), p_(I), 3, None)
    assert iv(out) == R.KAT_H3_ZERO  # /root/reference/src/indexed_merkle_tree.rs:247-251


@pytest.mark.parametrize("t,r_f,r_p", [(2, 8, 56), (3, 8, 57), (4, 8, 56), (5, 8, 60), (3, 6, 10), (4, 2, 0)])
def test_any_width_sponge_source_on_the_host(shim, t, r_f, r_p):
    """csrc/poseidon_spec.cuh (the any-width permutation + sponge the GPU kernels run) compiled for the host with the emulated
    carry flag: digests and per-round traces for every input length 0 .. 2t + 1 against the Python oracle"""
    sp = R.Spec(r_f, r_p, t)
    rng = random.Random(t * 100 + r_p)
    per_perm = (1 + r_f + r_p) * t
    for arity in range(0, 2 * t + 2):
        x = [rng.randrange(P) for _ in range(arity)]
        I = np.concatenate([w(v) for v in x]) if arity else np.zeros(8, np.uint32)
        perms = arity // (t - 1) + 1
        out = np.zeros(8, np.uint32)
        st = np.zeros(perms * per_perm * 8, np.uint32)
        assert shim.shim_spec_hash(t, r_f, r_p, p_(I), ctypes.c_size_t(arity), p_(out), p_(st)) == 0
        d, tr = R.hash_trace_n(x, sp)
        assert iv(out) == d, (t, arity)
        assert [iv(st[8 * i:8 * i + 8]) for i in range(perms * per_perm)] == [v for s_ in tr for v in s_]
        assert shim.shim_spec_hash(t, r_f, r_p, p_(I), ctypes.c_size_t(arity), p_(out), None) == 0 and iv(out) == d
    assert shim.shim_spec_hash(6, 8, 57, p_(out), ctypes.c_size_t(0), p_(out), None) == 1


def test_lead_helper_recurrence_equals_the_partial_rounds(shim):
    """csrc/poseidon_lh_math.cuh: the affine reformulation of a partial round the lead / helper latency kernel (k_hash_lh) runs
    across warps — x' = x^4 (row_0 x) + K, s_i' = x^4 (col_i x) + (col_i c + s_i), K' = x^4 (rho x) + (kappa + row_1' s_1 + row_2' s_2)
    with the derived per-round tables rho, kappa, col_i c — produces the same 57-round state as the product's partial_round, and the
    tables are what the closed forms say (big-int arithmetic on the oracle's parameters)."""
    rng = random.Random(57)
    states = [[0, 0, 0], [P - 1, P - 1, P - 1], [1, 0, P - 1]] + [[rng.randrange(P) for _ in range(3)] for _ in range(12)]
    for st in states:
        a = np.concatenate([w(v) for v in st])
        b = a.copy()
        shim.shim_partial_rounds(p_(a), 0)
        shim.shim_partial_rounds(p_(b), 1)
        assert [iv(a[8 * i:8 * i + 8]) for i in range(3)] == [iv(b[8 * i:8 * i + 8]) for i in range(3)]
        assert all(iv(a[8 * i:8 * i + 8]) < P for i in range(3))
    # and against the independent Python permutation: 57 partial rounds of the optimized schedule on big ints
    sp = R.spec()
    st = [rng.randrange(P) for _ in range(3)]
    want = list(st)
    for k in range(sp.r_p):  # poseidon_ref.permute_trace, partial rounds
        row, col_hat = sp.sparse[k]
        x2 = want[0] * want[0] % P
        s0 = (x2 * x2 % P * want[0] + sp.partial[k]) % P
        v = [s0] + want[1:]
        want = [sum(r * x for r, x in zip(row, v)) % P] + [(col_hat[i - 1] * s0 + v[i]) % P for i in range(1, 3)]
    a = np.concatenate([w(v) for v in st])
    shim.shim_partial_rounds(p_(a), 1)
    assert [iv(a[8 * i:8 * i + 8]) for i in range(3)] == want

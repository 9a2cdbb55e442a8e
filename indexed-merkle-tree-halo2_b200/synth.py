"""Deterministic synthetic inputs: counter-based splitmix64 -> field elements -> indexed leaves.

FE e of stream `seed` = words splitmix(seed, 4e .. 4e+3), top word masked to 62 bits, minus p if >= p.
Vectorised numpy; the oracle has an independent C version (imto_synth_fe) the tests compare against.
"""
import numpy as np

DEFAULT_SEED = 0x494D54  # "IMT"
_P_WORDS = np.array([0x43E1F593F0000001, 0x2833E84879B97091, 0xB85045B68181585D, 0x30644E72E131A029], dtype=np.uint64)


def _splitmix(seed, ctr):
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (ctr + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _ge_p(w):
    ge = np.ones(w.shape[0], dtype=bool)
    decided = np.zeros(w.shape[0], dtype=bool)
    for k in (3, 2, 1, 0):
        gt = (w[:, k] > _P_WORDS[k]) & ~decided
        lt = (w[:, k] < _P_WORDS[k]) & ~decided
        ge[lt] = False
        decided |= gt | lt
    return ge


def _sub_p(w, mask):
    borrow = np.zeros(w.shape[0], dtype=np.uint64)
    with np.errstate(over="ignore"):
        for k in range(4):
            a = w[:, k]
            t = a - _P_WORDS[k]
            b1 = (a < _P_WORDS[k]).astype(np.uint64)
            t2 = t - borrow
            b2 = (t < borrow).astype(np.uint64)
            w[:, k] = np.where(mask, t2, a)
            borrow = b1 | b2


def field_elements(n, seed=DEFAULT_SEED, first=0):
    """(n, 4) uint64 canonical field elements"""
    ctr = (np.arange(n, dtype=np.uint64) + np.uint64(first)) * np.uint64(4)
    w = np.stack([_splitmix(seed, ctr + np.uint64(k)) for k in range(4)], axis=1)
    w[:, 3] &= np.uint64(0x3FFFFFFFFFFFFFFF)
    _sub_p(w, _ge_p(w))
    return np.ascontiguousarray(w)


def random_preimages(n, seed=DEFAULT_SEED):
    """n x 3 independent uniform FE — hashing cost is data independent, so this is the throughput workload"""
    return field_elements(3 * n, seed).reshape(n, 3, 4)


def indexed_preimages(n, occupied=None, seed=DEFAULT_SEED):
    """A well-formed indexed tree: slot 0 is the head {0,..}, slots 1..occupied-1 hold distinct random values in
    insertion (random) order, each pointing at its successor in sorted order; the largest points at (0, 0);
    the remaining slots are empty {0,0,0}.  Returns (n, 3, 4) uint64 canonical."""
    m = n if occupied is None else occupied
    assert 1 <= m <= n
    pre = np.zeros((n, 3, 4), dtype=np.uint64)
    if m == 1:
        return pre
    vals = field_elements(m - 1, seed)
    order = np.lexsort((vals[:, 0], vals[:, 1], vals[:, 2], vals[:, 3]))  # ascending by value
    sv = vals[order]
    if m > 2 and (np.any(np.all(sv[1:] == sv[:-1], axis=1)) or not sv[0].any()):
        raise ValueError("synthetic values collided; pick another seed")
    pre[1:m, 0] = vals
    slots = order.astype(np.uint64) + np.uint64(1)  # slot of the k-th smallest value
    # head -> smallest
    pre[0, 1] = sv[0]
    pre[0, 2, 0] = slots[0]
    # k-th smallest -> (k+1)-th smallest; largest -> (0, 0)
    pre[slots[:-1], 1] = sv[1:]
    pre[slots[:-1], 2, 0] = slots[1:]
    return pre


def _settle(torch, device):
    """The engine launches on its OWN non-blocking stream: a tensor handed to it must be complete, and — subtler — every torch kernel
    queued before must have finished with its temporaries, because a later torch.empty() may reuse their memory (safe on torch's
    stream, not for a writer on another stream). The synthetic-input helpers therefore return settled tensors."""
    if str(device) != "cpu":
        torch.cuda.synchronize(device)


def field_elements_torch(n, seed=DEFAULT_SEED, first=0, device="cuda"):
    """Same stream as field_elements(), generated with torch integer ops on `device`: (n, 4) int64 (bit pattern =
    uint64). Setup helper for bench.py so that 2^24 x 3 elements need no host round trip."""
    import torch

    def u(x):  # python int -> int64 bit pattern
        x &= 0xFFFFFFFFFFFFFFFF
        return x - (1 << 64) if x >= (1 << 63) else x

    def lsr(x, k):  # logical shift right on int64
        return (x >> k) & u((1 << (64 - k)) - 1)

    MIN = u(1 << 63)
    chunks = []
    step = 1 << 22
    for lo in range(0, n, step):
        cnt = min(step, n - lo)
        e = torch.arange(first + lo, first + lo + cnt, dtype=torch.int64, device=device)
        ws = []
        for k in range(4):
            ctr = e * 4 + k
            z = (ctr + 1) * u(0x9E3779B97F4A7C15) + u(seed)
            z = (z ^ lsr(z, 30)) * u(0xBF58476D1CE4E5B9)
            z = (z ^ lsr(z, 27)) * u(0x94D049BB133111EB)
            ws.append(z ^ lsr(z, 31))
        ws[3] = ws[3] & u(0x3FFFFFFFFFFFFFFF)
        pw = [u(int(v)) for v in _P_WORDS]
        # unsigned comparison via sign flip
        ge = torch.ones(cnt, dtype=torch.bool, device=device)
        decided = torch.zeros(cnt, dtype=torch.bool, device=device)
        for k in (3, 2, 1, 0):
            a, b = ws[k] ^ MIN, pw[k] ^ MIN
            gt, lt = (a > b) & ~decided, (a < b) & ~decided
            ge = ge & ~lt
            decided = decided | gt | lt
        borrow = torch.zeros(cnt, dtype=torch.int64, device=device)
        out = []
        for k in range(4):
            a = ws[k]
            t = a - pw[k]
            b1 = ((a ^ MIN) < (pw[k] ^ MIN)).to(torch.int64)
            t2 = t - borrow
            b2 = ((t ^ MIN) < (borrow ^ MIN)).to(torch.int64)
            out.append(torch.where(ge, t2, a))
            borrow = b1 | b2
        chunks.append(torch.stack(out, dim=1))
    out = torch.cat(chunks, dim=0).contiguous() if chunks else torch.zeros((0, 4), dtype=torch.int64, device=device)
    _settle(torch, device)
    return out


def indexed_preimages_torch(n, occupied=None, seed=DEFAULT_SEED, device="cuda"):
    """indexed_preimages() built on `device` with torch (same values, same layout): (n, 3, 4) int64 (bit pattern = uint64,
    canonical). Setup helper for the depth-24 query benchmarks — the host version needs a 16M-row numpy lexsort."""
    import torch
    m = n if occupied is None else occupied
    assert 1 <= m <= n
    pre = torch.zeros((n, 3, 4), dtype=torch.int64, device=device)
    if m == 1:
        _settle(torch, device)
        return pre
    vals = field_elements_torch(m - 1, seed, device=device)
    MIN = -(1 << 63)
    order = torch.arange(m - 1, device=device)
    for k in range(4):  # LSD: four stable sorts, least significant word first, unsigned order via the sign flip
        key = vals[order, k] ^ MIN
        order = order[torch.sort(key, stable=True).indices]
    sv = vals[order]
    if m > 2 and bool((sv[1:] == sv[:-1]).all(dim=1).any()):
        raise ValueError("synthetic values collided; pick another seed")
    pre[1:m, 0] = vals
    slots = order + 1
    pre[0, 1] = sv[0]
    pre[0, 2, 0] = slots[0]
    pre[slots[:-1], 1] = sv[1:]
    pre[slots[:-1], 2, 0] = slots[1:]
    _settle(torch, device)
    return pre

"""TEST INFRASTRUCTURE — ctypes binding of oracle/_build/libimt_oracle.so (oracle/imt_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Field elements cross this boundary as CANONICAL little-endian 4 x uint64 rows of a numpy array.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libimt_oracle.so")
P = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001

_u64p = ctypes.POINTER(ctypes.c_uint64)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    src = os.path.join(_HERE, "imt_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.imto_init.restype = ctypes.c_int
        L.imto_constants.restype = ctypes.c_size_t
        L.imto_constants.argtypes = [ctypes.c_int, _u64p]
        L.imto_permute.argtypes = [_u64p, _u64p, ctypes.c_int]
        L.imto_hash2.argtypes = [_u64p, ctypes.c_size_t, _u64p, ctypes.c_int]
        L.imto_hash3.argtypes = [_u64p, ctypes.c_size_t, _u64p, ctypes.c_int]
        L.imto_hash_trace.argtypes = [_u64p, ctypes.c_int, _u64p, _u64p]
        L.imto_tree_build.argtypes = [_u64p, ctypes.c_size_t, _u64p, ctypes.c_int]
        L.imto_tree_build.restype = ctypes.c_int
        L.imto_get_proof.argtypes = [_u64p, ctypes.c_size_t, ctypes.c_size_t, _u64p, _u8p]
        L.imto_get_proof.restype = ctypes.c_int
        L.imto_verify_proof.argtypes = [_u64p, ctypes.c_size_t, _u64p, _u64p, ctypes.c_size_t]
        L.imto_verify_proof.restype = ctypes.c_int
        L.imto_update_idx_leaf.argtypes = [_u64p, ctypes.c_size_t, _u64p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_int)]
        L.imto_update_idx_leaf.restype = ctypes.c_size_t
        L.imto_low_leaf.argtypes = [_u64p, ctypes.c_size_t, _u64p, ctypes.POINTER(ctypes.c_int)]
        L.imto_low_leaf.restype = ctypes.c_size_t
        rnd = [_u64p, _u64p, ctypes.c_size_t, _u64p, ctypes.c_uint64]
        outs = [_u64p, _u64p, _u64p, _u64p, _u8p, _u64p, _u64p, _u64p, _u8p, _u8p]
        L.imto_insert_round_rebuild.argtypes = rnd + outs + [ctypes.c_int]
        L.imto_insert_round_rebuild.restype = ctypes.c_int
        L.imto_insert_round_incremental.argtypes = rnd + [ctypes.c_size_t, ctypes.c_int] + outs
        L.imto_insert_round_incremental.restype = ctypes.c_int
        L.imto_synth_fe.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_size_t, _u64p]
        L.imto_max_threads.restype = ctypes.c_int
        L.imto_build_from_preimages.argtypes = [_u64p, ctypes.c_size_t, _u64p, ctypes.c_int]
        L.imto_build_from_preimages.restype = ctypes.c_int
        L.imto_convert.argtypes = [_u64p, ctypes.c_size_t, _u64p, ctypes.c_int, ctypes.c_int]
        L.imto_init()
        _lib = L
    return _lib


# ----------------------------------------------------------------------------- conversions
def fe(x):
    """int -> (4,) uint64 canonical LE"""
    x %= P
    return np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def fes(xs):
    return np.stack([fe(int(x)) for x in xs]) if len(xs) else np.zeros((0, 4), np.uint64)


def to_int(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1)
    return sum(int(a[i]) << (64 * i) for i in range(4))


def to_ints(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    return [to_int(r) for r in a]


def _p(a, t=_u64p):
    return a.ctypes.data_as(t)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


# ----------------------------------------------------------------------------- API
def max_threads():
    return lib().imto_max_threads()


def constants(kind):
    n = lib().imto_constants(kind, None)
    out = np.zeros((n, 4), np.uint64)
    lib().imto_constants(kind, _p(out))
    return out


def permute(state, naive=False):
    s = _c(state).reshape(3, 4)
    out = np.zeros_like(s)
    lib().imto_permute(_p(s), _p(out), 1 if naive else 0)
    return out


def hash2(pairs, threads=1):
    a = _c(pairs).reshape(-1, 2, 4)
    out = np.zeros((a.shape[0], 4), np.uint64)
    lib().imto_hash2(_p(a), a.shape[0], _p(out), threads)
    return out


def hash3(triples, threads=1):
    a = _c(triples).reshape(-1, 3, 4)
    out = np.zeros((a.shape[0], 4), np.uint64)
    lib().imto_hash3(_p(a), a.shape[0], _p(out), threads)
    return out


def hash_trace(inputs):
    a = _c(inputs).reshape(-1, 4)
    states = np.zeros((132, 3, 4), np.uint64)
    digest = np.zeros(4, np.uint64)
    lib().imto_hash_trace(_p(a), a.shape[0], _p(states), _p(digest))
    return digest, states


ERRORS = {1: "Cannot create Merkle Tree with no leaves", 2: "Leaves must be even", 3: "not a power of two", 4: "index out of bounds"}


def tree_build(leaves, threads=1):
    """returns the concatenated levels ((2n-1), 4) — raises ValueError with the reference's message on error"""
    a = _c(leaves).reshape(-1, 4)
    n = a.shape[0]
    out = np.zeros((max(2 * n - 1, 1), 4), np.uint64)
    rc = lib().imto_tree_build(_p(a), n, _p(out), threads)
    if rc:
        raise ValueError(ERRORS[rc])
    return out


def levels(tree, n):
    out, off = [], 0
    ln = n
    while ln >= 1:
        out.append(tree[off:off + ln])
        off += ln
        ln >>= 1
    return out


def get_proof(tree, n, index):
    d = n.bit_length() - 1
    sib = np.zeros((d, 4), np.uint64)
    hel = np.zeros(d, np.uint8)
    rc = lib().imto_get_proof(_p(_c(tree)), n, index, _p(sib), _p(hel, _u8p))
    if rc:
        raise IndexError(ERRORS[rc])
    return sib, hel


def verify_proof(leaf, index, root, proof):
    pr = _c(proof).reshape(-1, 4)
    return bool(lib().imto_verify_proof(_p(_c(leaf)), index, _p(_c(root)), _p(pr), pr.shape[0]))


def update_idx_leaf(pre, new_val, new_idx):
    """(IMT:632-660) returns (new preimages, low_idx, matched)"""
    a = _c(pre).reshape(-1, 3, 4).copy()
    m = ctypes.c_int(0)
    low = lib().imto_update_idx_leaf(_p(a), a.shape[0], _p(_c(new_val)), new_idx, ctypes.byref(m))
    return a, int(low), bool(m.value)


def low_leaf(pre, new_val):
    a = _c(pre).reshape(-1, 3, 4)
    m = ctypes.c_int(0)
    low = lib().imto_low_leaf(_p(a), a.shape[0], _p(_c(new_val)), ctypes.byref(m))
    return int(low), bool(m.value)


class InsertState:
    """Preimages + tree, advanced one insert at a time (IMT:710-741)."""

    def __init__(self, pre, threads=1):
        self.pre = _c(pre).reshape(-1, 3, 4).copy()
        self.n = self.pre.shape[0]
        self.d = self.n.bit_length() - 1
        self.threads = threads
        self.tree = tree_build(hash3(self.pre, threads), threads)

    def root(self):
        return self.tree[-1].copy()

    def insert(self, new_val, new_idx, incremental=True, low_hint=None):
        d = self.d
        o = dict(old_root=np.zeros(4, np.uint64), low_idx=np.zeros(1, np.uint64), low_leaf=np.zeros((3, 4), np.uint64),
                 low_proof=np.zeros((d, 4), np.uint64), low_helper=np.zeros(d, np.uint8),
                 new_root=np.zeros(4, np.uint64), new_leaf=np.zeros((3, 4), np.uint64),
                 new_proof=np.zeros((d, 4), np.uint64), new_helper=np.zeros(d, np.uint8), is_largest=np.zeros(1, np.uint8))
        outs = [_p(o["old_root"]), _p(o["low_idx"]), _p(o["low_leaf"]), _p(o["low_proof"]), _p(o["low_helper"], _u8p),
                _p(o["new_root"]), _p(o["new_leaf"]), _p(o["new_proof"]), _p(o["new_helper"], _u8p), _p(o["is_largest"], _u8p)]
        head = [_p(self.pre), _p(self.tree), self.n, _p(_c(new_val)), new_idx]
        if incremental:
            rc = lib().imto_insert_round_incremental(*head, 0 if low_hint is None else low_hint, 0 if low_hint is None else 1, *outs)
        else:
            rc = lib().imto_insert_round_rebuild(*head, *outs, self.threads)
        if rc:
            raise ValueError(ERRORS[rc])
        o["low_idx"] = int(o["low_idx"][0])
        o["is_largest"] = int(o["is_largest"][0])
        return o


def synth_fe(seed, first, n):
    out = np.zeros((n, 4), np.uint64)
    lib().imto_synth_fe(seed, first, n, _p(out))
    return out


def convert(a, to_montgomery, threads=1):
    """dense FE array canonical <-> halo2curves' in-memory Montgomery form (x * 2^256 mod p)"""
    a = _c(a)
    out = np.empty_like(a)
    lib().imto_convert(_p(a), a.size // 4, _p(out), 1 if to_montgomery else 0, threads)
    return out


def build_from_preimages(pre, threads=1):
    a = _c(pre).reshape(-1, 3, 4)
    root = np.zeros(4, np.uint64)
    rc = lib().imto_build_from_preimages(_p(a), a.shape[0], _p(root), threads)
    if rc:
        raise ValueError(ERRORS[rc])
    return root

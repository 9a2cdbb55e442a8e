"""Host-side mirror of the reference's Rust API for this path, backed by the GPU library.

Same names, argument meaning and error behaviour as /root/reference/src/utils.rs, so the parity tests read like
the reference's own tests (src/indexed_merkle_tree.rs:360-803). Field elements are Python ints here (the way the
reference's tests write `Fr::from(30)`); everything is forwarded to the C-ABI as canonical 32-byte values.
The Rust binding a maintainer would add instead of this file is in INTEGRATION.md.
"""
from dataclasses import dataclass

import numpy as np

from .engine import Engine, ImtError, P, fe_from_int, fe_to_int, fes_from_ints, fes_to_ints
from . import _ffi

_engines = {}


def default_engine(t=3, rate=2, r_f=8, r_p=57):
    """one canonical-format engine per Poseidon instance, on device 0"""
    key = (t, rate, r_f, r_p)
    if key not in _engines:
        _engines[key] = Engine(0, "canonical", t=t, rate=rate, r_f=r_f, r_p=r_p)
    return _engines[key]


class Poseidon:
    """`Poseidon::<Fr, T, RATE>::new(r_f, r_p)` + `update` + `squeeze_and_reset` (indexed_merkle_tree.rs:370-376).

    The sponge state is only ever materialised on the GPU: `update` buffers, `squeeze_and_reset` runs the whole hash
    (len // RATE + 1 permutations) as one kernel. The reference's instance <3, 2>(8, 57) with its input lengths 2 (node
    hashing, utils.rs:46) and 3 (leaf hashing, indexed_merkle_tree.rs:374) takes the tuned kernels; any other length or
    instance (T in 2..5, RATE = T - 1) the any-width ones."""

    def __init__(self, r_f=8, r_p=57, engine=None, t=3, rate=2):
        self.T, self.RATE = t, rate
        self.engine = engine or default_engine(t, rate, r_f, r_p)
        if (self.engine.t, self.engine.rate, self.engine.r_f, self.engine.r_p) != (t, rate, r_f, r_p):
            raise ValueError("the engine was created for a different Poseidon instance")
        self._buf = []

    def update(self, elements):
        for e in elements:
            e = int(e)
            if not 0 <= e < P:
                raise ValueError("field element out of range")
            self._buf.append(e)

    def squeeze_and_reset(self):
        buf, self._buf = self._buf, []
        return fe_to_int(self.engine.hash(fes_from_ints(buf), len(buf))[0])


@dataclass
class IndexedMerkleTreeLeaf:
    """utils.rs:12-17 — field order val, next_val, next_idx"""
    val: int = 0
    next_val: int = 0
    next_idx: int = 0

    def as_tuple(self):
        return (self.val, self.next_val, self.next_idx)


class IndexedMerkleTree:
    """utils.rs:5-108. `new` raises ValueError with the reference's two messages (utils.rs:25, 35)."""

    def __init__(self, hash, leaves):  # noqa: A002 - the reference names the parameter `hash`
        self.hash = hash
        eng = hash.engine
        try:
            self._tree = eng.build_from_hashes(fes_from_ints([int(x) for x in leaves]))
        except ImtError as e:
            if e.status in (_ffi.ERR_EMPTY, _ffi.ERR_ODD):
                raise ValueError(str(e)) from None
            if e.status == _ffi.ERR_NOT_POW2:  # the reference panics with an index out of bounds at utils.rs:45
                raise IndexError(str(e)) from None
            raise
        self._n = len(leaves)
        self.root = fe_to_int(self._tree.root())

    @classmethod
    def new(cls, hash, leaves):  # noqa: A002
        return cls(hash, leaves)

    @property
    def tree(self):
        """tree: Vec<Vec<F>> (utils.rs:8)"""
        return [fes_to_ints(self._tree.level(l, self._n >> l)) for l in range(self._tree.depth + 1)]

    def get_root(self):
        return self.root

    def get_proof(self, index):
        if not 0 <= index < self._n:
            raise IndexError("index out of bounds")  # the reference panics at utils.rs:76
        sib, hel = self._tree.get_proofs(np.array([index], np.uint64), helpers_as_fe=True)
        return fes_to_ints(sib[0]), fes_to_ints(hel[0])

    def verify_proof(self, leaf, index, root, proof):
        eng = self.hash.engine
        ok = eng.verify_proofs(fes_from_ints([leaf]), np.array([index], np.uint64), fes_from_ints([root]),
                               fes_from_ints(list(proof)).reshape(1, -1, 4))
        return bool(ok[0])


def hash_nullifier_pre_images(leaves, engine=None):
    """indexed_merkle_tree.rs:662-671, batched: one kernel for all leaves."""
    eng = engine or default_engine()
    flat = [v for leaf in leaves for v in (leaf.as_tuple() if isinstance(leaf, IndexedMerkleTreeLeaf) else leaf)]
    return fes_to_ints(eng.hash3(fes_from_ints(flat).reshape(-1, 3, 4)))

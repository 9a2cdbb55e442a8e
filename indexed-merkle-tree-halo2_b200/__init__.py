"""imt_b200 — B200-native BN254 Poseidon / indexed Merkle tree engine.

Drop-in for the hot path of aerius-labs/indexed-merkle-tree-halo2: a C-ABI shared library of hand-written sm_100a
kernels (csrc/, include/imt_b200.h) plus this thin host layer:
  reference_api   the reference's own names (Poseidon, IndexedMerkleTree, IndexedMerkleTreeLeaf)
  engine          batched numpy/torch-facing calls (Engine, Tree)
  sharding        one-process-per-GPU subtree sharding over torch.distributed
  synth           deterministic synthetic leaves
The directory name contains '-', so import it through the repo-root shim:  `import imt_b200`.
"""
from . import build as _build  # noqa: F401
from . import _ffi  # noqa: F401
from .engine import Engine, Tree, Multi, MTree, ImtError, P, STATES_PER_HASH, fe_from_int, fe_to_int, fes_from_ints, fes_to_ints  # noqa: F401
from .reference_api import Poseidon, IndexedMerkleTree, IndexedMerkleTreeLeaf, hash_nullifier_pre_images, default_engine  # noqa: F401
from . import synth  # noqa: F401
from .sharding import ShardedTree  # noqa: F401


def build(force=False):
    """compile libimt_b200.so in-tree for sm_100a"""
    return _build.build(force=force)

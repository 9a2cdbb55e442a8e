"""ctypes binding of libimt_b200.so — the same C-ABI (include/imt_b200.h) a Rust `extern "C"` block would bind.

The library is the product; this file only marshals numpy / torch buffers into it. There is no CPU fallback: if the
library cannot be built or loaded, importing callers fail loudly."""
import ctypes
import os

from . import build as _build

c_void_p, c_size_t, c_int, c_uint, c_u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint, ctypes.c_uint64
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u64p = ctypes.POINTER(ctypes.c_uint64)

OK, ERR_EMPTY, ERR_ODD, ERR_NOT_POW2, ERR_INDEX_OOB, ERR_NON_CANONICAL, ERR_INVALID_ARG, ERR_TREE_FULL, ERR_NOT_WELL_FORMED = range(9)
ERR_CUDA = 100
FE_CANONICAL, FE_MONTGOMERY = 0, 1


class InsertWitness(ctypes.Structure):
    _fields_ = [("old_roots", c_void_p), ("low_idx", c_void_p), ("low_leaves", c_void_p), ("low_siblings", c_void_p),
                ("low_helpers", c_void_p), ("new_roots", c_void_p), ("new_leaves", c_void_p), ("new_siblings", c_void_p),
                ("new_helpers", c_void_p), ("is_largest", c_void_p), ("fold_nodes", c_void_p)]


# every symbol include/imt_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "imt_ctx_create": (c_int, [c_int, c_int, ctypes.POINTER(c_void_p)]),
    "imt_ctx_destroy": (None, [c_void_p]),
    "imt_last_error": (ctypes.c_char_p, [c_void_p]),
    "imt_status_string": (ctypes.c_char_p, [c_int]),
    "imt_ctx_launch_count": (c_u64, [c_void_p]),
    "imt_ctx_set_stream": (c_int, [c_void_p, c_void_p]),
    "imt_ctx_reset_stream": (c_int, [c_void_p]),
    "imt_ctx_trim": (c_int, [c_void_p]),
    "imt_ctx_enable_timing": (c_int, [c_void_p, c_int]),
    "imt_ctx_kernel_time": (c_int, [c_void_p, c_int, ctypes.POINTER(ctypes.c_double), c_u64p, c_u64p]),
    "imt_ctx_reset_timing": (c_int, [c_void_p]),
    "imt_tree_root_dev": (c_int, [c_void_p, c_void_p]),
    "imt_poseidon_hash2": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "imt_poseidon_hash3": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "imt_poseidon_hash2_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "imt_poseidon_hash3_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "imt_trace_hashes": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_void_p]),
    "imt_trace_hashes_dev": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_void_p]),
    "imt_tree_build_from_hashes": (c_int, [c_void_p, c_void_p, c_size_t, ctypes.POINTER(c_void_p)]),
    "imt_tree_build_from_leaves": (c_int, [c_void_p, c_void_p, c_size_t, ctypes.POINTER(c_void_p)]),
    "imt_tree_build_from_hashes_dev": (c_int, [c_void_p, c_void_p, c_size_t, ctypes.POINTER(c_void_p)]),
    "imt_tree_build_from_leaves_dev": (c_int, [c_void_p, c_void_p, c_size_t, ctypes.POINTER(c_void_p)]),
    "imt_tree_rebuild_from_leaves": (c_int, [c_void_p, c_void_p]),
    "imt_tree_rebuild_from_leaves_dev": (c_int, [c_void_p, c_void_p]),
    "imt_tree_destroy": (None, [c_void_p]),
    "imt_tree_num_leaves": (c_size_t, [c_void_p]),
    "imt_tree_depth": (c_uint, [c_void_p]),
    "imt_tree_root": (c_int, [c_void_p, c_void_p]),
    "imt_tree_level": (c_int, [c_void_p, c_uint, c_void_p]),
    "imt_tree_preimages": (c_int, [c_void_p, c_void_p]),
    "imt_tree_get_proofs": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "imt_tree_get_proofs_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "imt_verify_proofs_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_uint, c_void_p]),
    "imt_trace_merkle_proofs_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_uint, c_void_p, c_void_p]),
    "imt_low_leaf_lookup_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "imt_tree_get_proofs_fe": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "imt_verify_proofs": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_uint, c_void_p]),
    "imt_trace_merkle_proofs": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_uint, c_void_p, c_void_p]),
    "imt_low_leaf_lookup": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "imt_tree_occupied": (c_int, [c_void_p, ctypes.POINTER(c_size_t)]),
    "imt_non_inclusion_paths": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "imt_insert_batch": (c_int, [c_void_p, c_void_p, c_size_t, c_u64, ctypes.POINTER(InsertWitness)]),
    "imt_tree_trace_proofs": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "imt_tree_trace_proofs_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "imt_non_inclusion_limbs": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "imt_tree_subtree_root_dev": (c_int, [c_void_p, ctypes.POINTER(c_void_p)]),
    "imt_tree_attach_cap": (c_int, [c_void_p, c_uint, c_uint, c_void_p]),
    "imt_tree_attach_cap_dev": (c_int, [c_void_p, c_uint, c_uint, c_void_p]),
    "imt_tree_set_shard": (c_int, [c_void_p, c_uint, c_uint]),
    "imt_tree_shard_info": (c_int, [c_void_p, ctypes.POINTER(c_uint), ctypes.POINTER(c_uint), ctypes.POINTER(c_size_t)]),
    "imt_tree_head_next_zero": (c_int, [c_void_p, ctypes.POINTER(c_int)]),
    "imt_low_leaf_candidates": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "imt_low_leaf_merge": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint, c_size_t, c_u64, c_u64, c_int, c_void_p, c_void_p]),
    "imt_shard_insert_neighbors": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "imt_shard_insert_plan": (c_int, [c_void_p, c_void_p, c_size_t, c_u64, c_uint, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p]),
    "imt_shard_insert_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "imt_shard_insert_cap": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "imt_tree_leaves": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "imt_ctx_create_spec": (c_int, [c_int, c_int, c_uint, c_uint, c_uint, c_uint, ctypes.POINTER(c_void_p)]),
    "imt_ctx_spec": (c_int, [c_void_p, ctypes.POINTER(c_uint), ctypes.POINTER(c_uint), ctypes.POINTER(c_uint), ctypes.POINTER(c_uint),
                             ctypes.POINTER(c_int)]),
    "imt_trace_fe_per_hash": (c_int, [c_void_p, c_size_t, ctypes.POINTER(c_size_t)]),
    "imt_poseidon_hash": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_void_p]),
    "imt_poseidon_hash_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_void_p]),
    "imt_poseidon_trace": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_void_p]),
    "imt_poseidon_trace_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_void_p]),
    "imt_trace_sbox_fe_per_hash": (c_int, [c_void_p, c_size_t, ctypes.POINTER(c_size_t)]),
    "imt_poseidon_trace_ext": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_void_p, c_void_p]),
    "imt_poseidon_trace_ext_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_void_p, c_void_p]),
    "imt_tree_trace_proofs_ext": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "imt_tree_trace_proofs_ext_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "imt_poseidon_permute": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "imt_spec_params_host": (c_int, [c_uint, c_uint, c_uint, c_uint, c_void_p, c_size_t, ctypes.POINTER(c_size_t)]),
    "imt_calibrate_imad": (c_int, [c_void_p, ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    # ---- multi-GPU inside the library (NCCL)
    "imt_comm_unique_id": (c_int, [c_void_p]),
    "imt_comm_create": (c_int, [c_void_p, c_uint, c_uint, c_void_p]),
    "imt_comm_destroy": (c_int, [c_void_p]),
    "imt_comm_info": (c_int, [c_void_p, ctypes.POINTER(c_uint), ctypes.POINTER(c_uint), ctypes.POINTER(c_int)]),
    "imt_tree_exchange_roots": (c_int, [c_void_p]),
    "imt_sharded_build_from_leaves": (c_int, [c_void_p, c_void_p, c_size_t, ctypes.POINTER(c_void_p)]),
    "imt_sharded_build_from_leaves_dev": (c_int, [c_void_p, c_void_p, c_size_t, ctypes.POINTER(c_void_p)]),
    "imt_sharded_rebuild_from_leaves": (c_int, [c_void_p, c_void_p]),
    "imt_sharded_rebuild_from_leaves_dev": (c_int, [c_void_p, c_void_p]),
    "imt_sharded_get_proofs": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "imt_sharded_leaves": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "imt_sharded_low_leaf_lookup": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "imt_sharded_occupied": (c_int, [c_void_p, c_u64p]),
    "imt_sharded_non_inclusion_paths": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "imt_sharded_insert_batch": (c_int, [c_void_p, c_void_p, c_size_t, c_u64, ctypes.POINTER(InsertWitness)]),
    "imt_sharded_trace_proofs": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, ctypes.POINTER(c_size_t), c_void_p]),
    "imt_multi_create": (c_int, [ctypes.POINTER(c_int), c_uint, c_int, ctypes.POINTER(c_void_p)]),
    "imt_multi_destroy": (None, [c_void_p]),
    "imt_multi_size": (c_uint, [c_void_p]),
    "imt_multi_ctx": (c_void_p, [c_void_p, c_uint]),
    "imt_multi_last_error": (ctypes.c_char_p, [c_void_p]),
    "imt_multi_uses_nccl": (c_int, [c_void_p]),
    "imt_multi_build_from_leaves": (c_int, [c_void_p, c_void_p, c_size_t, ctypes.POINTER(c_void_p)]),
    "imt_mtree_rebuild_from_leaves": (c_int, [c_void_p, c_void_p]),
    "imt_mtree_destroy": (None, [c_void_p]),
    "imt_mtree_shard": (c_void_p, [c_void_p, c_uint]),
    "imt_mtree_num_leaves": (c_size_t, [c_void_p]),
    "imt_mtree_depth": (c_uint, [c_void_p]),
    "imt_mtree_root": (c_int, [c_void_p, c_void_p]),
    "imt_mtree_get_proofs": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "imt_mtree_leaves": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "imt_mtree_low_leaf_lookup": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "imt_mtree_occupied": (c_int, [c_void_p, c_u64p]),
    "imt_mtree_non_inclusion_paths": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "imt_mtree_insert_batch": (c_int, [c_void_p, c_void_p, c_size_t, c_u64, ctypes.POINTER(InsertWitness)]),
    "imt_mtree_trace_proofs": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "imt_fe_convert": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "imt_fe_convert_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "imt_insert_trace_hashes": (c_size_t, [c_uint]),
    "imt_non_inclusion_trace_hashes": (c_size_t, [c_uint]),
    "imt_non_inclusion_witness_trace": (c_int, [c_void_p, c_void_p, c_size_t] + [c_void_p] * 9),
    "imt_non_inclusion_witness_trace_dev": (c_int, [c_void_p, c_void_p, c_size_t] + [c_void_p] * 9),
    "imt_insert_witness_trace": (c_int, [c_void_p, ctypes.POINTER(InsertWitness), c_size_t, c_uint, c_u64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "imt_insert_witness_trace_dev": (c_int, [c_void_p, ctypes.POINTER(InsertWitness), c_size_t, c_uint, c_u64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    # ---- checkpoints
    "imt_tree_save": (c_int, [c_void_p, ctypes.c_char_p]),
    "imt_tree_load": (c_int, [c_void_p, ctypes.c_char_p, ctypes.POINTER(c_void_p)]),
    "imt_mtree_save": (c_int, [c_void_p, ctypes.c_char_p]),
    "imt_multi_load": (c_int, [c_void_p, ctypes.c_char_p, ctypes.POINTER(c_void_p)]),
    "imt_checkpoint_read_info": (c_int, [ctypes.c_char_p, c_void_p]),
}


class CheckpointInfo(ctypes.Structure):
    _fields_ = [("num_leaves", c_u64), ("version", c_uint), ("t", c_uint), ("rate", c_uint), ("r_f", c_uint), ("r_p", c_uint),
                ("depth", c_uint), ("root", ctypes.c_uint8 * 32)]

_lib = None


def library_path():
    return _build.LIB


def load(build_if_stale=True):
    """dlopen the C-ABI library and bind every declared symbol. Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("IMT_B200_LIB")  # A/B measurements of another build of the same library; normally unset
    if not path:
        if build_if_stale and _build.stale():
            _build.build()
        path = _build.LIB
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing and could not be built; the CUDA library is required (no CPU fallback)")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib

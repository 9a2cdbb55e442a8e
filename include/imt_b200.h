/*
 * imt_b200 — C-ABI of the B200-native BN254-Poseidon / indexed-Merkle-tree engine.
 *
 * Drop-in boundary for the hot path of aerius-labs/indexed-merkle-tree-halo2. Every entry point names the
 * reference interface it replaces (paths relative to /root/reference). The reference has no FFI of its own (it is
 * pure Rust); INTEGRATION.md shows the `extern "C"` block + build.rs a maintainer adds so that
 * `IndexedMerkleTree::new / get_root / get_proof / verify_proof` (src/utils.rs) and the witness loading in front
 * of `insert_leaf` / `verify_non_inclusion` (src/indexed_merkle_tree.rs:444-474) call into this library.
 *
 * Conventions
 *  - A field element (FE) is 32 bytes: BN254 Fr, little-endian. Format is per context:
 *      IMT_FE_CANONICAL   canonical integer < p          (== halo2curves `to_repr()` bytes)
 *      IMT_FE_MONTGOMERY  value * 2^256 mod p, < p       (== halo2curves' in-memory [u64; 4]; zero-copy from Rust)
 *    Inputs that are >= p are rejected with IMT_ERR_NON_CANONICAL.
 *  - A leaf preimage is 3 FE in the reference's struct order  val, next_val, next_idx  (src/utils.rs:12-17;
 *    hash order src/indexed_merkle_tree.rs:667).
 *  - Paths: siblings bottom-up; helper = 1 when the current node is the LEFT child (src/utils.rs:70, 79).
 *  - All arrays are dense, caller-allocated. `*_dev` variants take DEVICE pointers (same layout) and do not copy.
 *  - Every function returns an imt_status; nothing aborts the process. A context is not thread-safe (the
 *    reference's hasher is a `&mut` borrow: src/utils.rs:7, 21, 87). Calls are synchronous on return.
 *  - There is no CPU fallback: without a CUDA device imt_ctx_create fails with IMT_ERR_CUDA.
 */
#ifndef IMT_B200_H
#define IMT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct imt_ctx imt_ctx;
typedef struct imt_tree imt_tree;

typedef enum imt_status {
    IMT_OK = 0,
    IMT_ERR_EMPTY = 1,         /* "Cannot create Merkle Tree with no leaves"            src/utils.rs:24-26 */
    IMT_ERR_ODD = 2,           /* "Leaves must be even"                                 src/utils.rs:34-36 */
    IMT_ERR_NOT_POW2 = 3,      /* even but not a power of two: the reference panics at  src/utils.rs:45    */
    IMT_ERR_INDEX_OOB = 4,     /* index >= leaf count: the reference panics at          src/utils.rs:76    */
    IMT_ERR_NON_CANONICAL = 5, /* an input field element is >= p                                           */
    IMT_ERR_INVALID_ARG = 6,   /* null pointer, bad arity/format/level, tree without preimages, ...        */
    IMT_ERR_TREE_FULL = 7,     /* insert batch does not fit into the remaining empty slots                  */
    IMT_ERR_NOT_WELL_FORMED = 8, /* preimages are not a consistent indexed (sorted linked list) tree        */
    IMT_ERR_CUDA = 100         /* CUDA runtime / driver failure; text in imt_last_error()                   */
} imt_status;

typedef enum imt_fe_format { IMT_FE_CANONICAL = 0, IMT_FE_MONTGOMERY = 1 } imt_fe_format;

#define IMT_FE_BYTES 32
#define IMT_STATES_PER_HASH 132 /* 2 permutations x (1 + R_F + R_P) states of 3 FE: SURVEY.md 8a row 9 */

/* ---------------------------------------------------------------- context ------------------------------------ */
/* Replaces Poseidon::<Fr,3,2>::new(8, 57) (src/indexed_merkle_tree.rs:370, 663, 681, 807): derives the parameter
 * set (Grain LFSR, Cauchy MDS, optimized constants, sparse matrices) and uploads it to the device. */
imt_status imt_ctx_create(int device, imt_fe_format format, imt_ctx** out);
void imt_ctx_destroy(imt_ctx* ctx);
/* Text of the last failure on this context ("" if none). Valid until the next call on the context. */
const char* imt_last_error(const imt_ctx* ctx);
/* The message the reference returns for a status (exact strings of src/utils.rs:25, 35), or a description. */
const char* imt_status_string(imt_status st);
/* Number of kernels this context has launched so far (bench.py reports it as gpu_launches). */
uint64_t imt_ctx_launch_count(const imt_ctx* ctx);
/* Launch on a caller-owned CUDA stream (a cudaStream_t, e.g. torch's current stream) instead of the context's own
 * non-blocking stream, so that callers can bracket calls with their own events and order them against NCCL. NULL is
 * the CUDA legacy default stream. imt_ctx_reset_stream returns to the internal stream.
 * Device buffers passed to the _dev calls are read and written ON THE CONTEXT'S STREAM, which by default is not ordered against any
 * other stream: inputs produced on another stream (and every earlier use of recycled memory behind the outputs — a caching allocator
 * may hand out a block whose previous users are still queued on ITS stream) must be complete before the call, or the caller shares
 * its stream with the context through this function. Every call returns with its own work finished. */
imt_status imt_ctx_set_stream(imt_ctx* ctx, void* cuda_stream);
imt_status imt_ctx_reset_stream(imt_ctx* ctx);
/* Tree and scratch buffers are stream-ordered allocations from a memory pool that belongs to the context (the device's default
 * pool is not touched); freed memory stays cached there between calls (a depth-24 trace call can leave ~5 GB). imt_ctx_trim
 * returns the cached, unused part to the driver — for processes that share the GPU with another allocator. */
imt_status imt_ctx_trim(imt_ctx* ctx);
/* Per-kernel device timing: when enabled every hash launch is bracketed by CUDA events on its stream.
 * imt_ctx_kernel_time returns, for hash kernels of the given arity (2 = node levels, 3 = leaf hashing), the summed
 * device time in ms, the launch count and the number of hashes since the last reset. */
imt_status imt_ctx_enable_timing(imt_ctx* ctx, int enabled);
imt_status imt_ctx_kernel_time(imt_ctx* ctx, int arity, double* total_ms, uint64_t* launches, uint64_t* hashes);
imt_status imt_ctx_reset_timing(imt_ctx* ctx);

/* ---------------------------------------------------------------- batched hashing ---------------------------- */
/* out[i] = H(in[2i], in[2i+1]): update(&[l, r]) + squeeze_and_reset()   (src/utils.rs:46-47, 96-100)           */
imt_status imt_poseidon_hash2(imt_ctx* ctx, const void* in, size_t n, void* out);
/* out[i] = H(in[3i], in[3i+1], in[3i+2]): leaf hashing, order val,next_val,next_idx
 * (src/indexed_merkle_tree.rs:373-376, 407-415, 510-518, 662-671)                                              */
imt_status imt_poseidon_hash3(imt_ctx* ctx, const void* in, size_t n, void* out);
imt_status imt_poseidon_hash2_dev(imt_ctx* ctx, const void* d_in, size_t n, void* d_out);
imt_status imt_poseidon_hash3_dev(imt_ctx* ctx, const void* d_in, size_t n, void* d_out);

/* Dense field-element arrays between the two formats (canonical `to_repr()` bytes <-> halo2curves' in-memory Montgomery
 * form), n elements; inputs >= p are rejected. Independent of the context's own format. */
imt_status imt_fe_convert(imt_ctx* ctx, const void* in, size_t n, imt_fe_format from, imt_fe_format to, void* out);
imt_status imt_fe_convert_dev(imt_ctx* ctx, const void* d_in, size_t n, imt_fe_format from, imt_fe_format to, void* d_out);

/* Witness trace of `hash_fix_len_array` (src/indexed_merkle_tree.rs:92, 194, 271, 299): for each of the n hashes
 * of `arity` (2 or 3) inputs, the 132 x 3 FE states (per permutation: after the pre-constant add, then after the
 * linear layer of each of the 4 + 57 + 4 rounds) and the digest. states may be NULL. Other input lengths and
 * any-width contexts are forwarded to imt_poseidon_trace (state count: imt_trace_fe_per_hash). */
imt_status imt_trace_hashes(imt_ctx* ctx, const void* in, int arity, size_t n, void* states, void* digests);
imt_status imt_trace_hashes_dev(imt_ctx* ctx, const void* d_in, int arity, size_t n, void* d_states, void* d_digests);

/* ---------------------------------------------------------------- any-width instances ------------------------ */
/* The reference's tree and chip are generic over the Poseidon instance: `IndexedMerkleTree<'a, F, const T, const RATE>`
 * (src/utils.rs:5-10, 19), `verify_merkle_proof / verify_non_inclusion / insert_leaf<F, const T, const RATE>`
 * (src/indexed_merkle_tree.rs:65, 127, 231); its tests instantiate <3, 2> with R_F = 8, R_P = 57 only
 * (src/indexed_merkle_tree.rs:362-365), which is what imt_ctx_create gives (tuned kernels).
 * imt_ctx_create_spec replaces `Poseidon::<Fr, T, RATE>::new(r_f, r_p)` for any other instance: t in 2..5,
 * rate == t - 1, r_f even >= 2, r_f + r_p <= 256. EVERY entry point of this header then hashes with that instance
 * (tree build, leaf hashing, paths, folds, traces, inserts, sharding cap) on the any-width kernels; a witness trace has
 * (arity / rate + 1) x (1 + r_f + r_p) states of t FE per hash (imt_trace_fe_per_hash). */
imt_status imt_ctx_create_spec(int device, imt_fe_format format, unsigned t, unsigned rate, unsigned r_f, unsigned r_p, imt_ctx** out);
/* The instance of a context (any output pointer may be NULL); generic_kernels = 1 for imt_ctx_create_spec contexts. */
imt_status imt_ctx_spec(const imt_ctx* ctx, unsigned* t, unsigned* rate, unsigned* r_f, unsigned* r_p, int* generic_kernels);
/* FE per hash of `arity` inputs in a witness trace of this context (396 = 132 x 3 for <3, 2>(8, 57), arity 2 or 3). */
imt_status imt_trace_fe_per_hash(const imt_ctx* ctx, size_t arity, size_t* fe);
/* out[i] = { update(&in[arity*i .. arity*(i+1)]); squeeze_and_reset() } for ANY input length (arity >= 0): full
 * rate-chunks are absorbed and permuted, the remainder plus the padding element 1 once more (pse-poseidon's sponge,
 * call sites src/utils.rs:46-47, src/indexed_merkle_tree.rs:374-375). arity / rate + 1 permutations per hash. */
imt_status imt_poseidon_hash(imt_ctx* ctx, const void* in, size_t arity, size_t n, void* out);
imt_status imt_poseidon_hash_dev(imt_ctx* ctx, const void* d_in, size_t arity, size_t n, void* d_out);
/* imt_trace_hashes for any input length / instance: states = n x imt_trace_fe_per_hash(arity) FE (may be NULL). */
imt_status imt_poseidon_trace(imt_ctx* ctx, const void* in, size_t arity, size_t n, void* states, void* digests);
imt_status imt_poseidon_trace_dev(imt_ctx* ctx, const void* d_in, size_t arity, size_t n, void* d_states, void* d_digests);
/* EXTENDED witness trace (optional part): besides the per-round states, the three product cells the
 * in-circuit S-box materialises (x^2, x^4, x^5 + c — halo2-base computes x^5 + c as mul, mul, mul_add) for EVERY S-box in
 * execution order: per permutation r_f/2 full rounds x t lanes, r_p partial rounds x 1, r_f/2 full rounds x t lanes.
 * sbox = n x imt_trace_sbox_fe_per_hash(arity) FE (486 = 2 x 81 x 3 for <3, 2>(8, 57)); may be NULL. The column layout of
 * halo2-base itself is not reproduced (its source is not part of the reference). */
imt_status imt_trace_sbox_fe_per_hash(const imt_ctx* ctx, size_t arity, size_t* fe);
imt_status imt_poseidon_trace_ext(imt_ctx* ctx, const void* in, size_t arity, size_t n, void* states, void* sbox, void* digests);
imt_status imt_poseidon_trace_ext_dev(imt_ctx* ctx, const void* d_in, size_t arity, size_t n, void* d_states, void* d_sbox, void* d_digests);
/* The bare permutation on n states of t FE each, in place semantics (in -> out): what the published Poseidon test
 * vectors (poseidonperm_x5_254_3 / _5) are stated on. Not used by the tree. */
imt_status imt_poseidon_permute(imt_ctx* ctx, const void* in_states, size_t n, void* out_states);
/* Host-only: the parameter array `Poseidon::new` derives for an instance, Montgomery form, in the order
 * cap(2^64), one, pre[t], full[r_f][t], mds[t][t], pre_sparse[t][t], r_p x { c, row[t], col[t-1] }.
 * out == NULL just returns the element count. No device is touched. */
imt_status imt_spec_params_host(unsigned t, unsigned rate, unsigned r_f, unsigned r_p, void* out, size_t capacity_fe, size_t* count_fe);

/* ---------------------------------------------------------------- native tree -------------------------------- */
/* IndexedMerkleTree::new(hasher, leaves)  (src/utils.rs:20-57): all levels, bottom-up, kept on the device.
 * n == 0 -> IMT_ERR_EMPTY, n == 1 -> tree = [leaf], odd n -> IMT_ERR_ODD, other non powers of two -> NOT_POW2. */
imt_status imt_tree_build_from_hashes(imt_ctx* ctx, const void* leaf_hashes, size_t n, imt_tree** out);
/* Leaf hashing (src/indexed_merkle_tree.rs:662-671) fused in front of the build; keeps the preimages on the
 * device so that low-leaf lookups and inserts can follow. */
imt_status imt_tree_build_from_leaves(imt_ctx* ctx, const void* preimages, size_t n, imt_tree** out);
imt_status imt_tree_build_from_hashes_dev(imt_ctx* ctx, const void* d_leaf_hashes, size_t n, imt_tree** out);
imt_status imt_tree_build_from_leaves_dev(imt_ctx* ctx, const void* d_preimages, size_t n, imt_tree** out);
/* Re-run the build into an existing tree of the same size (no allocation): the steady-state call bench.py times. */
imt_status imt_tree_rebuild_from_leaves(imt_tree* tree, const void* preimages);
imt_status imt_tree_rebuild_from_leaves_dev(imt_tree* tree, const void* d_preimages);
/* A tree points at its context: destroy every tree BEFORE imt_ctx_destroy (the Rust wrapper's lifetime 'a does this). */
void imt_tree_destroy(imt_tree* tree);

size_t imt_tree_num_leaves(const imt_tree* tree);
unsigned imt_tree_depth(const imt_tree* tree); /* number of sibling levels = log2(n) */
/* get_root()  (src/utils.rs:59-61) */
imt_status imt_tree_root(imt_tree* tree, void* out_fe);
/* The same into a DEVICE buffer of 32 bytes (context format); asynchronous on the context's stream. */
imt_status imt_tree_root_dev(imt_tree* tree, void* d_out_fe);
/* tree[level] (src/utils.rs:8): level 0 = leaf hashes, level depth = [root]. Copies (n >> level) FE. */
imt_status imt_tree_level(imt_tree* tree, unsigned level, void* out);
/* Current preimages (n x 3 FE) of a tree built from leaves. */
imt_status imt_tree_preimages(imt_tree* tree, void* out);

/* get_proof(index) batched  (src/utils.rs:63-85): siblings[q][depth] FE, helpers[q][depth] bytes (1 = left). */
imt_status imt_tree_get_proofs(imt_tree* tree, const uint64_t* indices, size_t q, void* siblings, uint8_t* helpers);
/* Device-pointer variant (indices, siblings, helper bytes all in HBM): for provers that consume witnesses on the GPU. */
imt_status imt_tree_get_proofs_dev(imt_tree* tree, const uint64_t* d_indices, size_t q, void* d_siblings, uint8_t* d_helpers);
/* The same helpers as field elements, as the reference returns them (F::from(1) / F::from(0), src/utils.rs:79). */
imt_status imt_tree_get_proofs_fe(imt_tree* tree, const uint64_t* indices, size_t q, void* siblings, void* helpers_fe);

/* verify_proof(leaf, index, root, proof) batched  (src/utils.rs:87-107): ok[i] = 1 iff the fold equals roots[i].
 * roots holds q FE (one per query). */
imt_status imt_verify_proofs(imt_ctx* ctx, const void* leaves, const uint64_t* indices, const void* roots,
                             const void* siblings, size_t q, unsigned depth, uint8_t* ok);

imt_status imt_verify_proofs_dev(imt_ctx* ctx, const void* d_leaves, const uint64_t* d_indices, const void* d_roots,
                                 const void* d_siblings, size_t q, unsigned depth, uint8_t* d_ok);

/* Witness trace of compute_merkle_root (src/indexed_merkle_tree.rs:78-96): for each query the `depth` hashes of
 * the fold, in order, each as 132 x 3 FE; plus the computed roots. states[q][depth][132][3] FE. */
imt_status imt_trace_merkle_proofs(imt_ctx* ctx, const void* leaves, const uint64_t* indices, const void* siblings,
                                   size_t q, unsigned depth, void* states, void* roots);

imt_status imt_trace_merkle_proofs_dev(imt_ctx* ctx, const void* d_leaves, const uint64_t* d_indices, const void* d_siblings,
                                       size_t q, unsigned depth, void* d_states, void* d_roots);

/* The same traces for leaves OF A TREE, by index: get_proof (src/utils.rs:63-85) + the verify_merkle_proof witness
 * (src/indexed_merkle_tree.rs:65-96) in one call. Every operand of every hash is a stored node, so the q x depth traced
 * hashes run independently on the device instead of as q serial folds. states[q][depth][fe per hash] FE — byte-identical
 * to imt_tree_get_proofs followed by imt_trace_merkle_proofs on the tree's leaf hashes. */
imt_status imt_tree_trace_proofs(imt_tree* tree, const uint64_t* indices, size_t q, void* states);
imt_status imt_tree_trace_proofs_dev(imt_tree* tree, const uint64_t* d_indices, size_t q, void* d_states);
/* ... with the extended S-box trace (see imt_poseidon_trace_ext): sbox[q][depth][sbox fe per hash], may be NULL. The host
 * variant sizes its device buffers for the whole batch. */
imt_status imt_tree_trace_proofs_ext(imt_tree* tree, const uint64_t* indices, size_t q, void* states, void* sbox);
imt_status imt_tree_trace_proofs_ext_dev(imt_tree* tree, const uint64_t* d_indices, size_t q, void* d_states, void* d_sbox);

/* ---------------------------------------------------------------- indexed-leaf logic ------------------------- */
/* Low-leaf (predecessor) lookup, the read-only half of update_idx_leaf (src/indexed_merkle_tree.rs:632-660):
 * low_idx[i] = first slot with  val < v && (next_val > v || next_val == 0)  (or slot 0 for the very first insert).
 * matched[i] = 0 where no slot qualifies (v == 0 or v already present with no empty slot): the reference then
 * returns index 0 and leaves the leaves unchanged. Requires a tree built from leaves whose occupied slots form a
 * prefix and a consistent sorted linked list (IMT_ERR_NOT_WELL_FORMED otherwise). */
imt_status imt_low_leaf_lookup(imt_tree* tree, const void* values, size_t q, uint64_t* low_idx, uint8_t* matched);
imt_status imt_low_leaf_lookup_dev(imt_tree* tree, const void* d_values, size_t q, uint64_t* d_low_idx, uint8_t* d_matched);
/* Number of occupied slots (they form the prefix [0, occupied)): the slot the next insert goes to (IMT:733). */
imt_status imt_tree_occupied(imt_tree* tree, size_t* occupied);
/* Non-inclusion witnesses for verify_non_inclusion (src/indexed_merkle_tree.rs:127-137): lookup + the low leaf's
 * preimage (3 FE), its path, and is_largest = (low.next_val == 0). Any output pointer may be NULL. */
imt_status imt_non_inclusion_paths(imt_tree* tree, const void* values, size_t q, uint64_t* low_idx, uint8_t* matched,
                                   void* low_leaves, void* siblings, uint8_t* helpers, uint8_t* is_largest);

/* Sequential inserts with per-insert witnesses — the native half of test_insert_leaf_multiple_round
 * (src/indexed_merkle_tree.rs:710-741) — computed with O(depth) hashes per insert instead of a full re-hash and
 * rebuild. Insert i puts new_vals[i] into slot first_idx + i. Outputs (any may be NULL), per insert:
 *   old_roots FE, low_idx u64, low_leaves 3 FE (OLD preimage), low_siblings depth FE + low_helpers (OLD tree),
 *   new_roots FE, new_leaves 3 FE, new_siblings depth FE + new_helpers (NEW tree), is_largest u8,
 *   fold_nodes 4 x depth FE = the chain values of the four folds insert_leaf constrains (:196-204, 277-294, 305-313), levels
 *   0 .. depth-1 each: [0] the low leaf's path BEFORE the insert (level 0 = H3(low leaf)), [1] the same path after the low
 *   leaf was rewired, [2] the new leaf's path before the new leaf is written (level 0 = the stored empty leaf), [3] after.
 *   They cost nothing extra here (the batch computes them anyway) and turn imt_insert_witness_trace into one launch. Single-GPU
 *   trees only: imt_sharded_insert_batch / imt_mtree_insert_batch do not write fold_nodes (pass NULL there).
 * first_idx must equal imt_tree_occupied(). The batch is validated first (IMT_ERR_INVALID_ARG for a value that is 0,
 * already in the tree or repeated; IMT_ERR_TREE_FULL): on a validation error nothing is modified. The tree, its
 * preimages and its sorted index are updated in place, chunk by chunk (up to 65536 inserts each): a CUDA / allocation
 * failure (IMT_ERR_CUDA) in a later chunk leaves the earlier chunks applied — imt_tree_occupied() tells how far the
 * tree got, and the witnesses of those inserts are valid. */
typedef struct imt_insert_witness {
    void* old_roots;
    uint64_t* low_idx;
    void* low_leaves;
    void* low_siblings;
    uint8_t* low_helpers;
    void* new_roots;
    void* new_leaves;
    void* new_siblings;
    uint8_t* new_helpers;
    uint8_t* is_largest;
    void* fold_nodes;
} imt_insert_witness;
imt_status imt_insert_batch(imt_tree* tree, const void* new_vals, size_t b, uint64_t first_idx, imt_insert_witness* w);

/* The 128-bit limb witnesses of verify_non_inclusion (src/indexed_merkle_tree.rs:143-172, 206-222; is_less_than :98-125),
 * batched: for each of the b pairs (low_leaves[3i..3i+3) = val,next_val,next_idx ; new_vals[i]) writes 6 FE
 *   nl_q, nl_r, ll_q, ll_r, llv_q, llv_r      (x = x_q * 2^128 + x_r; nl = new value, ll = low.next_val, llv = low.val)
 * in the order the chip loads them, and (flags may be NULL) 3 bytes: nl < ll (is_next_val_greater, :178),
 * llv < nl (check_less_than, :226), and whether the pair passes both prover-side assertions (:180-191 with
 * is_new_leaf_largest = (low.next_val == 0), :226-228) — a witness that fails them makes the chip panic before MockProver. */
imt_status imt_non_inclusion_limbs(imt_ctx* ctx, const void* low_leaves, const void* new_vals, size_t b, void* limbs, uint8_t* flags);

/* The Poseidon witness trace of everything the chip's insert_leaf hashes (src/indexed_merkle_tree.rs:231-314), for a whole
 * batch in ONE call: per insert 3 + 4 x depth hashes in the chip's call order —
 *     [0] H3(low leaf before)                      :193-194      [1 .. d]          its fold up the low path -> old root    :196-204
 *     [d+1] H3(low.val, new.val, new_idx)          :265-275      [d+2 .. 2d+1]     its fold up the SAME path -> interim    :277-284
 *     [2d+2 .. 3d+1] fold of the empty leaf H3(0,0,0) (a constant in the chip, :247-251) up the new leaf's path -> interim :286-294
 *     [3d+2] H3(new leaf)                          :299-303      [3d+3 .. 4d+2]    its fold up the same path -> new root   :305-313
 * each as imt_trace_fe_per_hash states (132 x 3 FE). Input: the witnesses imt_insert_batch returned for the batch (low_leaves,
 * low_idx, low_siblings, new_leaves, new_siblings must be set; the rest of `w` is not read) and the slot of its first insert.
 * Outputs (any may be NULL): states[b][3 + 4 depth][132][3] FE; roots[b][4] FE = the roots the four folds end in (old,
 * interim, interim again via the empty leaf, new); new_low_leaves[b][3] FE = the rewired low leaf's preimage; limbs[b][6] FE +
 * limb_flags[b][3] = imt_non_inclusion_limbs of (low leaf, new value). With w->fold_nodes (as imt_insert_batch returned them)
 * every operand of every hash is known up front and ALL b x (3 + 4 depth) traced hashes run as independent threads of one
 * launch (the multiply-pipe rate of the trace kernel); every digest is compared with the next chain value, so fold_nodes that
 * do not belong to these witnesses give IMT_ERR_INVALID_ARG, never a silently inconsistent trace. Without fold_nodes the four folds of all inserts advance level by level, one traced launch
 * of 4b hashes per level (1 + depth dependent launches: hash latency for small b). The _dev variant takes DEVICE pointers in
 * `w` and for every output. Default instance only (any-width contexts: compose imt_poseidon_trace / imt_trace_merkle_proofs). */
size_t imt_insert_trace_hashes(unsigned depth); /* 3 + 4 depth */
imt_status imt_insert_witness_trace(imt_ctx* ctx, const imt_insert_witness* w, size_t b, unsigned depth, uint64_t first_idx, void* states,
                                    void* roots, void* new_low_leaves, void* limbs, uint8_t* limb_flags);
imt_status imt_insert_witness_trace_dev(imt_ctx* ctx, const imt_insert_witness* d_w, size_t b, unsigned depth, uint64_t first_idx,
                                        void* d_states, void* d_roots, void* d_new_low_leaves, void* d_limbs, uint8_t* d_limb_flags);

/* The whole witness of the chip's verify_non_inclusion (src/indexed_merkle_tree.rs:127-229) for q values in ONE call: the low-leaf
 * lookup, everything imt_non_inclusion_paths returns, the 128-bit limb witnesses of imt_non_inclusion_limbs (:143-172, 206-222) and
 * the Poseidon states of the 1 + depth hashes it constrains, in its call order:
 *     [0] H3(low leaf)   :193-194        [1 .. depth] its fold up the low leaf's path to the root (compute_merkle_root :65-96, :196-204)
 * states[q][1 + depth][imt_trace_fe_per_hash] FE. Every output may be NULL in the host variant; the _dev variant takes DEVICE pointers
 * and needs d_low_idx and d_low_leaves (they feed the trace). All operands are stored nodes, so the q x (1 + depth) traced hashes are
 * independent threads of two launches (the multiply-pipe rate of the trace kernel); the host variant streams the states out in
 * chunks behind the hashing. matched[i] = 0 (value 0 or already present) still yields the witness of slot low_idx[i] = 0, as the
 * reference's helper returns it; limb_flags[3 i + 2] tells whether the chip's prover-side assertions would pass. Single-GPU trees
 * built from leaves, default Poseidon instance (any-width contexts: imt_non_inclusion_paths + imt_poseidon_trace +
 * imt_tree_trace_proofs). */
size_t imt_non_inclusion_trace_hashes(unsigned depth); /* 1 + depth */
imt_status imt_non_inclusion_witness_trace(imt_tree* tree, const void* values, size_t q, uint64_t* low_idx, uint8_t* matched, void* low_leaves,
                                           void* siblings, uint8_t* helpers, uint8_t* is_largest, void* limbs, uint8_t* limb_flags,
                                           void* states);
imt_status imt_non_inclusion_witness_trace_dev(imt_tree* tree, const void* d_values, size_t q, uint64_t* d_low_idx, uint8_t* d_matched,
                                               void* d_low_leaves, void* d_siblings, uint8_t* d_helpers, uint8_t* d_is_largest, void* d_limbs,
                                               uint8_t* d_limb_flags, void* d_states);

/* ---------------------------------------------------------------- subtree sharding (one process per GPU) ----- */
/* A depth-d tree over N = 2^k ranks: rank g owns leaves [g n/N, (g+1) n/N) and builds that subtree with the calls
 * above (n = leaves per rank). The N subtree roots are exchanged by the caller (ncclAllGather / torch.distributed
 * all_gather of N x 32 bytes) and attached here; the k cap levels are then built on this rank's device.
 * Afterwards root / get_proofs / depth refer to the GLOBAL tree: indices passed to get_proofs are global and must
 * fall inside this rank's range. `d_subtree_root` from imt_tree_subtree_root_dev is a device pointer to this
 * rank's root (Montgomery-format contexts only), usable directly as the all-gather send buffer. A rebuild makes the
 * attached cap stale: root()/depth() then refer to the subtree again until the new roots are attached. */
imt_status imt_tree_subtree_root_dev(imt_tree* tree, const void** d_subtree_root);
imt_status imt_tree_attach_cap(imt_tree* tree, unsigned rank, unsigned world, const void* subtree_roots);
imt_status imt_tree_attach_cap_dev(imt_tree* tree, unsigned rank, unsigned world, const void* d_subtree_roots);

/* Indexed-leaf lookups on a sharded tree. The preimages of a shard carry GLOBAL slot numbers in next_idx, and the
 * occupied slots are a global prefix, so each shard's occupied slots are a local prefix. imt_tree_set_shard declares
 * the shard (attach_cap does the same) so that the sorted index of this rank reports global slots; the linked-list
 * consistency check of imt_low_leaf_lookup needs the whole list and is skipped on shards (distinctness is kept).
 *   1. every rank: imt_low_leaf_candidates -> for each query value the largest LOCAL key below it (cand_keys: canonical
 *      integers, NOT the context format), its global slot, flags (bit 0: candidate exists, bit 1: value present locally)
 *   2. the caller all-gathers the three arrays ([world][q], rank-major) and the per-rank imt_tree_occupied counts
 *   3. any rank: imt_low_leaf_merge -> low_idx / matched, identical to imt_low_leaf_lookup on the unsharded tree
 *      (head_next_zero = imt_tree_head_next_zero of rank 0: the reference's first-insert branch, IMT:640)
 *   4. the owner of each low_idx serves its preimage (imt_tree_leaves) and path (imt_tree_get_proofs), global indices. */
imt_status imt_tree_set_shard(imt_tree* tree, unsigned rank, unsigned world);
/* rank / world of a shard (0 / 1 for an unsharded tree) and the leaves this rank holds; any pointer may be NULL */
imt_status imt_tree_shard_info(const imt_tree* tree, unsigned* rank, unsigned* world, size_t* n_local);
imt_status imt_tree_head_next_zero(imt_tree* tree, int* flag);
imt_status imt_low_leaf_candidates(imt_tree* tree, const void* values, size_t q, void* cand_keys, uint64_t* cand_slots, uint8_t* flags);
imt_status imt_low_leaf_merge(imt_ctx* ctx, const void* values, const void* cand_keys, const uint64_t* cand_slots, const uint8_t* flags,
                              unsigned world, size_t q, uint64_t occupied_total, uint64_t n_total, int head_next_zero,
                              uint64_t* low_idx, uint8_t* matched);
/* Insert batches on a sharded tree (<= 4096 inserts per round of calls; next_idx / slots are GLOBAL). Same result as
 * imt_insert_batch on the unsharded tree — the reference's sequence of rebuilds (src/indexed_merkle_tree.rs:710-741):
 *   1. every rank: imt_shard_insert_neighbors -> each value's neighbours in the rank's own sorted index (keys as
 *      canonical integers; flags bit 0 predecessor exists, bit 1 value present, bit 2 successor exists)
 *   2. all-gather ([world][b], rank-major); every rank: imt_shard_insert_plan -> the replicated plan: x[2b] = slot of
 *      write t (t = 2k low leaf of insert k, 2k+1 its new leaf), upd[2b][3] = preimage each write stores,
 *      low_old[b][3] = low leaves before (IMT:720), is_largest[b]. IMT_ERR_INVALID_ARG for a zero / present / repeated value.
 *   3. every rank: imt_shard_insert_apply -> hashes and applies ITS writes to its subtree (levels, preimages, index);
 *      sub_roots[2b] = subtree root after each own write, sib_local[2b][log2 n] = its path inside the subtree (zeros
 *      for other ranks' writes)
 *   4. all-gather (each write has one owner); every rank: imt_shard_insert_cap -> applies all writes to the replicated
 *      cap: roots[2b] = global root after every write, sib_cap[2b][log2 world] = the cap part of every path. */
imt_status imt_shard_insert_neighbors(imt_tree* tree, const void* values, size_t b, void* pred_keys, uint64_t* pred_slots,
                                      void* succ_keys, uint64_t* succ_slots, uint8_t* flags);
imt_status imt_shard_insert_plan(imt_ctx* ctx, const void* values, size_t b, uint64_t first_idx, unsigned world, const void* pred_keys,
                                 const uint64_t* pred_slots, const void* succ_keys, const uint64_t* succ_slots, const uint8_t* flags,
                                 uint64_t* x, void* upd, void* low_old, uint8_t* is_largest);
imt_status imt_shard_insert_apply(imt_tree* tree, const uint64_t* x, const void* upd, size_t b, void* sub_roots, void* sib_local);
imt_status imt_shard_insert_cap(imt_tree* tree, const uint64_t* x, const void* sub_roots, size_t b, void* roots, void* sib_cap);
/* Preimages (3 FE each, context format) of the given slots and is_largest = (next_val == 0). Global indices inside this
 * rank's range for a shard. Either output may be NULL. */
imt_status imt_tree_leaves(imt_tree* tree, const uint64_t* indices, size_t q, void* leaves, uint8_t* is_largest);

/* ---------------------------------------------------------------- multi-GPU inside the library (NCCL) --------- */
/* The calls above leave the exchange of the subtree roots to the caller. The calls below do it INSIDE the library with NCCL
 * (bound at run time: libnccl.so.2 is only needed once a communicator is made), so that the reference's single entry point
 * `IndexedMerkleTree::new(&mut hasher, leaves)` (src/utils.rs:20-57) maps to ONE call on a multi-GPU box:
 *     N local subtree builds  ->  one ncclAllGather of N x 32 bytes over NVLink  ->  log2 N cap levels on every rank.
 * Results are bit-identical to the single-GPU tree (and to the reference). Two ways to form the group of ranks:
 *
 * (1) ONE PROCESS PER GPU (torchrun / MPI layout). Rank 0 calls imt_comm_unique_id and passes the IMT_COMM_ID_BYTES bytes
 *     to every rank by any means (file, socket, MPI_Bcast, a torch.distributed store); every rank attaches a communicator
 *     to its own context. world must be a power of two. The imt_sharded_* calls are collective: every rank calls them
 *     with the same arguments (query arrays are replicated inputs, results are replicated outputs). */
#define IMT_COMM_ID_BYTES 128
imt_status imt_comm_unique_id(void* id);
imt_status imt_comm_create(imt_ctx* ctx, unsigned rank, unsigned world, const void* id);
imt_status imt_comm_destroy(imt_ctx* ctx); /* before imt_ctx_destroy; trees of the context first */
/* rank / world of the context's group (0 / 1 without one) and the NCCL version in use (any pointer may be NULL) */
imt_status imt_comm_info(const imt_ctx* ctx, unsigned* rank, unsigned* world, int* nccl_version);
/* Exchange after a local (re)build made with the single-GPU calls: all-gathers the subtree roots straight out of / into
 * the trees' device buffers on the context's stream and builds the cap; afterwards root / depth / get_proofs refer to the
 * GLOBAL tree (as after imt_tree_attach_cap). */
imt_status imt_tree_exchange_roots(imt_tree* tree);
/* IndexedMerkleTree::new for this rank's contiguous slice of the leaves (n_local = n / world), exchange included. */
imt_status imt_sharded_build_from_leaves(imt_ctx* ctx, const void* local_preimages, size_t n_local, imt_tree** out);
imt_status imt_sharded_build_from_leaves_dev(imt_ctx* ctx, const void* d_local_preimages, size_t n_local, imt_tree** out);
imt_status imt_sharded_rebuild_from_leaves(imt_tree* tree, const void* local_preimages);
imt_status imt_sharded_rebuild_from_leaves_dev(imt_tree* tree, const void* d_local_preimages);
/* get_proof (src/utils.rs:63-85) for GLOBAL indices: the owner of each leaf serves its path, one all-reduce assembles the
 * replicated result. Same outputs as imt_tree_get_proofs on the unsharded tree. */
imt_status imt_sharded_get_proofs(imt_tree* tree, const uint64_t* indices, size_t q, void* siblings, uint8_t* helpers);
/* imt_tree_leaves for GLOBAL indices, replicated. */
imt_status imt_sharded_leaves(imt_tree* tree, const uint64_t* indices, size_t q, void* leaves, uint8_t* is_largest);
/* update_idx_leaf's scan (src/indexed_merkle_tree.rs:632-660) over the sharded tree: per-rank predecessor candidates, one
 * all-gather on the device, the same decision as imt_low_leaf_lookup. occupied_total may be NULL. */
imt_status imt_sharded_low_leaf_lookup(imt_tree* tree, const void* values, size_t q, uint64_t* low_idx, uint8_t* matched);
imt_status imt_sharded_occupied(imt_tree* tree, uint64_t* occupied_total);
/* imt_non_inclusion_paths over the sharded tree (any output pointer may be NULL). */
imt_status imt_sharded_non_inclusion_paths(imt_tree* tree, const void* values, size_t q, uint64_t* low_idx, uint8_t* matched,
                                           void* low_leaves, void* siblings, uint8_t* helpers, uint8_t* is_largest);
/* imt_insert_batch over the sharded tree: same witnesses, same final state as on one GPU — the reference's sequence of
 * rebuilds (src/indexed_merkle_tree.rs:710-741). Insert i goes to GLOBAL slot first_idx + i; first_idx must equal the total
 * number of occupied slots (imt_sharded_occupied). */
imt_status imt_sharded_insert_batch(imt_tree* tree, const void* new_vals, size_t b, uint64_t first_idx, imt_insert_witness* w);
/* Witness traces of verify_merkle_proof (imt_tree_trace_proofs) sharded by leaf OWNER with no exchange: this rank traces the
 * queries whose leaves it owns. positions[j] (capacity q) = index into `indices` of the j-th traced query, *n_mine = how
 * many; states[j] = its trace. The traces stay on the rank that produced them (19.9 GB for 2^16 depth-24 paths). */
imt_status imt_sharded_trace_proofs(imt_tree* tree, const uint64_t* indices, size_t q, uint64_t* positions, size_t* n_mine, void* states);

/* (2) ONE PROCESS DRIVES N GPUs — the natural shape for the Rust host, whose `IndexedMerkleTree::new` is one call on one
 *     thread. imt_multi_create makes one context per listed device (a power of two of them) and their communicators
 *     (ncclCommInitAll); an imt_mtree is the sharded tree: shard i (leaves [i n/N, (i+1) n/N)) lives on devices[i].
 *     Host arrays are whole-tree arrays; every call fans out to the devices and returns the assembled result. A device may be
 *     listed more than once (one-GPU test boxes): NCCL refuses duplicate devices, so such a group exchanges by
 *     device-to-device copies instead — everything else is the same code. */
typedef struct imt_multi imt_multi;
typedef struct imt_mtree imt_mtree;
imt_status imt_multi_create(const int* devices, unsigned n_dev, imt_fe_format format, imt_multi** out);
void imt_multi_destroy(imt_multi* m); /* destroy its trees first */
unsigned imt_multi_size(const imt_multi* m);
imt_ctx* imt_multi_ctx(imt_multi* m, unsigned i); /* context of device i: batched hashing, folds, traces, timing, ... */
const char* imt_multi_last_error(const imt_multi* m);
int imt_multi_uses_nccl(const imt_multi* m); /* the NCCL version code in use, 0 = copy transport */
/* IndexedMerkleTree::new (src/utils.rs:20-57) + leaf hashing (src/indexed_merkle_tree.rs:662-671) over all devices: the N
 * host->device pipelines, leaf kernels and level launches are queued on all devices before anything is waited for. */
imt_status imt_multi_build_from_leaves(imt_multi* m, const void* preimages, size_t n, imt_mtree** out);
imt_status imt_mtree_rebuild_from_leaves(imt_mtree* tree, const void* preimages);
void imt_mtree_destroy(imt_mtree* tree);
imt_tree* imt_mtree_shard(imt_mtree* tree, unsigned i); /* shard i as an imt_tree (levels, preimages, device-pointer calls) */
size_t imt_mtree_num_leaves(const imt_mtree* tree);
unsigned imt_mtree_depth(const imt_mtree* tree);
imt_status imt_mtree_root(imt_mtree* tree, void* out_fe);
imt_status imt_mtree_get_proofs(imt_mtree* tree, const uint64_t* indices, size_t q, void* siblings, uint8_t* helpers);
imt_status imt_mtree_leaves(imt_mtree* tree, const uint64_t* indices, size_t q, void* leaves, uint8_t* is_largest);
imt_status imt_mtree_low_leaf_lookup(imt_mtree* tree, const void* values, size_t q, uint64_t* low_idx, uint8_t* matched);
imt_status imt_mtree_occupied(imt_mtree* tree, uint64_t* occupied_total);
imt_status imt_mtree_non_inclusion_paths(imt_mtree* tree, const void* values, size_t q, uint64_t* low_idx, uint8_t* matched,
                                         void* low_leaves, void* siblings, uint8_t* helpers, uint8_t* is_largest);
imt_status imt_mtree_insert_batch(imt_mtree* tree, const void* new_vals, size_t b, uint64_t first_idx, imt_insert_witness* w);
/* imt_tree_trace_proofs for GLOBAL indices: every device traces the queries it owns and drains them over its own PCIe link
 * into states[q][depth][fe per hash] (caller order). */
imt_status imt_mtree_trace_proofs(imt_mtree* tree, const uint64_t* indices, size_t q, void* states);

/* ---------------------------------------------------------------- checkpoints -------------------------------- */
/* A built tree on disk, readable by a serde consumer of the reference's native leaf struct (src/utils.rs:12-17): a 64-byte
 * header, then n x 96 bytes = the leaves as `val, next_val, next_idx`, each the canonical 32-byte little-endian `to_repr()`
 * bytes (whatever the context's format), then the 32-byte canonical root. Layout:
 *     0 "IMTB200\0" | 8 u32 version = 1 | 12 u32 t | 16 u32 rate | 20 u32 r_f | 24 u32 r_p | 28 u32 depth | 32 u64 n | 40 zero[24]
 *     64 leaves[n][3][32] | 64 + 96 n root[32]
 * The levels are not stored: imt_tree_load re-hashes the leaves on the GPU and fails with IMT_ERR_INVALID_ARG ("checkpoint
 * is corrupt") when the rebuilt root differs from the stored one; leaves >= p give IMT_ERR_NON_CANONICAL. */
typedef struct imt_checkpoint_info {
    uint64_t num_leaves;
    uint32_t version, t, rate, r_f, r_p, depth;
    uint8_t root[32];
} imt_checkpoint_info;
imt_status imt_tree_save(imt_tree* tree, const char* path);
imt_status imt_tree_load(imt_ctx* ctx, const char* path, imt_tree** out);
/* the same file from / into a tree sharded over the devices of an imt_multi (shards are written in rank order = leaf order) */
imt_status imt_mtree_save(imt_mtree* tree, const char* path);
imt_status imt_multi_load(imt_multi* m, const char* path, imt_mtree** out);
/* header + root of a checkpoint; touches no device */
imt_status imt_checkpoint_read_info(const char* path, imt_checkpoint_info* info);

/* ---------------------------------------------------------------- calibration -------------------------------- */
/* Integer-multiply roofline calibration: saturates every SM with IMAD.WIDE.U32.X carry chains (the instruction the
 * field arithmetic issues) for about `ms` milliseconds per launch and returns the sustained 32x32->64
 * multiply-accumulates per second; implied_clock_mhz = rate / (32 lanes x SMs), i.e. the SM clock this rate
 * corresponds to if the pipe retires its nominal 32 wide MACs per clock per SM. */
imt_status imt_calibrate_imad(imt_ctx* ctx, double ms, double* wide_mac_per_s, double* sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* IMT_B200_H */

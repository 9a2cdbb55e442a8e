// Internal declarations shared by the translation units of libimt_b200.so (not part of the C-ABI).
//   imt_capi.cu     context, batched hashing, tree build, paths, folds, traces, sharding cap, calibration
//   imt_indexed.cu  sorted-key index, low-leaf lookups, non-inclusion witnesses, batched inserts
//   imt_spec.cu     any-width Poseidon instances (T = 2..5, run-time r_f / r_p / input length): kernels + entry points
//   imt_comm.cu     NCCL inside the library: communicators (one process per GPU, or one process driving N GPUs), the
//                   root exchange of a sharded build, the collectives the sharded lookups / inserts use
// All hashing kernels live in imt_capi.cu (one __constant__ copy of the Poseidon parameters); imt_indexed.cu prepares
// operands and calls them through imt_host::launch_hash / launch_level.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "imt_b200.h"
#include "poseidon.cuh"
#include "poseidon_spec.cuh"

struct imt_group;  // the ranks of a sharded tree this process drives + their transport (imt_comm.cu)

struct imt_ctx {
    int device = 0;
    int fmt = 0;                         // imt_fe_format
    cudaStream_t stream = nullptr;       // compute (own_stream unless the caller supplied one)
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // host<->device staging, overlapped with compute
    cudaStream_t aux_stream = nullptr;   // second compute stream: consecutive chunk kernels of a pipelined call alternate between
                                         // `stream` and this one so that chunk k+1 fills the SMs while chunk k drains
    cudaMemPool_t pool = nullptr;        // this context's OWN stream-ordered pool: every tree and scratch buffer comes from it (its release
                                         // threshold is raised so that calls recycle memory; the device's default pool is left alone)
    uint32_t* d_err = nullptr;           // device error bits, see kErr*
    uint32_t* h_err = nullptr;           // pinned mirror
    imt::PoseidonParams* d_params = nullptr;  // global-memory copy of the parameters (lane-dependent reads of the cooperative kernel)
    void* d_lh_aux = nullptr;            // imt::LhAux: per-round tables of the lead / helper latency kernel (poseidon_lh.cuh), made at context creation
    int sm_count = 0;                    // one block of that kernel per SM at most
    // The Poseidon instance of this context. imt_ctx_create: <3, 2>(8, 57) on the tuned kernels (generic == false; the
    // any-width kernels then only serve input lengths other than 2 and 3). imt_ctx_create_spec: every hash of the context
    // runs on the any-width kernels of imt_spec.cu (generic == true).
    imt::SpecLayout spec{3, 8, 57};
    bool generic = false;
    imt::Fr* d_spec = nullptr;  // SpecLayout-ordered parameter array (derived and uploaded on first use)
    imt::Fr* d_zero_leaf = nullptr;  // H3(0, 0, 0) in Montgomery form: the empty leaf (indexed_merkle_tree.rs:247-251); made on first use
    // multi-GPU: the group this context belongs to (imt_comm_create / imt_multi_create) and its position among the group's
    // local contexts; null for a single-GPU context
    imt_group* group = nullptr;
    unsigned group_slot = 0;
    uint64_t launches = 0;
    std::string last_error;
    // optional per-launch device timing of the hash kernels
    bool timing = false;
    struct Timed {
        cudaEvent_t a, b;
        int arity;
        size_t hashes;
    };
    std::vector<Timed> pending;
    double kernel_ms[4] = {0, 0, 0, 0};
    uint64_t kernel_launches[4] = {0, 0, 0, 0};
    uint64_t kernel_hashes[4] = {0, 0, 0, 0};
};

struct imt_tree {
    imt_ctx* ctx = nullptr;
    size_t n = 0;                 // leaves on this rank
    unsigned depth = 0;           // log2(n)
    imt::Fr* d_levels = nullptr;  // 2n - 1 FE, Montgomery, level 0 first
    imt::Fr* d_pre = nullptr;     // 3n FE in the context format (only when built from leaves)
    bool owns_pre = false;
    // subtree sharding
    unsigned rank = 0, world = 1, cap_depth = 0;
    imt::Fr* d_cap = nullptr;  // 2*world - 1 FE, Montgomery
    unsigned cap_alloc_world = 0;
    bool cap_valid = false;  // a rebuild makes the attached cap stale until the roots are exchanged again
    // sorted index over the occupied leaves (low-leaf lookups); built lazily, kept current by imt_insert_batch
    bool index_valid = false;
    size_t occupied = 0;                 // occupied slots form the prefix [0, occupied)
    size_t index_capacity = 0;           // entries the two arrays below can hold
    imt::Fr* d_sorted_keys = nullptr;    // canonical integer values, ascending
    uint32_t* d_sorted_slots = nullptr;  // slot of each key
    imt::Fr* d_alt_keys = nullptr;       // merge target of the next insert batch (allocated by the first one; the two
    uint32_t* d_alt_slots = nullptr;     // buffer pairs swap roles after every merge)
    bool head_next_zero = false;         // preimage[0].next_val == 0  (the reference's first-insert branch, IMT:640)
    // search accelerator of the lookups (rebuilt lazily after the sorted keys change): the top 64 bits of every sorted key
    // (8 B per key: four per 32-byte sector) and kIndexTop evenly spaced samples of those (staged in shared memory)
    unsigned long long* d_prefix = nullptr;
    unsigned long long* d_top = nullptr;
    size_t prefix_capacity = 0, top_stride = 0;
    bool prefix_valid = false;
};

namespace imt {

constexpr int kFmtCanonical = 0;
constexpr int kFmtMontgomery = 1;
#ifndef IMT_HASH_THREADS
#define IMT_HASH_THREADS 128  // swept 64 / 128 / 256
#endif
constexpr int kHashThreads = IMT_HASH_THREADS;

// bits of imt_ctx::d_err
constexpr uint32_t kErrNonCanonical = 1u;   // an input FE >= p
constexpr uint32_t kErrIndexOob = 2u;       // a leaf index outside the tree
constexpr uint32_t kErrNotWellFormed = 4u;  // preimages are not a consistent indexed tree
constexpr uint32_t kErrBadInsert = 8u;      // insert value is 0, already present, or repeated inside the batch
constexpr uint32_t kErrBadFold = 16u;       // imt_insert_witness::fold_nodes are not the chain values of the witnesses they came with

// FE offset of level `lvl` inside the concatenated level buffer of an n-leaf tree (n a power of two)
__host__ __device__ __forceinline__ size_t level_offset(size_t n, unsigned lvl) { return 2 * n - 2 * (n >> lvl); }

}  // namespace imt

#define IMT_TRY_CUDA(ctx, expr)                                                     \
    do {                                                                            \
        cudaError_t e_ = (expr);                                                    \
        if (e_ != cudaSuccess) {                                                    \
            (ctx)->last_error = std::string(#expr) + ": " + cudaGetErrorString(e_); \
            return IMT_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

#define IMT_TRY(expr)                \
    do {                             \
        imt_status s_ = (expr);      \
        if (s_ != IMT_OK) return s_; \
    } while (0)

namespace imt_host {

inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

inline imt_status fail(imt_ctx* ctx, imt_status st, const char* what) {
    ctx->last_error = what;
    return st;
}

// RAII device scratch buffer, stream-ordered: cudaMallocFromPoolAsync / cudaFreeAsync on the context's compute stream from
// the context's own memory pool (imt_ctx_create raises ITS release threshold, so steady-state calls recycle memory
// instead of paying a synchronous cudaMalloc / cudaFree per buffer — those cost ~100 ms per insert batch at depth 24 —
// while other users of the device's default pool, e.g. torch's cudaMallocAsync backend, are not affected).
struct DevBuf {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    cudaMemPool_t pool = nullptr;
    explicit DevBuf(imt_ctx* ctx) : s(ctx->stream), pool(ctx->pool) {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() {
        if (p) cudaFreeAsync(p, s);
    }
    cudaError_t alloc(size_t bytes) { return cudaMallocFromPoolAsync(&p, bytes ? bytes : 1, pool, s); }
    template <class T>
    T* as() { return static_cast<T*>(p); }
};

// RAII CUDA event (destroyed on every return path)
struct Event {
    cudaEvent_t e = nullptr;
    Event() = default;
    Event(const Event&) = delete;
    Event& operator=(const Event&) = delete;
    ~Event() {
        if (e) cudaEventDestroy(e);
    }
    cudaError_t create(unsigned flags = cudaEventDisableTiming) { return cudaEventCreateWithFlags(&e, flags); }
    operator cudaEvent_t() const { return e; }
};

// The long-lived buffers of a tree (levels, preimages, cap, sorted index) come from the same pool: a caller that maps
// the reference's `IndexedMerkleTree::new` to build + destroy per call recycles memory instead of paying cudaMalloc /
// cudaFree of 2.5 GiB (133 ms per depth-24 tree, measured).
inline cudaError_t tree_malloc(imt_ctx* ctx, void** p, size_t bytes) {
    cudaError_t e = cudaMallocFromPoolAsync(p, bytes ? bytes : 1, ctx->pool, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // usable from the copy stream right away
    return e;
}
inline void tree_free(imt_ctx* ctx, void* p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}

// ---- implemented in imt_capi.cu
imt_status clear_err(imt_ctx* ctx);
// waits for the compute stream and turns the device error bits into a status
imt_status finish(imt_ctx* ctx);
// out[i] = H(in[arity*i ..]) on `s`; formats are imt::kFmt*
imt_status launch_hash(imt_ctx* ctx, int arity, const void* d_in, void* d_out, size_t n, int in_fmt, int out_fmt, cudaStream_t s);
// dense FE array between formats (validates < p)
imt_status launch_convert(imt_ctx* ctx, const void* d_in, void* d_out, size_t n, int from_fmt, int to_fmt);
// batched get_proof from device indices into device buffers (any of d_helpers / d_helpers_fe may be null)
// select = true (sharded calls): indices owned by other ranks yield zeros instead of IMT_ERR_INDEX_OOB
imt_status launch_gather_proofs(imt_tree* t, const uint64_t* d_idx, size_t q, void* d_siblings, uint8_t* d_helpers, void* d_helpers_fe,
                                bool select = false);

// witness traces of the paths of a resident tree, queued on the compute stream (q x depth independent traced hashes)
// lead_slots: hash slots left free in front of every query's path hashes in d_states (default instance only when non-zero)
imt_status launch_tree_trace(imt_tree* t, const uint64_t* d_idx, size_t q, void* d_states, void* d_sbox = nullptr, unsigned lead_slots = 0);

// one tree level, Montgomery in / out: d_dst[i] = H(d_src[2i], d_src[2i+1]); small levels take the cooperative kernel
imt_status launch_level(imt_ctx* ctx, const imt::Fr* d_src, imt::Fr* d_dst, size_t nodes);

// asynchronous halves of the build, for callers that drive several devices from one host thread:
// enqueue_rebuild queues the leaf hashing (with the H2D pipeline when the leaves are on the host) and every level on the
// context's streams and returns; wait_staging blocks until the caller's HOST buffer has been consumed; finish() ends the call.
imt_status enqueue_rebuild(imt_tree* t, const void* preimages, bool device_src);
imt_status wait_staging(imt_ctx* ctx);
imt_status tree_alloc(imt_ctx* ctx, size_t n, bool with_pre, imt_tree** out);
imt_status check_leaf_count(imt_ctx* ctx, size_t n);

// ---- implemented in imt_latency.cu (compiled with free carry chains): the kernels of small batches
cudaError_t latency_upload_params(const imt::PoseidonParams* host_params);
// 3 lanes per hash (poseidon_coop.cuh): batches of <= coop_max_nodes() hashes
// `concurrent`: launches of this size that run side by side on other streams (the two half-trees of a build). While all hashes in flight
// fit one block per SM the batch goes to the lead / helper kernel (poseidon_lh.cuh), otherwise to the 3-lanes-per-hash kernel.
void launch_hash_coop(imt_ctx* ctx, int arity, const void* d_in, void* d_out, size_t n, int in_fmt, int out_fmt, cudaStream_t s,
                      unsigned concurrent = 1);
// force one of the two latency kernels (self-test, A/B measurements): which = 0 -> 3 lanes per hash, 1 -> lead / helper
void launch_hash_latency(imt_ctx* ctx, int which, int arity, const void* d_in, void* d_out, size_t n, int in_fmt, int out_fmt, cudaStream_t s);
// per-context tables of the latency kernels (after the parameters are on the device); freed by latency_teardown
cudaError_t latency_setup(imt_ctx* ctx);
void latency_teardown(imt_ctx* ctx);
void launch_fold_coop(imt_ctx* ctx, const void* d_leaves, const uint64_t* d_indices, const void* d_roots, const void* d_siblings, size_t q,
                      unsigned depth, uint8_t* d_ok, void* d_roots_out, void* d_states);

// ---- implemented in imt_indexed.cu
// 128-bit limb witnesses of (low leaf, new value) pairs on the compute stream; new values are read with a stride in FE
imt_status launch_limb_witness(imt_ctx* ctx, const void* d_low_leaves, const void* d_new_vals, size_t new_stride, size_t b, void* d_limbs,
                               uint8_t* d_flags);
void invalidate_index(imt_tree* t);
imt_status ensure_index(imt_tree* t);
// lookup of q device values + everything verify_non_inclusion loads about their low leaves, queued on the compute stream (no wait)
imt_status queue_non_inclusion(imt_tree* t, const void* d_values, size_t q, uint64_t* d_low, uint8_t* d_matched, void* d_low_leaves,
                               uint8_t* d_is_largest, void* d_siblings, uint8_t* d_helpers);

// ---- implemented in imt_comm.cu: collectives over the ranks of a group. Element i of every vector belongs to the group's
// i-th LOCAL context (one entry in process-per-GPU mode, all ranks in single-process mode); every operation is queued on
// that context's compute stream (stream-ordered after the kernels that produced the send buffers).
// all_gather: recv[i] = concatenation over ranks r of that rank's `bytes` bytes (rank-major), on every local rank
imt_status group_all_gather(imt_group* g, const std::vector<const void*>& send, const std::vector<void*>& recv, size_t bytes);
// in-place sum over ranks of `count` uint64 / uint8 values (exactly one rank holds a non-zero value per element: a select)
imt_status group_all_reduce_sum(imt_group* g, const std::vector<void*>& buf, size_t count, bool bytes8);
unsigned group_world(const imt_group* g);
unsigned group_local_count(const imt_group* g);
imt_ctx* group_ctx(imt_group* g, unsigned slot);
unsigned group_rank(const imt_group* g, unsigned slot);
// the group and the local shards behind a single-process multi-GPU tree
imt_status mtree_parts(imt_mtree* mt, imt_group** g, std::vector<imt_tree*>* trees);
imt_status mtree_alloc(imt_multi* m, size_t n, imt_mtree** out);  // buffers only, nothing built
imt_status mtree_rebuild_resident(imt_mtree* mt);                // re-hash the preimages resident in every shard + exchange

// ---- implemented in imt_spec.cu (any-width instances)
// permutations of one hash of `arity` inputs, and FE per hash of its witness trace
inline size_t spec_perms(const imt_ctx* ctx, size_t arity) { return arity / (ctx->spec.t - 1) + 1; }
inline size_t trace_fe_per_hash(const imt_ctx* ctx, size_t arity) {
    return spec_perms(ctx, arity) * ctx->spec.states_per_perm() * ctx->spec.t;
}
// FE per hash of the extended S-box trace: (x^2, x^4, x^5 + c) for each of the r_f t + r_p S-boxes of every permutation
inline size_t sbox_fe_per_hash(const imt_ctx* ctx, size_t arity) {
    return spec_perms(ctx, arity) * ((size_t)ctx->spec.r_f * ctx->spec.t + ctx->spec.r_p) * 3;
}
// derives + uploads ctx->d_spec if it is not there yet
imt_status ensure_spec(imt_ctx* ctx);
// out[i] = squeeze(update(in[arity*i ..])) with the context's instance; d_states may be null
imt_status launch_spec_hash(imt_ctx* ctx, size_t arity, const void* d_in, void* d_out, size_t n, int in_fmt, int out_fmt, void* d_states,
                            cudaStream_t s, void* d_sbox = nullptr);
// witness traces of the paths of a resident tree, one thread per (query, level), with the context's instance
imt_status launch_spec_tree_trace(imt_tree* t, const uint64_t* d_idx, size_t q, void* d_states, void* d_sbox = nullptr);
// batched verify_proof / compute_merkle_root (+ trace) with the context's instance
imt_status launch_spec_fold(imt_ctx* ctx, const void* d_leaves, const uint64_t* d_indices, const void* d_roots, const void* d_siblings, size_t q,
                            unsigned depth, uint8_t* d_ok, void* d_roots_out, void* d_states);

}  // namespace imt_host

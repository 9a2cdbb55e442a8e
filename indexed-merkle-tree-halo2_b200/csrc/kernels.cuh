// sm_100a kernels for the Poseidon / indexed-Merkle-tree hot path. One thread = one hash (two permutations,
// ~138k IMAD-pipe instructions); the integer-multiply pipe is the bound, HBM is ~500x away (96-128 B per hash),
// so the memory side only needs vectorised 128-bit accesses. See DESIGN.md for the per-kernel rooflines.
#pragma once
#include <cuda_runtime.h>

#include "kernels_common.cuh"

namespace imt {

// out[i] = H(in[ARITY*i .. ARITY*i + ARITY)). in_fmt/out_fmt select canonical <-> Montgomery conversion at the
// edges; tree levels are Montgomery on both sides.
// Occupancy hint of the throughput kernels: 7 (node levels) / 6 (leaf hashing) blocks of 128 threads per SM = 72 / 80 registers per
// thread instead of the 84 / 92 ptxas picks on its own. Round 1 (serial carry chains) swept 1 / 6 / 7 / 8 on the depth-24 build: 7 was
// best for both. With carry mask 22 (round 2) the leaf kernel prefers 6: 2^22 leaves 66.7 ms at 7, 64.5 ms at 6, 65.9 ms at 5; the node
// kernel stays at 7 (tools/lab/latency_lab.cu, profiles/r02_latency_lab.md).
#ifndef IMT_HASH_MIN_BLOCKS
#define IMT_HASH_MIN_BLOCKS 7
#endif
#ifndef IMT_HASH_MIN_BLOCKS_LEAF
#define IMT_HASH_MIN_BLOCKS_LEAF 6
#endif
template <int ARITY>
__global__ void __launch_bounds__(kHashThreads, ARITY == 3 ? IMT_HASH_MIN_BLOCKS_LEAF : IMT_HASH_MIN_BLOCKS)
    k_hash(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n, int in_fmt, int out_fmt, uint32_t* __restrict__ err) {
    const size_t i = blockIdx.x * (size_t)kHashThreads + threadIdx.x;
    if (i >= n) return;
    uint32_t x[ARITY][8], d[8];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < ARITY; ++j) {
        load_fe(x[j], in + 2 * (ARITY * i + j));
        ok &= ingest(x[j], in_fmt);
    }
    if (!ok) atomicOr(err, kErrNonCanonical);
    NoTrace nt;
    hash_fixed<ARITY>(d, x, c_params, nt);
    egress(d, out_fmt);
    store_fe(out + 2 * i, d);
}

// Writes every traced state as 3 FE in the user format. One thread owns one hash: 132 x 96 contiguous bytes.
// `sbox` (may be null) receives the extended trace: (x^2, x^4, x^5 + c) of every S-box in execution order —
// kSboxPerHash x 3 FE per hash, contiguous.
struct TraceSink {
    uint4* dst;
    int fmt;
    uint4* sbox = nullptr;
    __device__ __forceinline__ void put(uint4*& p, const uint32_t* x) {
        uint32_t t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = x[i];
        if (fmt == kFmtCanonical) from_mont(t, t);
        else canonicalize(t);
        store_fe(p, t);
        p += 2;
    }
    __device__ __forceinline__ void emit_sbox(const uint32_t* x2, const uint32_t* x4, const uint32_t* u) {
        if (sbox) {
            put(sbox, x2);
            put(sbox, x4);
            put(sbox, u);
        }
    }
    __device__ __forceinline__ void emit(const uint32_t (*s)[8]) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            uint32_t t[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = s[j][i];
            if (fmt == kFmtCanonical) from_mont(t, t);
            else canonicalize(t);
            store_fe(dst, t);
            dst += 2;
        }
    }
};

// hash i writes its states at hash slot i * slot_stride (1: dense; 1 + depth: the leaf hash in front of the path hashes of one query,
// imt_non_inclusion_witness_trace); the S-box trace and the digests are always dense
template <int ARITY>
__global__ void __launch_bounds__(kHashThreads) k_trace_hash(const uint4* __restrict__ in, uint4* __restrict__ states,
                                                             uint4* __restrict__ digests, size_t n, int fmt,
                                                             uint32_t* __restrict__ err, uint4* __restrict__ sbox, size_t slot_stride = 1) {
    const size_t i = blockIdx.x * (size_t)kHashThreads + threadIdx.x;
    if (i >= n) return;
    uint32_t x[ARITY][8], d[8];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < ARITY; ++j) {
        load_fe(x[j], in + 2 * (ARITY * i + j));
        ok &= ingest(x[j], fmt);
    }
    if (!ok) atomicOr(err, kErrNonCanonical);
    if (states) {
        TraceSink sink{states + i * slot_stride * (size_t)(kStatesPerHash * 3 * 2), fmt, sbox ? sbox + i * (size_t)(kSboxPerHash * 3 * 2) : nullptr};
        hash_fixed<ARITY>(d, x, c_params, sink);
    } else {
        NoTrace nt;
        hash_fixed<ARITY>(d, x, c_params, nt);
    }
    egress(d, fmt);
    if (digests) store_fe(digests + 2 * i, d);
}

// Batched get_proof: one thread per (query, level). For a sharded tree the top `cap_depth` levels come from the
// replicated cap (built over the gathered subtree roots) and `rank` locates this subtree inside it.
// n_total_select != 0 ("select" mode of the sharded calls): an index owned by ANOTHER rank (still < n_total_select) is not an
// error — this rank writes zeros for it, and the sum over ranks (one owner per query) assembles the replicated result.
__global__ void k_gather_proofs(const uint4* __restrict__ levels, const uint4* __restrict__ cap, size_t n_local,
                                unsigned depth_local, unsigned cap_depth, unsigned rank, const uint64_t* __restrict__ idx,
                                size_t q, int fmt, uint4* __restrict__ siblings, uint8_t* __restrict__ helpers,
                                uint4* __restrict__ helpers_fe, uint32_t* __restrict__ err, uint64_t n_total_select) {
    const unsigned depth = depth_local + cap_depth;
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= q * depth) return;
    const size_t qi = t / depth;
    const unsigned lvl = (unsigned)(t % depth);
    const uint64_t g = idx[qi];
    const uint64_t base = (uint64_t)rank * n_local;
    if (g < base || g >= base + n_local) {
        if (n_total_select && g < n_total_select) {
            const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            store_fe(siblings + 2 * t, z);
            if (helpers) helpers[t] = 0;
            if (helpers_fe) store_fe(helpers_fe + 2 * t, z);
            return;
        }
        atomicOr(err, kErrIndexOob);
        return;
    }
    const uint64_t local = g - base;
    uint32_t x[8];
    unsigned left;
    if (lvl < depth_local) {
        const uint64_t node = local >> lvl;
        left = (node & 1) == 0;
        load_fe(x, levels + 2 * (level_offset(n_local, lvl) + (node ^ 1)));
    } else {
        const unsigned cl = lvl - depth_local;
        const uint64_t node = (uint64_t)rank >> cl;
        left = (node & 1) == 0;
        load_fe(x, cap + 2 * (level_offset((size_t)1 << cap_depth, cl) + (node ^ 1)));
    }
    egress(x, fmt);
    store_fe(siblings + 2 * t, x);
    if (helpers) helpers[t] = (uint8_t)left;
    if (helpers_fe) {
        uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (left) {
            if (fmt == kFmtCanonical) h[0] = 1;
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) h[i] = c_params.one.l[i];
            }
        }
        store_fe(helpers_fe + 2 * t, h);
    }
}

// Batched verify_proof / compute_merkle_root: one thread folds one path. With `states` it also writes the
// witness trace of every hash of the fold.
__global__ void __launch_bounds__(kHashThreads) k_fold_paths(const uint4* __restrict__ leaves, const uint64_t* __restrict__ idx,
                                                             const uint4* __restrict__ siblings, const uint4* __restrict__ roots,
                                                             size_t q, unsigned depth, int fmt, uint8_t* __restrict__ ok_out,
                                                             uint4* __restrict__ roots_out, uint4* __restrict__ states,
                                                             uint32_t* __restrict__ err) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= q) return;
    uint32_t x[2][8], h[8];
    load_fe(h, leaves + 2 * i);
    bool ok = ingest(h, fmt);
    uint64_t index = idx[i];
#pragma unroll 1
    for (unsigned l = 0; l < depth; ++l) {
        uint32_t s[8];
        load_fe(s, siblings + 2 * (i * depth + l));
        ok &= ingest(s, fmt);
        const bool left = (index & 1) == 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x[0][k] = left ? h[k] : s[k];
            x[1][k] = left ? s[k] : h[k];
        }
        if (states) {
            TraceSink sink{states + (i * depth + l) * (size_t)(kStatesPerHash * 3 * 2), fmt};
            hash_fixed<2>(h, x, c_params, sink);
        } else {
            NoTrace nt;
            hash_fixed<2>(h, x, c_params, nt);
        }
        index >>= 1;
    }
    canonicalize(h);  // depth 0: the leaf itself, possibly semi-reduced after ingest
    if (ok_out) {
        uint32_t r[8];
        load_fe(r, roots + 2 * i);
        ok &= ingest(r, fmt);  // a root encoded as root + p must be rejected, not folded onto its canonical twin
        canonicalize(r);
        bool same = true;
#pragma unroll
        for (int k = 0; k < 8; ++k) same &= r[k] == h[k];
        ok_out[i] = (uint8_t)same;
    }
    if (!ok) atomicOr(err, kErrNonCanonical);
    if (roots_out) {
        egress(h, fmt);
        store_fe(roots_out + 2 * i, h);
    }
}

// (no occupancy hint here: with 7 / 6 blocks per SM the trace sink spills and the kernel drops from 60.6 to 58.1 / 60.0 M hashes/s)
// Witness traces of verify_merkle_proof (indexed_merkle_tree.rs:65-96) for paths of a tree that is resident on the device:
// hash l of the fold for leaf index g is H(level[l][2k], level[l][2k+1]) with k = g >> (l + 1), and both operands are already
// stored — so the q x depth hashes are INDEPENDENT. One thread per (query, level): a 2^16-path batch runs 1.5 M threads at
// full occupancy instead of 65 536 serial folds. Same bytes out as k_fold_paths with the state sink.
__global__ void __launch_bounds__(kHashThreads) k_trace_tree_paths(const uint4* __restrict__ levels, const uint4* __restrict__ cap,
                                                                   size_t n_local, unsigned depth_local, unsigned cap_depth, unsigned rank,
                                                                   const uint64_t* __restrict__ idx, size_t q, int fmt,
                                                                   uint4* __restrict__ states, uint32_t* __restrict__ err,
                                                                   uint4* __restrict__ sbox, unsigned lead_slots = 0) {
    // lead_slots: hash slots left free in front of every query's path hashes (1: the traced leaf hash of verify_non_inclusion)
    const unsigned depth = depth_local + cap_depth;
    const size_t t = blockIdx.x * (size_t)kHashThreads + threadIdx.x;
    if (t >= q * depth) return;
    const size_t qi = t / depth;
    const unsigned lvl = (unsigned)(t % depth);
    const uint64_t g = idx[qi];
    const uint64_t base = (uint64_t)rank * n_local;
    if (g < base || g >= base + n_local) {
        atomicOr(err, kErrIndexOob);
        return;
    }
    const uint4* src;
    if (lvl < depth_local) src = levels + 2 * (level_offset(n_local, lvl) + (((g - base) >> lvl) & ~(uint64_t)1));
    else src = cap + 2 * (level_offset((size_t)1 << cap_depth, lvl - depth_local) + (((uint64_t)rank >> (lvl - depth_local)) & ~(uint64_t)1));
    uint32_t x[2][8], d[8];
    load_fe(x[0], src);      // tree levels are Montgomery, canonical
    load_fe(x[1], src + 2);
    TraceSink sink{states + (t + (qi + 1) * lead_slots) * (size_t)(kStatesPerHash * 3 * 2), fmt, sbox ? sbox + t * (size_t)(kSboxPerHash * 3 * 2) : nullptr};
    hash_fixed<2>(d, x, c_params, sink);
}

// ---- witness trace of the chip's insert_leaf (indexed_merkle_tree.rs:253-313), one insert = 3 + 4 depth hashes in call order:
//   slot 0            H3(low leaf before)                       verify_non_inclusion, IMT:193-194
//   1 .. d            fold of that hash up the low path          -> old root, IMT:196-204
//   d + 1             H3(low.val, new.val, new_idx)              the rewired low leaf, IMT:265-275
//   d + 2 .. 2d + 1   its fold up the SAME siblings              -> interim root, IMT:277-284
//   2d + 2 .. 3d + 1  fold of the empty leaf H3(0,0,0) (a constant in the chip, IMT:247-251) up the new leaf's path -> interim root, IMT:286-294
//   3d + 2            H3(new leaf)                               IMT:299-303
//   3d + 3 .. 4d + 2  its fold up the same path                  -> new root, IMT:305-313
// The four folds of an insert are independent of each other and of every other insert, so a batch of b inserts is one launch
// of 3b traced leaf hashes and then, level by level, one launch of 4b traced node hashes (digests carried in `dig`, Montgomery).
__device__ __forceinline__ bool fe_words_equal(const uint32_t* x, const uint4* p) {
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    return x[0] == a.x && x[1] == a.y && x[2] == a.z && x[3] == a.w && x[4] == b.x && x[5] == b.y && x[6] == b.z && x[7] == b.w;
}
__device__ __forceinline__ unsigned insert_trace_slot(unsigned fold, unsigned level, unsigned depth) {
    return fold == 0 ? 1 + level : fold == 1 ? depth + 2 + level : fold == 2 ? 2 * depth + 2 + level : 3 * depth + 3 + level;
}
__global__ void __launch_bounds__(kHashThreads) k_trace_insert_leaves(const uint4* __restrict__ low_leaves, const uint4* __restrict__ new_leaves,
                                                                      uint64_t first_idx, size_t b, unsigned depth, int fmt,
                                                                      const uint4* __restrict__ zero_leaf_hash, uint4* __restrict__ states,
                                                                      uint4* __restrict__ dig, uint4* __restrict__ new_low_out,
                                                                      uint32_t* __restrict__ err, const uint4* __restrict__ fold_nodes = nullptr) {
    const size_t i = blockIdx.x * (size_t)kHashThreads + threadIdx.x;
    if (i >= 3 * b) return;
    const size_t k = i / 3;
    const unsigned j = (unsigned)(i % 3);
    uint32_t x[3][8], d[8];
    bool ok = true;
    if (j == 1) {  // { low.val, new.val, new_idx }: the low leaf after it has been pointed at the new one (IMT:265-270)
        load_fe(x[0], low_leaves + 2 * (3 * k));
        load_fe(x[1], new_leaves + 2 * (3 * k));
        ok &= ingest(x[0], fmt);
        ok &= ingest(x[1], fmt);
        const uint64_t slot = first_idx + k;
        uint32_t c[8] = {(uint32_t)slot, (uint32_t)(slot >> 32), 0, 0, 0, 0, 0, 0};
        if (new_low_out) {  // the preimage itself, in the caller's format
            uint32_t o[3][8];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
#pragma unroll
                for (int t = 0; t < 8; ++t) o[q][t] = x[q][t];
                canonicalize(o[q]);
                egress(o[q], fmt);
                store_fe(new_low_out + 2 * (3 * k + q), o[q]);
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) o[2][t] = c[t];
            if (fmt == kFmtMontgomery) {
                to_mont(o[2], o[2]);
                canonicalize(o[2]);
            }
            store_fe(new_low_out + 2 * (3 * k + 2), o[2]);
        }
        to_mont(x[2], c);
    } else {
        const uint4* src = (j == 0 ? low_leaves : new_leaves) + 2 * (3 * k);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            load_fe(x[q], src + 2 * q);
            ok &= ingest(x[q], fmt);
        }
    }
    if (!ok) atomicOr(err, kErrNonCanonical);
    const unsigned S = 3 + 4 * depth, slot = j == 0 ? 0 : j == 1 ? depth + 1 : 3 * depth + 2;
    if (states) {
        TraceSink sink{states + (k * S + slot) * (size_t)(kStatesPerHash * 3 * 2), fmt};
        hash_fixed<3>(d, x, c_params, sink);
    } else {
        NoTrace nt;
        hash_fixed<3>(d, x, c_params, nt);
    }
    if (fold_nodes) {  // one-launch form: the chain values are given; level 0 of each fold must be the leaf hash computed here
        if (depth == 0) return;
        const uint4* fn = fold_nodes + 2 * (4 * k * (size_t)depth);
        egress(d, fmt);
        if (!fe_words_equal(d, fn + 2 * ((j == 2 ? 3 : j) * (size_t)depth))) atomicOr(err, kErrBadFold);
        if (j == 0) {  // fold 2 starts from the empty leaf, a constant in the chip (IMT:247-251)
            uint32_t z[8];
            load_fe(z, zero_leaf_hash);
            egress(z, fmt);
            if (!fe_words_equal(z, fn + 2 * (2 * (size_t)depth))) atomicOr(err, kErrBadFold);
        }
        return;
    }
    store_fe(dig + 2 * (4 * k + (j == 2 ? 3 : j)), d);
    if (j == 0) {  // the empty leaf the new one replaces: fold 2 starts from the constant
        uint32_t z[8];
        load_fe(z, zero_leaf_hash);
        store_fe(dig + 2 * (4 * k + 2), z);
    }
}
// One-launch form of the level loop below: with the chain values of the four folds known (imt_insert_witness::fold_nodes, a by-product
// of imt_insert_batch) every one of the 4 b depth node hashes has both operands up front — one independent traced hash per thread, the
// same shape as k_trace_tree_paths. Each digest is checked against the next chain value (the top one is the fold's root and goes to
// roots_out), so a fold_nodes array that does not belong to these witnesses is reported (kErrBadFold), never traced silently.
__global__ void __launch_bounds__(kHashThreads) k_trace_insert_folds(const uint4* __restrict__ low_sib, const uint4* __restrict__ new_sib,
                                                                     const uint4* __restrict__ fold_nodes, const uint64_t* __restrict__ low_idx,
                                                                     uint64_t first_idx, size_t b, unsigned depth, int fmt,
                                                                     uint4* __restrict__ states, uint4* __restrict__ roots_out,
                                                                     uint32_t* __restrict__ err) {
    const size_t i = blockIdx.x * (size_t)kHashThreads + threadIdx.x;
    if (i >= 4 * b * depth) return;
    const size_t k = i / (4 * (size_t)depth);
    const unsigned r = (unsigned)(i - k * 4 * depth), f = r / depth, level = r - f * depth;
    uint32_t h[8], s[8], x[2][8], d[8];
    const uint4* node = fold_nodes + 2 * i;  // [k][f][level]
    load_fe(h, node);
    load_fe(s, (f < 2 ? low_sib : new_sib) + 2 * (k * depth + level));
    bool ok = ingest(h, fmt);
    ok &= ingest(s, fmt);
    if (!ok) atomicOr(err, kErrNonCanonical);
    const bool left = (((f < 2 ? low_idx[k] : first_idx + k) >> level) & 1) == 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        x[0][t] = left ? h[t] : s[t];
        x[1][t] = left ? s[t] : h[t];
    }
    if (states) {
        TraceSink sink{states + (k * (3 + 4 * depth) + insert_trace_slot(f, level, depth)) * (size_t)(kStatesPerHash * 3 * 2), fmt};
        hash_fixed<2>(d, x, c_params, sink);
    } else {
        NoTrace nt;
        hash_fixed<2>(d, x, c_params, nt);
    }
    egress(d, fmt);
    if (level + 1 == depth) {
        if (roots_out) store_fe(roots_out + 2 * (4 * k + f), d);
    } else if (!fe_words_equal(d, node + 2)) {
        atomicOr(err, kErrBadFold);
    }
}
__global__ void __launch_bounds__(kHashThreads) k_trace_insert_level(const uint4* __restrict__ low_sib, const uint4* __restrict__ new_sib,
                                                                     const uint64_t* __restrict__ low_idx, uint64_t first_idx, size_t b,
                                                                     unsigned depth, unsigned level, int fmt, uint4* __restrict__ states,
                                                                     uint4* __restrict__ dig, uint4* __restrict__ roots_out,
                                                                     uint32_t* __restrict__ err) {
    const size_t i = blockIdx.x * (size_t)kHashThreads + threadIdx.x;
    if (i >= 4 * b) return;
    const size_t k = i >> 2;
    const unsigned f = (unsigned)(i & 3);
    uint32_t h[8], s[8], x[2][8], d[8];
    load_fe(h, dig + 2 * i);
    load_fe(s, (f < 2 ? low_sib : new_sib) + 2 * (k * depth + level));
    if (!ingest(s, fmt)) atomicOr(err, kErrNonCanonical);
    const uint64_t node = (f < 2 ? low_idx[k] : first_idx + k) >> level;
    const bool left = (node & 1) == 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        x[0][t] = left ? h[t] : s[t];
        x[1][t] = left ? s[t] : h[t];
    }
    if (states) {
        TraceSink sink{states + (k * (3 + 4 * depth) + insert_trace_slot(f, level, depth)) * (size_t)(kStatesPerHash * 3 * 2), fmt};
        hash_fixed<2>(d, x, c_params, sink);
    } else {
        NoTrace nt;
        hash_fixed<2>(d, x, c_params, nt);
    }
    store_fe(dig + 2 * i, d);
    if (roots_out && level + 1 == depth) {
        egress(d, fmt);
        store_fe(roots_out + 2 * i, d);
    }
}

// Format conversion of a dense FE array (used for roots / levels / preimages crossing the boundary)
__global__ void k_convert(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n, int from_fmt, int to_fmt,
                          uint32_t* __restrict__ err) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x[8];
    load_fe(x, in + 2 * i);
    if (!is_canonical(x)) atomicOr(err, kErrNonCanonical);
    if (from_fmt != to_fmt) {
        if (to_fmt == kFmtMontgomery) {
            to_mont(x, x);
            canonicalize(x);
        } else {
            from_mont(x, x);
        }
    }
    store_fe(out + 2 * i, x);
}

// Integer-multiply roofline probe: the exact instruction form the field arithmetic issues — IMAD.WIDE.U32.X, each
// one accumulating a 32x32->64 product into a 64-bit register pair with the carry chained through the flag — as one
// long dependent chain per thread. With enough resident warps this saturates the multiply pipe: it measures the
// sustained wide multiply-accumulates per second that ANY schedule of such instructions can reach on this chip.
__global__ void __launch_bounds__(256) k_imad_probe(uint64_t* __restrict__ out, uint32_t seed, int iters) {
    uint32_t lo[8], hi[8];
    const uint32_t b = seed * 2654435761u + threadIdx.x;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        lo[j] = seed + 977u * j + blockIdx.x;
        hi[j] = j;
    }
    cc::clear();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int j = 0; j < 8; ++j) cc::madwc_cc(lo[j], hi[j], lo[(j + 3) & 7], b);  // multiplier varies: nothing is loop invariant
        }
    }
    uint64_t x = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) x ^= ((uint64_t)hi[j] << 32) | lo[j];
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = x;
}

}  // namespace imt

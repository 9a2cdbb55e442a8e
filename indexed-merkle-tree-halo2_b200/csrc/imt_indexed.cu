// Indexed-leaf logic on the device: sorted-key index, low-leaf (predecessor) lookups, non-inclusion witnesses and
// batched inserts with per-insert roots.
//
// Replaces the O(n)-per-call test helpers of the reference:
//   update_idx_leaf                  /root/reference/src/indexed_merkle_tree.rs:632-660   (linear scan + list rewiring)
//   insert orchestration             /root/reference/src/indexed_merkle_tree.rs:710-741   (re-hash ALL leaves + rebuild per insert)
// with (a) a sorted array of the occupied leaves' canonical values + binary search per query, and (b) a
// level-synchronous versioned update: a batch of B inserts is 2B single-leaf writes at times t = 0..2B-1 (t = 2k:
// the low leaf of insert k, t = 2k+1: its new leaf); every (write, level) pair is one hash whose sibling operand is
// the latest earlier write to the sibling node, else the stored tree. All 2B hashes of a level run in one launch, so
// a batch costs `depth` launches instead of B full rebuilds, and yields the same roots and paths as the reference's
// sequence of rebuilds (checked against the oracle's restatement of that sequence).
#include <cub/device/device_merge_sort.cuh>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <new>

#include "imt_internal.h"
#include "kernels_common.cuh"

using namespace imt;
using namespace imt_host;

namespace {

// Inserts resolved together. Every level of a chunk is ONE launch of 2 x chunk hashes, so larger chunks amortise the hash
// latency of the `depth` dependent levels better (depth 24, 131072 inserts: 381 k inserts/s with chunks of 4096, 775 k with
// 16384, 858 k with 32768, see DESIGN.md for 65536); scratch is 25 x 2 x chunk field elements (105 MB at 65536).
// IMT_INSERT_CHUNK overrides it (tests, A/B).
constexpr size_t kInsertChunk = 65536;
constexpr size_t kShardInsertRound = 4096;  // inserts per round of the sharded insert calls (host-side plan arrays)

__device__ __forceinline__ int cmp256(const uint32_t* a, const uint32_t* b) {
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    }
    return 0;
}
__device__ __forceinline__ bool zero256(const uint32_t* a) { return (a[0] | a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7]) == 0; }
__device__ __forceinline__ void copy256(uint32_t* d, const uint32_t* s) {
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = s[i];
}
// context-format FE -> canonical integer (`Fr: Ord` compares canonical values, IMT:647). false when the input is >= p
__device__ __forceinline__ bool to_int(uint32_t* x, int fmt) {
    const bool ok = is_canonical(x);
    if (fmt == kFmtMontgomery) from_mont(x, x);
    return ok;
}
// canonical integer -> context-format FE
__device__ __forceinline__ void from_int(uint32_t* x, int fmt) {
    if (fmt == kFmtMontgomery) {
        to_mont(x, x);
        canonicalize(x);
    }
}
__device__ __forceinline__ void slot_to_int(uint32_t* x, uint64_t slot) {
    x[0] = (uint32_t)slot, x[1] = (uint32_t)(slot >> 32);
#pragma unroll
    for (int i = 2; i < 8; ++i) x[i] = 0;
}

struct KeyLess {
    __device__ __forceinline__ bool operator()(const Fr& a, const Fr& b) const { return cmp256(a.l, b.l) < 0; }
};

// first j in [0, m) with keys[j] >= v, else m
__device__ __forceinline__ size_t lower_bound(const uint4* __restrict__ keys, size_t m, const uint32_t* v) {
    size_t lo = 0, hi = m;
    while (lo < hi) {
        const size_t mid = lo + ((hi - lo) >> 1);
        uint32_t k[8];
        load_fe(k, keys + 2 * mid);
        if (cmp256(k, v) < 0) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------- index build
// keys[i] = canonical val of slot i, slots[i] = i; stats[0] = first empty slot (min), stats[1] = last occupied slot + 1
// (max; 0 = none: an all-empty shard). Global slot 0 is the head and always counts as occupied; any other slot is occupied iff val != 0.
// `base` = global slot of local slot 0 (rank * n for a subtree shard).
__global__ void __launch_bounds__(256) k_index_extract(const uint4* __restrict__ pre, size_t n, uint64_t base, int fmt, uint4* __restrict__ keys,
                                                       uint32_t* __restrict__ slots, unsigned long long* __restrict__ stats,
                                                       uint32_t* __restrict__ head_next_zero, uint32_t* __restrict__ err) {
    __shared__ unsigned long long s_min, s_max;
    if (threadIdx.x == 0) s_min = ~0ull, s_max = 0ull;
    __syncthreads();
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) {
        uint32_t v[8], a[8], b[8];
        load_fe(v, pre + 2 * (3 * i));
        load_fe(a, pre + 2 * (3 * i + 1));
        load_fe(b, pre + 2 * (3 * i + 2));
        if (!to_int(v, fmt)) atomicOr(err, kErrNonCanonical);
        store_fe(keys + 2 * i, v);
        slots[i] = (uint32_t)i;
        const bool occ = (base + i) == 0 || !zero256(v);
        if (occ) atomicMax(&s_max, (unsigned long long)i + 1);
        else {
            atomicMin(&s_min, (unsigned long long)i);
            if (!zero256(a) || !zero256(b)) atomicOr(err, kErrNotWellFormed);  // an empty slot is {0, 0, 0}
        }
        if (base + i == 0) *head_next_zero = zero256(a) ? 1u : 0u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_min != ~0ull) atomicMin(&stats[0], s_min);
        atomicMax(&stats[1], s_max);
    }
}

// the sorted order must be the linked list: next_val / next_idx of the j-th smallest point at the (j+1)-th smallest,
// the largest points at (0, 0), values are distinct
// (a shard of a larger tree only sees part of the list: there only distinctness is checked)
__global__ void __launch_bounds__(256) k_index_check(const uint4* __restrict__ pre, int fmt, const uint4* __restrict__ keys,
                                                     const uint32_t* __restrict__ slots, size_t m, bool links, uint32_t* __restrict__ err) {
    const size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (j >= m) return;
    if (!links) {
        if (j + 1 < m) {
            uint32_t k0[8], k1[8];
            load_fe(k0, keys + 2 * j);
            load_fe(k1, keys + 2 * (j + 1));
            if (cmp256(k0, k1) >= 0) atomicOr(err, kErrNotWellFormed);
        }
        return;
    }
    const size_t s = slots[j];
    uint32_t nv[8], ni[8];
    load_fe(nv, pre + 2 * (3 * s + 1));
    load_fe(ni, pre + 2 * (3 * s + 2));
    bool ok = to_int(nv, fmt);
    ok &= to_int(ni, fmt);
    if (!ok) atomicOr(err, kErrNonCanonical);
    bool good;
    if (j + 1 < m) {
        uint32_t k0[8], k1[8], want[8];
        load_fe(k0, keys + 2 * j);
        load_fe(k1, keys + 2 * (j + 1));
        slot_to_int(want, slots[j + 1]);
        good = cmp256(k0, k1) < 0 && cmp256(nv, k1) == 0 && cmp256(ni, want) == 0;
    } else {
        good = zero256(nv) && zero256(ni);
    }
    if (!good) atomicOr(err, kErrNotWellFormed);
}

// ------------------------------------------------------------------------------------------------- lookups
// The read-only half of update_idx_leaf (IMT:638-647) answered from the sorted index. In a well-formed tree the
// scan's predicate  val < v && (next_val > v || next_val == 0)  holds for exactly one occupied slot — the
// predecessor of v — unless v is 0 or already present; then the scan falls through to the first EMPTY slot
// (val 0 < v, next_val 0), or matches nothing.
__global__ void __launch_bounds__(256) k_low_leaf_lookup(const uint4* __restrict__ keys, const uint32_t* __restrict__ slots, size_t m,
                                                         size_t n, uint32_t head_next_zero, const uint4* __restrict__ values, size_t q,
                                                         int fmt, uint64_t* __restrict__ low_idx, uint8_t* __restrict__ matched,
                                                         uint32_t* __restrict__ err) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= q) return;
    uint32_t v[8];
    load_fe(v, values + 2 * i);
    if (!to_int(v, fmt)) atomicOr(err, kErrNonCanonical);
    uint64_t low = 0;
    uint8_t hit = 0;
    if (head_next_zero) {  // IMT:640: `node.next_val == 0 && i == 0` matches before anything is compared
        hit = 1;
    } else {
        const size_t j = lower_bound(keys, m, v);
        bool present = false;
        if (j < m) {
            uint32_t k[8];
            load_fe(k, keys + 2 * j);
            present = cmp256(k, v) == 0;
        }
        if (j >= 1 && !present) {
            low = slots[j - 1];
            hit = 1;
        } else if (!zero256(v) && m < n) {
            low = m;
            hit = 1;
        }
    }
    low_idx[i] = low;
    if (matched) matched[i] = hit;
}

// ---- the fast lookup. A binary search over m 32-byte keys is ~log2(m) DEPENDENT 32-byte sector reads per query (24 at depth
// 24, half of them from DRAM: the key array is 512 MB). Instead:
//   * prefix[i] = the top 64 bits of sorted key i: 8 bytes per key, four keys per sector, 128 MB at depth 24 (L2-sized);
//   * top[kIndexTop] = the last prefix of each of kIndexTop equal blocks of the prefix array, staged in SHARED memory: the first
//     12 levels of the search cost no global traffic at all;
//   * the remaining log2(m / 4096) levels run on the prefix array inside one block of it (the last two share a sector);
//   * the 32-byte keys are read only when a 64-bit prefix ties with the query's (a present value, or — 2^-40 per pair for
//     random keys — a coincidence): ties are resolved on the full keys, long runs of equal prefixes by the plain binary search.
// ~10 sector reads per query, mostly L2 hits, instead of ~26.
constexpr unsigned kIndexTop = 4096;
__device__ __forceinline__ unsigned long long key_prefix(const uint32_t* k) { return ((unsigned long long)k[7] << 32) | k[6]; }
__global__ void __launch_bounds__(256) k_index_prefix(const uint4* __restrict__ keys, size_t m, unsigned long long* __restrict__ prefix) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint4 hi = __ldg(keys + 2 * i + 1);
    prefix[i] = ((unsigned long long)hi.w << 32) | hi.z;
}
__global__ void __launch_bounds__(256) k_index_top(const unsigned long long* __restrict__ prefix, size_t m, size_t stride,
                                                   unsigned long long* __restrict__ top) {
    const size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (b >= kIndexTop) return;
    const size_t first = b * stride;
    top[b] = first < m ? prefix[(first + stride < m ? first + stride : m) - 1] : ~0ull;  // blocks past the end never match
}
// first j in [0, m) with keys[j] >= v, else m — through the shared-memory top and the prefix array
__device__ __forceinline__ size_t lower_bound_fast(const uint4* __restrict__ keys, const unsigned long long* __restrict__ prefix,
                                                   const unsigned long long* s_top, size_t stride, size_t m, const uint32_t* v) {
    const unsigned long long pv = key_prefix(v);
    unsigned lo = 0, hi = kIndexTop;  // first block whose last prefix is >= pv
    while (lo < hi) {
        const unsigned mid = (lo + hi) >> 1;
        if (s_top[mid] < pv) lo = mid + 1;
        else hi = mid;
    }
    if (lo == kIndexTop) return m;
    size_t a = (size_t)lo * stride, b = a + stride < m ? a + stride : m;
    if (a >= m) return m;
    while (a < b) {  // first index in the block with prefix >= pv (the block's last entry qualifies)
        const size_t mid = a + ((b - a) >> 1);
        if (__ldg(prefix + mid) < pv) a = mid + 1;
        else b = mid;
    }
    // ties on the 64-bit prefix: decide on the full keys
    for (int step = 0; step < 4; ++step) {
        if (a >= m || __ldg(prefix + a) != pv) return a;
        uint32_t k[8];
        load_fe(k, keys + 2 * a);
        if (cmp256(k, v) >= 0) return a;
        ++a;
    }
    return lower_bound(keys, m, v);  // a long run of equal prefixes (adversarial keys): the plain search is always right
}
__global__ void __launch_bounds__(512) k_low_leaf_lookup_fast(const uint4* __restrict__ keys, const uint32_t* __restrict__ slots,
                                                              const unsigned long long* __restrict__ prefix,
                                                              const unsigned long long* __restrict__ top, size_t stride, size_t m, size_t n,
                                                              uint32_t head_next_zero, const uint4* __restrict__ values, size_t q, int fmt,
                                                              uint64_t* __restrict__ low_idx, uint8_t* __restrict__ matched,
                                                              uint32_t* __restrict__ err) {
    __shared__ unsigned long long s_top[kIndexTop];
    for (unsigned i = threadIdx.x; i < kIndexTop; i += blockDim.x) s_top[i] = top[i];
    __syncthreads();
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= q) return;
    uint32_t v[8];
    load_fe(v, values + 2 * i);
    if (!to_int(v, fmt)) atomicOr(err, kErrNonCanonical);
    uint64_t low = 0;
    uint8_t hit = 0;
    if (head_next_zero) {  // IMT:640
        hit = 1;
    } else {
        const size_t j = lower_bound_fast(keys, prefix, s_top, stride, m, v);
        bool present = false;
        if (j < m && __ldg(prefix + j) == key_prefix(v)) {
            uint32_t k[8];
            load_fe(k, keys + 2 * j);
            present = cmp256(k, v) == 0;
        }
        if (j >= 1 && !present) {
            low = slots[j - 1];
            hit = 1;
        } else if (!zero256(v) && m < n) {
            low = m;
            hit = 1;
        }
    }
    low_idx[i] = low;
    if (matched) matched[i] = hit;
}

// Sharded lookup, per-rank half: the largest LOCAL key below v (as a canonical integer) with its GLOBAL slot.
// flags bit 0: a candidate exists, bit 1: v itself is a local key.
__global__ void __launch_bounds__(256) k_low_leaf_candidates(const uint4* __restrict__ keys, const uint32_t* __restrict__ slots, size_t m,
                                                             uint64_t base, const uint4* __restrict__ values, size_t q, int fmt,
                                                             uint4* __restrict__ cand_keys, uint64_t* __restrict__ cand_slots,
                                                             uint8_t* __restrict__ flags, uint32_t* __restrict__ err) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= q) return;
    uint32_t v[8], k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    load_fe(v, values + 2 * i);
    if (!to_int(v, fmt)) atomicOr(err, kErrNonCanonical);
    const size_t j = lower_bound(keys, m, v);
    uint8_t f = 0;
    uint64_t slot = 0;
    if (j < m) {
        uint32_t w[8];
        load_fe(w, keys + 2 * j);
        if (cmp256(w, v) == 0) f |= 2;
    }
    if (j >= 1) {
        load_fe(k, keys + 2 * (j - 1));
        slot = base + slots[j - 1];
        f |= 1;
    }
    store_fe(cand_keys + 2 * i, k);
    cand_slots[i] = slot;
    flags[i] = f;
}

// Sharded lookup, replicated half: the winner among the `world` gathered candidates of each query, then the same
// decision as k_low_leaf_lookup. occupied / n are the GLOBAL counts. v_zero[i] != 0 marks a zero query value.
__global__ void __launch_bounds__(256) k_low_leaf_merge(const uint4* __restrict__ cand_keys, const uint64_t* __restrict__ cand_slots,
                                                        const uint8_t* __restrict__ flags, unsigned world, size_t q, uint64_t occupied,
                                                        uint64_t n, uint32_t head_next_zero, const uint8_t* __restrict__ v_zero,
                                                        uint64_t* __restrict__ low_idx, uint8_t* __restrict__ matched) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= q) return;
    uint32_t best[8];
    bool have = false, present = false;
    uint64_t slot = 0;
    for (unsigned r = 0; r < world; ++r) {
        const uint8_t f = flags[(size_t)r * q + i];
        present |= (f & 2) != 0;
        if (f & 1) {
            uint32_t k[8];
            load_fe(k, cand_keys + 2 * ((size_t)r * q + i));
            if (!have || cmp256(k, best) > 0) copy256(best, k), slot = cand_slots[(size_t)r * q + i], have = true;
        }
    }
    uint64_t low = 0;
    uint8_t hit = 0;
    if (head_next_zero) hit = 1;
    else if (have && !present) low = slot, hit = 1;
    else if (!v_zero[i] && occupied < n) low = occupied, hit = 1;
    low_idx[i] = low;
    if (matched) matched[i] = hit;
}

__global__ void __launch_bounds__(256) k_is_zero(const uint4* __restrict__ values, size_t q, uint8_t* __restrict__ out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= q) return;
    uint32_t v[8];
    load_fe(v, values + 2 * i);
    out[i] = zero256(v) ? 1 : 0;
}

// n_total_select != 0: an index owned by another rank writes zeros instead of raising kErrIndexOob (see k_gather_proofs)
__global__ void __launch_bounds__(256) k_gather_leaves(const uint4* __restrict__ pre, size_t n, uint64_t base,
                                                       const uint64_t* __restrict__ idx, size_t q, uint4* __restrict__ leaves,
                                                       uint8_t* __restrict__ is_largest, uint32_t* __restrict__ err,
                                                       uint64_t n_total_select = 0) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= q) return;
    if (idx[i] < base || idx[i] - base >= n) {
        if (n_total_select && idx[i] < n_total_select) {
            if (leaves) {
#pragma unroll
                for (int k = 0; k < 6; ++k) leaves[6 * i + k] = make_uint4(0, 0, 0, 0);
            }
            if (is_largest) is_largest[i] = 0;
            return;
        }
        atomicOr(err, kErrIndexOob);
        return;
    }
    const size_t s = idx[i] - base;
    uint4 w[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) w[k] = __ldg(pre + 6 * s + k);
    if (leaves) {
#pragma unroll
        for (int k = 0; k < 6; ++k) leaves[6 * i + k] = w[k];
    }
    if (is_largest) is_largest[i] = ((w[2].x | w[2].y | w[2].z | w[2].w | w[3].x | w[3].y | w[3].z | w[3].w) == 0) ? 1 : 0;  // IMT:736-741
}

// The 128-bit limb witnesses verify_non_inclusion assigns (indexed_merkle_tree.rs:143-172, 206-222): for every (low leaf,
// new value) pair  nl = new value, ll = low.next_val, llv = low.val  each split as q * 2^128 + r, in the order the chip loads
// them: nl_q, nl_r, ll_q, ll_r, llv_q, llv_r; plus the three bits it derives from them (is_less_than, IMT:98-125):
// flags[3i] = nl < ll (is_next_val_greater, IMT:178), flags[3i+1] = llv < nl (check_less_than, IMT:226), flags[3i+2] = the
// witness satisfies both prover-side assertions (IMT:180-191 with is_new_leaf_largest = (ll == 0), IMT:226-228).
__global__ void __launch_bounds__(256) k_limb_witness(const uint4* __restrict__ low_leaves, const uint4* __restrict__ new_vals, size_t new_stride,
                                                      size_t b, int fmt, uint4* __restrict__ limbs, uint8_t* __restrict__ flags,
                                                      uint32_t* __restrict__ err) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= b) return;
    uint32_t v[3][8];  // nl, ll, llv as canonical integers
    load_fe(v[0], new_vals + 2 * new_stride * i);  // new_stride = 3: the val field of an array of leaf preimages
    load_fe(v[1], low_leaves + 2 * (3 * i + 1));
    load_fe(v[2], low_leaves + 2 * (3 * i));
    bool ok = true;
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
        ok &= is_canonical(v[k]);
        if (fmt == kFmtMontgomery) from_mont(v[k], v[k]);
    }
    if (!ok) atomicOr(err, kErrNonCanonical);
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
        uint32_t q[8] = {v[k][4], v[k][5], v[k][6], v[k][7], 0, 0, 0, 0};
        uint32_t r[8] = {v[k][0], v[k][1], v[k][2], v[k][3], 0, 0, 0, 0};
        if (fmt == kFmtMontgomery) {
            to_mont(q, q);
            canonicalize(q);
            to_mont(r, r);
            canonicalize(r);
        }
        store_fe(limbs + 2 * (6 * i + 2 * k), q);
        store_fe(limbs + 2 * (6 * i + 2 * k + 1), r);
    }
    if (flags) {
        const bool nl_lt_ll = cmp256(v[0], v[1]) < 0, llv_lt_nl = cmp256(v[2], v[0]) < 0;
        flags[3 * i] = nl_lt_ll;
        flags[3 * i + 1] = llv_lt_nl;
        flags[3 * i + 2] = (zero256(v[1]) || nl_lt_ll) && llv_lt_nl;
    }
}

// ------------------------------------------------------------------------------------------------- inserts
// whole-batch validation on the sorted batch: no zero, no repeats, none already in the tree
__global__ void __launch_bounds__(256) k_ins_validate(const uint4* __restrict__ sorted_vals, size_t b, const uint4* __restrict__ keys, size_t m,
                                                      uint32_t* __restrict__ err) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= b) return;
    uint32_t v[8];
    load_fe(v, sorted_vals + 2 * i);
    bool bad = zero256(v);
    if (i + 1 < b) {
        uint32_t w[8];
        load_fe(w, sorted_vals + 2 * (i + 1));
        bad |= cmp256(v, w) == 0;
    }
    const size_t j = lower_bound(keys, m, v);
    if (j < m) {
        uint32_t k[8];
        load_fe(k, keys + 2 * j);
        bad |= cmp256(k, v) == 0;
    }
    bad |= j == 0;  // nothing below v: the tree has no head {0, ..}
    if (bad) atomicOr(err, kErrBadInsert);
}

// Insert k of the chunk: its predecessor / successor among (indexed leaves) U (chunk values 0..k-1), i.e. the state
// the reference's scan sees at that point, and from them the four preimages IMT:648-654 reads and writes.
//   x[2k] = low slot, x[2k+1] = new slot;  upd[2k] = low leaf AFTER, upd[2k+1] = new leaf;  low_old[k] = low leaf BEFORE
// shared tail: refine (pred, succ) — the neighbours among the leaves indexed before the chunk — with the chunk's own
// earlier values, then emit the slots and preimages. Returns false on a repeat inside the chunk.
__device__ __forceinline__ bool ins_finish(size_t k, const uint32_t* v, uint32_t* pred, uint64_t pred_slot, uint32_t* succ, uint64_t succ_slot,
                                           bool has_succ, const uint4* __restrict__ vals, uint64_t first_idx, int fmt, uint64_t* __restrict__ x,
                                           uint4* __restrict__ upd, uint4* __restrict__ low_old, uint8_t* __restrict__ is_largest) {
    // A WARP per insert: lane j scans the earlier chunk values j, j + 32, ... (coalesced), then the 32 partial
    // (pred, succ) pairs are combined by a butterfly of 256-bit compares. One thread per insert made the last insert of a
    // 4096-chunk walk 4095 values alone on 32 resident blocks: 1.2 ms per chunk.
    const unsigned lane = threadIdx.x & 31;
    bool distinct = true;
    for (size_t i = lane; i < k; i += 32) {
        uint32_t w[8];
        load_fe(w, vals + 2 * i);
        const int c = cmp256(w, v);
        distinct &= c != 0;
        if (c < 0) {
            if (cmp256(w, pred) > 0) copy256(pred, w), pred_slot = first_idx + i;
        } else if (!has_succ || cmp256(w, succ) < 0) {
            copy256(succ, w), succ_slot = first_idx + i, has_succ = true;
        }
    }
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
        uint32_t op[8], os[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            op[i] = __shfl_xor_sync(0xffffffffu, pred[i], d);
            os[i] = __shfl_xor_sync(0xffffffffu, succ[i], d);
        }
        const uint64_t ops = __shfl_xor_sync(0xffffffffu, pred_slot, d), oss = __shfl_xor_sync(0xffffffffu, succ_slot, d);
        const bool ohs = __shfl_xor_sync(0xffffffffu, (int)has_succ, d) != 0;
        distinct &= __shfl_xor_sync(0xffffffffu, (int)distinct, d) != 0;
        // equal keys come from the same source (the index neighbours every lane started with): keep either
        if (cmp256(op, pred) > 0) copy256(pred, op), pred_slot = ops;
        if (ohs && (!has_succ || cmp256(os, succ) < 0)) copy256(succ, os), succ_slot = oss, has_succ = true;
    }
    if (lane != 0) return distinct;  // every lane holds the combined result; lane 0 writes it
    if (!has_succ) {
#pragma unroll
        for (int i = 0; i < 8; ++i) succ[i] = 0;
    }
    const uint64_t new_slot = first_idx + k;
    x[2 * k] = pred_slot;
    x[2 * k + 1] = new_slot;
    is_largest[k] = has_succ ? 0 : 1;
    uint32_t fv[8], fp[8], fs[8], fsi[8], fni[8];
    copy256(fv, v), from_int(fv, fmt);
    copy256(fp, pred), from_int(fp, fmt);
    copy256(fs, succ), from_int(fs, fmt);
    slot_to_int(fsi, succ_slot), from_int(fsi, fmt);
    slot_to_int(fni, new_slot), from_int(fni, fmt);
    uint4* lo = low_old + 6 * k;  // {pred, succ, succ_slot}
    store_fe(lo, fp), store_fe(lo + 2, fs), store_fe(lo + 4, fsi);
    uint4* u0 = upd + 6 * (2 * k);  // {pred, v, new_slot}
    store_fe(u0, fp), store_fe(u0 + 2, fv), store_fe(u0 + 4, fni);
    uint4* u1 = upd + 6 * (2 * k + 1);  // {v, succ, succ_slot}
    store_fe(u1, fv), store_fe(u1 + 2, fs), store_fe(u1 + 4, fsi);
    return distinct;
}

__global__ void __launch_bounds__(128) k_ins_resolve(const uint4* __restrict__ keys, const uint32_t* __restrict__ slots, size_t m,
                                                     const uint4* __restrict__ vals, size_t b, uint64_t first_idx, int fmt,
                                                     uint64_t* __restrict__ x, uint4* __restrict__ upd, uint4* __restrict__ low_old,
                                                     uint8_t* __restrict__ is_largest) {
    const size_t k = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;  // a warp per insert (see ins_finish)
    if (k >= b) return;
    uint32_t v[8], pred[8], succ[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    load_fe(v, vals + 2 * k);
    const size_t j = lower_bound(keys, m, v);  // validated: j >= 1 and keys[j] != v
    load_fe(pred, keys + 2 * (j - 1));
    uint64_t succ_slot = 0;
    const bool has_succ = j < m;
    if (has_succ) {
        load_fe(succ, keys + 2 * j);
        succ_slot = slots[j];
    }
    ins_finish(k, v, pred, slots[j - 1], succ, succ_slot, has_succ, vals, first_idx, fmt, x, upd, low_old, is_largest);
}

// Sharded inserts, per-rank half: the neighbours of every value in THIS rank's sorted index (keys as canonical
// integers, slots global). flags bit 0: predecessor exists, bit 1: the value is a local key, bit 2: successor exists.
__global__ void __launch_bounds__(256) k_ins_neighbors(const uint4* __restrict__ keys, const uint32_t* __restrict__ slots, size_t m, uint64_t base,
                                                       const uint4* __restrict__ values, size_t b, int fmt, uint4* __restrict__ pred_keys,
                                                       uint64_t* __restrict__ pred_slots, uint4* __restrict__ succ_keys,
                                                       uint64_t* __restrict__ succ_slots, uint8_t* __restrict__ flags, uint32_t* __restrict__ err) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= b) return;
    uint32_t v[8], p[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    load_fe(v, values + 2 * i);
    if (!to_int(v, fmt)) atomicOr(err, kErrNonCanonical);
    const size_t j = lower_bound(keys, m, v);
    uint8_t f = 0;
    uint64_t ps = 0, ss = 0;
    size_t js = j;
    if (j < m) {
        uint32_t w[8];
        load_fe(w, keys + 2 * j);
        if (cmp256(w, v) == 0) f |= 2, js = j + 1;
    }
    if (j >= 1) load_fe(p, keys + 2 * (j - 1)), ps = base + slots[j - 1], f |= 1;
    if (js < m) load_fe(q, keys + 2 * js), ss = base + slots[js], f |= 4;
    store_fe(pred_keys + 2 * i, p);
    store_fe(succ_keys + 2 * i, q);
    pred_slots[i] = ps, succ_slots[i] = ss, flags[i] = f;
}

// Sharded inserts, replicated half: merge the gathered per-rank neighbours ([world][b], rank-major) into the global
// ones, then the same resolution as k_ins_resolve. Zero, present or repeated values raise kErrBadInsert.
__global__ void __launch_bounds__(128) k_ins_plan(const uint4* __restrict__ vals, size_t b, uint64_t first_idx, int fmt, unsigned world,
                                                  const uint4* __restrict__ pred_keys, const uint64_t* __restrict__ pred_slots,
                                                  const uint4* __restrict__ succ_keys, const uint64_t* __restrict__ succ_slots,
                                                  const uint8_t* __restrict__ flags, uint64_t* __restrict__ x, uint4* __restrict__ upd,
                                                  uint4* __restrict__ low_old, uint8_t* __restrict__ is_largest, uint32_t* __restrict__ err) {
    const size_t k = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;  // a warp per insert (see ins_finish)
    if (k >= b) return;
    uint32_t v[8], pred[8] = {0, 0, 0, 0, 0, 0, 0, 0}, succ[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    load_fe(v, vals + 2 * k);
    bool has_pred = false, has_succ = false, bad = zero256(v);
    uint64_t pred_slot = 0, succ_slot = 0;
    for (unsigned r = 0; r < world; ++r) {
        const size_t o = (size_t)r * b + k;
        const uint8_t f = flags[o];
        bad |= (f & 2) != 0;
        uint32_t w[8];
        if (f & 1) {
            load_fe(w, pred_keys + 2 * o);
            if (!has_pred || cmp256(w, pred) > 0) copy256(pred, w), pred_slot = pred_slots[o], has_pred = true;
        }
        if (f & 4) {
            load_fe(w, succ_keys + 2 * o);
            if (!has_succ || cmp256(w, succ) < 0) copy256(succ, w), succ_slot = succ_slots[o], has_succ = true;
        }
    }
    bad |= !has_pred;  // nothing below v anywhere: the tree has no head {0, ..}
    if (!ins_finish(k, v, pred, pred_slot, succ, succ_slot, has_succ, vals, first_idx, fmt, x, upd, low_old, is_largest)) bad = true;
    if (bad) atomicOr(err, kErrBadInsert);
}

// per write: the root after it (top versions) in the context format
__global__ void __launch_bounds__(256) k_export_versions(const uint4* __restrict__ ver_level, unsigned writes, int fmt, uint4* __restrict__ out) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= writes) return;
    uint32_t r[8];
    load_fe(r, ver_level + 2 * (size_t)t);
    egress(r, fmt);
    store_fe(out + 2 * (size_t)t, r);
}

// For write t and level l: prev = latest earlier write whose level-l node is the SIBLING of t's node (-1: none, take the
// stored tree); last = no later write touches t's own level-l node (t's version is the final one).
// Computed level by level over the writes kept sorted by (level-l node, time): in that order the writes of a node form a
// segment sorted by time, the sibling's segment is adjacent, and
//     rank r of t among the sibling's times   =>   prev = the sibling segment's element r - 1
//     position of t in the PARENT's segment   =   start of the even child's segment + (index in own segment) + r
// i.e. one step of a bottom-up merge sort per level: O(writes x log) per level instead of the pairwise O(writes^2) scan
// (0.9 ms per level-set at 8192 writes, 14 ms at 32768). Keys pack (slot << 20 | time): at most 2^20 writes per chunk, slots < 2^32.
constexpr unsigned kLinkTimeBits = 20;
__global__ void __launch_bounds__(256) k_ins_link_keys(const uint64_t* __restrict__ x, unsigned writes, uint64_t* __restrict__ keys) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < writes) keys[t] = (x[t] << kLinkTimeBits) | t;
}
// first position whose level-l node is >= v
__device__ __forceinline__ unsigned link_lower_bound(const uint64_t* __restrict__ a, unsigned writes, unsigned shift, uint64_t v) {
    unsigned lo = 0, hi = writes;
    while (lo < hi) {
        const unsigned mid = (lo + hi) >> 1;
        if ((a[mid] >> shift) < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
// prev_own (optional) = the latest earlier write to t's OWN level-l node (-1: none): the node's value BEFORE write t, which the
// witness trace of insert_leaf needs (the folds of the old low leaf and of the empty leaf, indexed_merkle_tree.rs:196-204, 286-294)
__global__ void __launch_bounds__(256) k_ins_links_level(const uint64_t* __restrict__ a, uint64_t* __restrict__ a_next, unsigned writes,
                                                         unsigned depth, unsigned l, int* __restrict__ prev, uint8_t* __restrict__ last,
                                                         int* __restrict__ prev_own) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= writes) return;
    const unsigned shift = kLinkTimeBits + l;
    const uint64_t key = a[i], v = key >> shift;
    const unsigned t = (unsigned)(key & ((1u << kLinkTimeBits) - 1));
    const unsigned own = link_lower_bound(a, writes, shift, v);
    const unsigned sb = link_lower_bound(a, writes, shift, v ^ 1), se = link_lower_bound(a, writes, shift, (v ^ 1) + 1);
    unsigned lo = sb, hi = se;  // rank of t among the sibling segment's times (the segment is sorted by time)
    while (lo < hi) {
        const unsigned mid = (lo + hi) >> 1;
        if ((unsigned)(a[mid] & ((1u << kLinkTimeBits) - 1)) < t) lo = mid + 1;
        else hi = mid;
    }
    const unsigned r = lo - sb;
    const size_t id = (size_t)t * (depth + 1) + l;
    prev[id] = r ? (int)(a[sb + r - 1] & ((1u << kLinkTimeBits) - 1)) : -1;
    last[id] = (i + 1 == writes || (a[i + 1] >> shift) != v) ? 1 : 0;
    if (prev_own) prev_own[id] = i > own ? (int)(a[i - 1] & ((1u << kLinkTimeBits) - 1)) : -1;  // the own segment is sorted by time
    if (l < depth) a_next[((v & 1) ? sb : own) + (i - own) + r] = key;  // merged by time into the parent's segment
}

// One level of all writes, operand half: pairs[t] = (left, right) children of write t's level-(l+1) node — its own
// child version and the sibling (latest earlier write to the sibling node, else the stored tree). The sibling is also
// the witness path element of that write (low path of insert t/2 in the OLD tree for even t, new-leaf path in the NEW
// tree for odd t — see the header of this file). The hashing half is the ordinary level kernel over `pairs`
// (imt_host::launch_level: with <= 8192 writes that is the 3-lanes-per-hash latency kernel).
// With fold_nodes (and prev_own) the chain values of the four folds the chip's insert_leaf constrains are kept for the one-launch witness
// trace (imt_insert_witness_trace): fold_nodes[t / 2][2 (t & 1) + {0, 1}][l] = the level-l node on write t's path BEFORE the write (latest
// earlier write to the same node, else the stored tree) and AFTER it, in the context format.
__global__ void __launch_bounds__(256) k_ins_pairs(const uint4* __restrict__ ver_level, const uint4* __restrict__ tree_levels, size_t n,
                                                   unsigned l, const uint64_t* __restrict__ x, const int* __restrict__ prev, unsigned writes,
                                                   unsigned depth, int fmt, uint4* __restrict__ pairs, uint4* __restrict__ sib_low,
                                                   uint4* __restrict__ sib_new, uint4* __restrict__ sib_all,
                                                   const int* __restrict__ prev_own = nullptr, uint4* __restrict__ fold_nodes = nullptr) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= writes) return;
    const uint64_t node = x[t] >> l;
    const int p = prev[(size_t)t * (depth + 1) + l];
    uint32_t own[8], sib[8];
    load_fe(own, ver_level + 2 * (size_t)t);
    if (p >= 0) load_fe(sib, ver_level + 2 * (size_t)p);
    else load_fe(sib, tree_levels + 2 * (level_offset(n, l) + (node ^ 1)));
    const bool left = (node & 1) == 0;
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        lo[i] = left ? own[i] : sib[i];
        hi[i] = left ? sib[i] : own[i];
    }
    store_fe(pairs + 4 * (size_t)t, lo);
    store_fe(pairs + 4 * (size_t)t + 2, hi);
    if (fold_nodes) {
        uint32_t before[8];
        const int q = prev_own[(size_t)t * (depth + 1) + l];
        if (q >= 0) load_fe(before, ver_level + 2 * (size_t)q);
        else load_fe(before, tree_levels + 2 * (level_offset(n, l) + node));
        egress(before, fmt);
        egress(own, fmt);
        uint4* dst = fold_nodes + 2 * (((size_t)(t >> 1) * 4 + 2 * (t & 1)) * depth + l);
        store_fe(dst, before);
        store_fe(dst + 2 * (size_t)depth, own);
    }
    uint4* wit = (t & 1) ? sib_new : sib_low;
    if (wit || sib_all) egress(sib, fmt);
    if (wit) store_fe(wit + 2 * ((size_t)(t >> 1) * depth + l), sib);
    if (sib_all) store_fe(sib_all + 2 * ((size_t)t * depth + l), sib);  // one row per write (sharded inserts)
}

// per insert: roots before / after, helper bits of both paths
__global__ void __launch_bounds__(256) k_ins_outputs(const uint4* __restrict__ ver, const uint4* __restrict__ tree_levels, size_t n,
                                                     const uint64_t* __restrict__ x, unsigned b, unsigned depth, int fmt,
                                                     uint4* __restrict__ old_roots, uint4* __restrict__ new_roots,
                                                     uint8_t* __restrict__ low_helpers, uint8_t* __restrict__ new_helpers,
                                                     uint64_t* __restrict__ low_idx) {
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= b) return;
    const unsigned writes = 2 * b;
    const uint4* top = ver + 2 * ((size_t)depth * writes);
    uint32_t r[8];
    if (old_roots) {
        if (k == 0) load_fe(r, tree_levels + 2 * level_offset(n, depth));
        else load_fe(r, top + 2 * (2 * k - 1));
        egress(r, fmt);
        store_fe(old_roots + 2 * k, r);
    }
    if (new_roots) {
        load_fe(r, top + 2 * (2 * k + 1));
        egress(r, fmt);
        store_fe(new_roots + 2 * k, r);
    }
    const uint64_t lo = x[2 * k], nw = x[2 * k + 1];
    if (low_idx) low_idx[k] = lo;
    for (unsigned l = 0; l < depth; ++l) {
        if (low_helpers) low_helpers[(size_t)k * depth + l] = ((lo >> l) & 1) == 0;  // 1 = current node is LEFT (utils.rs:70)
        if (new_helpers) new_helpers[(size_t)k * depth + l] = ((nw >> l) & 1) == 0;
    }
}

// final versions into the stored tree and preimages
__global__ void __launch_bounds__(256) k_ins_commit(const uint4* __restrict__ ver, const uint4* __restrict__ upd, const uint64_t* __restrict__ x,
                                                    const uint8_t* __restrict__ last, unsigned writes, unsigned depth, size_t n,
                                                    uint4* __restrict__ tree_levels, uint4* __restrict__ pre) {
    const unsigned levels = depth + 1;
    const size_t id = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (id >= (size_t)writes * levels) return;
    if (!last[id]) return;
    const unsigned t = (unsigned)(id / levels), l = (unsigned)(id % levels);
    const uint64_t node = x[t] >> l;
    const uint4* v = ver + 2 * ((size_t)l * writes + t);
    uint4* dst = tree_levels + 2 * (level_offset(n, l) + node);
    dst[0] = v[0], dst[1] = v[1];
    if (l == 0 && pre) {
#pragma unroll
        for (int k = 0; k < 6; ++k) pre[6 * node + k] = upd[6 * (size_t)t + k];
    }
}

// merge of two sorted, disjoint key arrays by ranking: every element's output position is its own rank plus the
// number of smaller elements in the other array
__global__ void __launch_bounds__(256) k_merge_rank(const uint4* __restrict__ ka, const uint32_t* __restrict__ sa, size_t na,
                                                    const uint4* __restrict__ kb, const uint32_t* __restrict__ sb, size_t nb,
                                                    uint4* __restrict__ ko, uint32_t* __restrict__ so) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= na + nb) return;
    uint32_t v[8];
    size_t pos;
    uint32_t slot;
    if (i < na) {
        load_fe(v, ka + 2 * i);
        pos = i + lower_bound(kb, nb, v);
        slot = sa[i];
    } else {
        const size_t j = i - na;
        load_fe(v, kb + 2 * j);
        pos = j + lower_bound(ka, na, v);
        slot = sb[j];
    }
    store_fe(ko + 2 * pos, v);
    so[pos] = slot;
}

__global__ void k_iota_slots(uint32_t* s, size_t b, uint64_t first) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < b) s[i] = (uint32_t)(first + i);
}

// ------------------------------------------------------------------------------------------------- host side
// Levels 1..depth of `writes` single-leaf writes to a tree (`tree_levels`, n leaves) from their level-0 versions in
// ver[0 .. writes): links, then per level the operand gather + one level launch. ver holds (depth+1) x writes FE.
imt_status versioned_levels(imt_ctx* ctx, const Fr* tree_levels, size_t n, unsigned depth, const uint64_t* d_x, unsigned writes, Fr* d_ver,
                            int* d_prev, uint8_t* d_last, Fr* d_pairs, uint4* sib_low, uint4* sib_new, uint4* sib_all,
                            int* d_prev_own = nullptr, uint4* fold_nodes = nullptr) {
    if (writes > (1u << kLinkTimeBits)) return fail(ctx, IMT_ERR_INVALID_ARG, "more than 2^20 writes in one insert chunk");
    {  // links of every (write, level): sort the writes by (slot, time) once, then one merge step per level
        DevBuf keys_a(ctx), keys_b(ctx), temp(ctx);
        IMT_TRY_CUDA(ctx, keys_a.alloc(writes * sizeof(uint64_t)));
        IMT_TRY_CUDA(ctx, keys_b.alloc(writes * sizeof(uint64_t)));
        k_ins_link_keys<<<grid_for(writes, 256), 256, 0, ctx->stream>>>(d_x, writes, keys_a.as<uint64_t>());
        size_t temp_bytes = 0;
        IMT_TRY_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, keys_a.as<uint64_t>(), keys_b.as<uint64_t>(), (int)writes, 0, 64,
                                                         ctx->stream));
        IMT_TRY_CUDA(ctx, temp.alloc(temp_bytes));
        IMT_TRY_CUDA(ctx, cub::DeviceRadixSort::SortKeys(temp.p, temp_bytes, keys_a.as<uint64_t>(), keys_b.as<uint64_t>(), (int)writes, 0, 64,
                                                         ctx->stream));
        uint64_t* cur = keys_b.as<uint64_t>();
        uint64_t* nxt = keys_a.as<uint64_t>();
        for (unsigned l = 0; l <= depth; ++l) {
            k_ins_links_level<<<grid_for(writes, 256), 256, 0, ctx->stream>>>(cur, nxt, writes, depth, l, d_prev, d_last, d_prev_own);
            std::swap(cur, nxt);
        }
        ctx->launches += depth + 4;
        IMT_TRY_CUDA(ctx, cudaGetLastError());
    }  // the scratch is freed in stream order
    for (unsigned l = 0; l < depth; ++l) {
        k_ins_pairs<<<grid_for(writes, 256), 256, 0, ctx->stream>>>((const uint4*)d_ver + 2 * ((size_t)l * writes), (const uint4*)tree_levels, n, l,
                                                                    d_x, d_prev, writes, depth, ctx->fmt, (uint4*)d_pairs, sib_low, sib_new, sib_all,
                                                                    d_prev_own, fold_nodes);
        ++ctx->launches;
        IMT_TRY_CUDA(ctx, cudaGetLastError());
        IMT_TRY(launch_level(ctx, d_pairs, d_ver + (size_t)(l + 1) * writes, writes));
    }
    return IMT_OK;
}

imt_status sort_pairs(imt_ctx* ctx, Fr* d_keys, uint32_t* d_slots, size_t count) {
    if (count < 2) return IMT_OK;
    size_t temp_bytes = 0;
    IMT_TRY_CUDA(ctx, cub::DeviceMergeSort::SortPairs(nullptr, temp_bytes, d_keys, d_slots, (long long)count, KeyLess(), ctx->stream));
    DevBuf temp(ctx);
    IMT_TRY_CUDA(ctx, temp.alloc(temp_bytes));
    IMT_TRY_CUDA(ctx, cub::DeviceMergeSort::SortPairs(temp.p, temp_bytes, d_keys, d_slots, (long long)count, KeyLess(), ctx->stream));
    ctx->launches += 2;  // block sort + merge passes (at least)
    IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // temp is freed on return
    return IMT_OK;
}

}  // namespace
imt_status imt_host::ensure_index(imt_tree* t) {
    imt_ctx* ctx = t->ctx;
    if (t->index_valid) return IMT_OK;
    if (!t->d_pre) return fail(ctx, IMT_ERR_INVALID_ARG, "tree was not built from leaves: no preimages to index");
    if (t->n > 0xffffffffull) return fail(ctx, IMT_ERR_INVALID_ARG, "index supports at most 2^32 slots");
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!t->d_sorted_keys || t->index_capacity < t->n) {
        if (t->d_sorted_keys) tree_free(ctx, t->d_sorted_keys), t->d_sorted_keys = nullptr;
        if (t->d_sorted_slots) tree_free(ctx, t->d_sorted_slots), t->d_sorted_slots = nullptr;
        if (t->d_alt_keys) tree_free(ctx, t->d_alt_keys), t->d_alt_keys = nullptr;
        if (t->d_alt_slots) tree_free(ctx, t->d_alt_slots), t->d_alt_slots = nullptr;
        t->index_capacity = 0;
        IMT_TRY_CUDA(ctx, tree_malloc(ctx, (void**)&t->d_sorted_keys, t->n * sizeof(Fr)));
        IMT_TRY_CUDA(ctx, tree_malloc(ctx, (void**)&t->d_sorted_slots, t->n * sizeof(uint32_t)));
        t->index_capacity = t->n;
    }
    DevBuf stats(ctx);  // [0] first empty, [1] last occupied, then the head flag
    IMT_TRY_CUDA(ctx, stats.alloc(3 * sizeof(unsigned long long)));
    const unsigned long long init[3] = {(unsigned long long)t->n, 0ull, 0ull};
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(stats.p, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY(clear_err(ctx));
    k_index_extract<<<grid_for(t->n, 256), 256, 0, ctx->stream>>>((const uint4*)t->d_pre, t->n, (uint64_t)t->rank * t->n, ctx->fmt, (uint4*)t->d_sorted_keys,
                                                                 t->d_sorted_slots, stats.as<unsigned long long>(),
                                                                 (uint32_t*)(stats.as<unsigned long long>() + 2), ctx->d_err);
    ++ctx->launches;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    unsigned long long got[3];
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(got, stats.p, sizeof(got), cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY(finish(ctx));
    const size_t m = (size_t)got[0];
    if (got[1] > m) return fail(ctx, IMT_ERR_NOT_WELL_FORMED, "occupied slots do not form a prefix");
    IMT_TRY(sort_pairs(ctx, t->d_sorted_keys, t->d_sorted_slots, m));
    IMT_TRY(clear_err(ctx));
    if (m) k_index_check<<<grid_for(m, 256), 256, 0, ctx->stream>>>((const uint4*)t->d_pre, ctx->fmt, (const uint4*)t->d_sorted_keys,
                                                             t->d_sorted_slots, m, t->world == 1, ctx->d_err);
    ++ctx->launches;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    IMT_TRY(finish(ctx));
    t->occupied = m;
    t->head_next_zero = (uint32_t)got[2] != 0;
    t->index_valid = true;
    t->prefix_valid = false;
    return IMT_OK;
}
namespace {

// the prefix array + its top samples, rebuilt when the sorted keys changed (one pass over the keys: ~0.1 ms at depth 24)
imt_status ensure_prefix(imt_tree* t) {
    imt_ctx* ctx = t->ctx;
    if (t->prefix_valid) return IMT_OK;
    if (t->prefix_capacity < t->index_capacity) {
        if (t->d_prefix) tree_free(ctx, t->d_prefix), t->d_prefix = nullptr;
        IMT_TRY_CUDA(ctx, tree_malloc(ctx, (void**)&t->d_prefix, t->index_capacity * sizeof(unsigned long long)));
        t->prefix_capacity = t->index_capacity;
    }
    if (!t->d_top) IMT_TRY_CUDA(ctx, tree_malloc(ctx, (void**)&t->d_top, kIndexTop * sizeof(unsigned long long)));
    const size_t m = t->occupied;
    t->top_stride = (m + kIndexTop - 1) / kIndexTop;
    if (t->top_stride == 0) t->top_stride = 1;
    if (m) k_index_prefix<<<grid_for(m, 256), 256, 0, ctx->stream>>>((const uint4*)t->d_sorted_keys, m, t->d_prefix);
    k_index_top<<<grid_for(kIndexTop, 256), 256, 0, ctx->stream>>>(t->d_prefix, m, t->top_stride, t->d_top);
    ctx->launches += 2;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    t->prefix_valid = true;
    return IMT_OK;
}

imt_status lookup_dev(imt_tree* t, const void* d_values, size_t q, uint64_t* d_low, uint8_t* d_matched) {
    imt_ctx* ctx = t->ctx;
    // small indices sit in L1 / L2 anyway: the plain search is as good. IMT_FAST_LOOKUP_MIN moves the threshold (tests force the fast
    // path on small trees with 0; a huge value gives the plain search for A/B measurements); read per call, it is one getenv.
    const char* env_min = std::getenv("IMT_FAST_LOOKUP_MIN");
    const size_t fast_min = env_min ? (size_t)std::strtoull(env_min, nullptr, 10) : 65536;
    if (t->occupied >= fast_min && t->occupied > 0) {
        IMT_TRY(ensure_prefix(t));
        k_low_leaf_lookup_fast<<<grid_for(q, 512), 512, 0, ctx->stream>>>((const uint4*)t->d_sorted_keys, t->d_sorted_slots, t->d_prefix, t->d_top,
                                                                          t->top_stride, t->occupied, t->n, t->head_next_zero ? 1u : 0u,
                                                                          (const uint4*)d_values, q, ctx->fmt, d_low, d_matched, ctx->d_err);
        ++ctx->launches;
        IMT_TRY_CUDA(ctx, cudaGetLastError());
        return IMT_OK;
    }
    k_low_leaf_lookup<<<grid_for(q, 256), 256, 0, ctx->stream>>>((const uint4*)t->d_sorted_keys, t->d_sorted_slots, t->occupied, t->n,
                                                                 t->head_next_zero ? 1u : 0u, (const uint4*)d_values, q, ctx->fmt, d_low,
                                                                 d_matched, ctx->d_err);
    ++ctx->launches;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    return IMT_OK;
}

bool sharded(const imt_tree* t) { return t->world > 1; }

}  // namespace

namespace imt_host {
imt_status launch_limb_witness(imt_ctx* ctx, const void* d_low_leaves, const void* d_new_vals, size_t new_stride, size_t b, void* d_limbs,
                               uint8_t* d_flags) {
    if (b == 0) return IMT_OK;
    k_limb_witness<<<grid_for(b, 256), 256, 0, ctx->stream>>>((const uint4*)d_low_leaves, (const uint4*)d_new_vals, new_stride, b, ctx->fmt,
                                                              (uint4*)d_limbs, d_flags, ctx->d_err);
    ++ctx->launches;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    return IMT_OK;
}
// lookup + gather of everything verify_non_inclusion loads about the low leaf, queued on the compute stream (no wait): the caller
// has cleared the error word and calls finish(). d_low is required; the other outputs may be null.
imt_status queue_non_inclusion(imt_tree* t, const void* d_values, size_t q, uint64_t* d_low, uint8_t* d_matched, void* d_low_leaves,
                               uint8_t* d_is_largest, void* d_siblings, uint8_t* d_helpers) {
    imt_ctx* ctx = t->ctx;
    if (sharded(t)) return fail(ctx, IMT_ERR_INVALID_ARG, "low-leaf lookups on a sharded tree go through the per-rank candidate call");
    IMT_TRY(ensure_index(t));
    if (q == 0) return IMT_OK;
    DevBuf dm(ctx);
    if (!d_matched) {  // the lookup kernels always write it
        IMT_TRY_CUDA(ctx, dm.alloc(q));
        d_matched = dm.as<uint8_t>();
    }
    IMT_TRY(lookup_dev(t, d_values, q, d_low, d_matched));
    if (d_low_leaves || d_is_largest) {
        k_gather_leaves<<<grid_for(q, 256), 256, 0, ctx->stream>>>((const uint4*)t->d_pre, t->n, 0, d_low, q, (uint4*)d_low_leaves, d_is_largest,
                                                                   ctx->d_err);
        ++ctx->launches;
        IMT_TRY_CUDA(ctx, cudaGetLastError());
    }
    if ((d_siblings || d_helpers) && t->depth) {
        DevBuf scratch(ctx);  // the gather kernel always writes siblings
        if (!d_siblings) {
            IMT_TRY_CUDA(ctx, scratch.alloc(q * (size_t)t->depth * sizeof(Fr)));
            d_siblings = scratch.p;
        }
        IMT_TRY(launch_gather_proofs(t, d_low, q, d_siblings, d_helpers, nullptr));
    }
    return IMT_OK;
}
void invalidate_index(imt_tree* t) {  // the buffers are kept for the next build of the index
    t->index_valid = false;
    t->prefix_valid = false;
    t->occupied = 0;
}
}  // namespace imt_host

extern "C" imt_status imt_tree_occupied(imt_tree* t, size_t* occupied) {
    if (!t || !occupied) return IMT_ERR_INVALID_ARG;
    IMT_TRY(ensure_index(t));
    *occupied = t->occupied;
    return IMT_OK;
}

extern "C" imt_status imt_low_leaf_lookup(imt_tree* t, const void* values, size_t q, uint64_t* low_idx, uint8_t* matched) {
    if (!t) return IMT_ERR_INVALID_ARG;
    imt_ctx* ctx = t->ctx;
    if (q && (!values || !low_idx)) return fail(ctx, IMT_ERR_INVALID_ARG, "null buffer");
    if (sharded(t)) return fail(ctx, IMT_ERR_INVALID_ARG, "low-leaf lookups on a sharded tree go through the per-rank candidate call");
    IMT_TRY(ensure_index(t));
    if (q == 0) return IMT_OK;
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf dv(ctx), dl(ctx), dm(ctx);
    IMT_TRY_CUDA(ctx, dv.alloc(q * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, dl.alloc(q * sizeof(uint64_t)));
    IMT_TRY_CUDA(ctx, dm.alloc(q));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(dv.p, values, q * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY(clear_err(ctx));
    IMT_TRY(lookup_dev(t, dv.p, q, dl.as<uint64_t>(), dm.as<uint8_t>()));
    IMT_TRY(finish(ctx));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(low_idx, dl.p, q * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (matched) IMT_TRY_CUDA(ctx, cudaMemcpyAsync(matched, dm.p, q, cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IMT_OK;
}

extern "C" imt_status imt_low_leaf_lookup_dev(imt_tree* t, const void* d_values, size_t q, uint64_t* d_low_idx, uint8_t* d_matched) {
    if (!t) return IMT_ERR_INVALID_ARG;
    imt_ctx* ctx = t->ctx;
    if (q && (!d_values || !d_low_idx)) return fail(ctx, IMT_ERR_INVALID_ARG, "null buffer");
    if (sharded(t)) return fail(ctx, IMT_ERR_INVALID_ARG, "low-leaf lookups on a sharded tree go through the per-rank candidate call");
    IMT_TRY(ensure_index(t));
    if (q == 0) return IMT_OK;
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    IMT_TRY(clear_err(ctx));
    IMT_TRY(lookup_dev(t, d_values, q, d_low_idx, d_matched));
    return finish(ctx);
}

extern "C" imt_status imt_non_inclusion_paths(imt_tree* t, const void* values, size_t q, uint64_t* low_idx, uint8_t* matched,
                                              void* low_leaves, void* siblings, uint8_t* helpers, uint8_t* is_largest) {
    if (!t) return IMT_ERR_INVALID_ARG;
    imt_ctx* ctx = t->ctx;
    if (q && !values) return fail(ctx, IMT_ERR_INVALID_ARG, "null buffer");
    if (sharded(t)) return fail(ctx, IMT_ERR_INVALID_ARG, "low-leaf lookups on a sharded tree go through the per-rank candidate call");
    IMT_TRY(ensure_index(t));
    if (q == 0) return IMT_OK;
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    const unsigned depth = t->depth;
    DevBuf dv(ctx), dl(ctx), dm(ctx), dlv(ctx), dsib(ctx), dhel(ctx), dlg(ctx);
    IMT_TRY_CUDA(ctx, dv.alloc(q * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, dl.alloc(q * sizeof(uint64_t)));
    IMT_TRY_CUDA(ctx, dm.alloc(q));
    if (low_leaves) IMT_TRY_CUDA(ctx, dlv.alloc(q * 3 * sizeof(Fr)));
    if (siblings) IMT_TRY_CUDA(ctx, dsib.alloc(q * (size_t)depth * sizeof(Fr)));
    if (helpers) IMT_TRY_CUDA(ctx, dhel.alloc(q * (size_t)depth));
    if (is_largest) IMT_TRY_CUDA(ctx, dlg.alloc(q));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(dv.p, values, q * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY(clear_err(ctx));
    IMT_TRY(lookup_dev(t, dv.p, q, dl.as<uint64_t>(), dm.as<uint8_t>()));
    if (low_leaves || is_largest) {
        k_gather_leaves<<<grid_for(q, 256), 256, 0, ctx->stream>>>((const uint4*)t->d_pre, t->n, 0, dl.as<uint64_t>(), q,
                                                                   low_leaves ? dlv.as<uint4>() : nullptr,
                                                                   is_largest ? dlg.as<uint8_t>() : nullptr, ctx->d_err);
        ++ctx->launches;
        IMT_TRY_CUDA(ctx, cudaGetLastError());
    }
    if (siblings || helpers) {
        DevBuf scratch(ctx);  // the gather kernel always writes siblings
        void* d_sib = dsib.p;
        if (!siblings) {
            IMT_TRY_CUDA(ctx, scratch.alloc(q * (size_t)depth * sizeof(Fr)));
            d_sib = scratch.p;
        }
        IMT_TRY(launch_gather_proofs(t, dl.as<uint64_t>(), q, d_sib, helpers ? dhel.as<uint8_t>() : nullptr, nullptr));
        IMT_TRY(finish(ctx));
    } else {
        IMT_TRY(finish(ctx));
    }
    if (low_idx) IMT_TRY_CUDA(ctx, cudaMemcpyAsync(low_idx, dl.p, q * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (matched) IMT_TRY_CUDA(ctx, cudaMemcpyAsync(matched, dm.p, q, cudaMemcpyDeviceToHost, ctx->stream));
    if (low_leaves) IMT_TRY_CUDA(ctx, cudaMemcpyAsync(low_leaves, dlv.p, q * 3 * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    if (siblings && depth) IMT_TRY_CUDA(ctx, cudaMemcpyAsync(siblings, dsib.p, q * (size_t)depth * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    if (helpers && depth) IMT_TRY_CUDA(ctx, cudaMemcpyAsync(helpers, dhel.p, q * (size_t)depth, cudaMemcpyDeviceToHost, ctx->stream));
    if (is_largest) IMT_TRY_CUDA(ctx, cudaMemcpyAsync(is_largest, dlg.p, q, cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IMT_OK;
}

extern "C" imt_status imt_insert_batch(imt_tree* t, const void* new_vals, size_t b, uint64_t first_idx, imt_insert_witness* w) {
    if (!t) return IMT_ERR_INVALID_ARG;
    imt_ctx* ctx = t->ctx;
    if (b && !new_vals) return fail(ctx, IMT_ERR_INVALID_ARG, "null buffer");
    if (sharded(t)) return fail(ctx, IMT_ERR_INVALID_ARG, "inserts into a sharded tree are not supported");
    IMT_TRY(ensure_index(t));
    if (first_idx != t->occupied) return fail(ctx, IMT_ERR_INVALID_ARG, "first_idx must be the next free slot (the number of occupied slots)");
    if (b > t->n - t->occupied) return fail(ctx, IMT_ERR_TREE_FULL, imt_status_string(IMT_ERR_TREE_FULL));
    if (b == 0) return IMT_OK;
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    const unsigned depth = t->depth;
    static const bool dbg = std::getenv("IMT_DEBUG_TIMING") != nullptr;
    auto tp0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!dbg) return;
        cudaStreamSynchronize(ctx->stream);
        auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[imt_insert_batch] %-12s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - tp0).count());
        tp0 = now;
    };
    const imt_insert_witness none = {};
    const imt_insert_witness out = w ? *w : none;

    // ---- the whole batch as canonical integers, validated before anything is modified
    DevBuf staged(ctx), vals(ctx), sorted_vals(ctx), sorted_slots(ctx);
    IMT_TRY_CUDA(ctx, staged.alloc(b * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, vals.alloc(b * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, sorted_vals.alloc(b * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, sorted_slots.alloc(b * sizeof(uint32_t)));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(staged.p, new_vals, b * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY(clear_err(ctx));
    IMT_TRY(launch_convert(ctx, staged.p, vals.p, b, ctx->fmt, kFmtCanonical));
    IMT_TRY(finish(ctx));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(sorted_vals.p, vals.p, b * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
    k_iota_slots<<<grid_for(b, 256), 256, 0, ctx->stream>>>(sorted_slots.as<uint32_t>(), b, first_idx);
    ++ctx->launches;
    IMT_TRY(sort_pairs(ctx, sorted_vals.as<Fr>(), sorted_slots.as<uint32_t>(), b));
    IMT_TRY(clear_err(ctx));
    k_ins_validate<<<grid_for(b, 256), 256, 0, ctx->stream>>>(sorted_vals.as<uint4>(), b, (const uint4*)t->d_sorted_keys, t->occupied, ctx->d_err);
    ++ctx->launches;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    IMT_TRY(finish(ctx));

    lap("validate");
    // ---- per-chunk scratch
    const char* chunk_env = std::getenv("IMT_INSERT_CHUNK");  // read per call: tests force small chunks to cross chunk boundaries
    const size_t chunk_override = chunk_env ? (size_t)std::strtoull(chunk_env, nullptr, 10) : 0;
    const size_t C = std::min(b, chunk_override ? std::min(chunk_override, (size_t)1 << (kLinkTimeBits - 1)) : kInsertChunk), W = 2 * C, L = depth + 1;
    DevBuf x(ctx), upd(ctx), low_old(ctx), largest(ctx), prev(ctx), last(ctx), ver(ctx), sib_low(ctx), sib_new(ctx), r_old(ctx), r_new(ctx), h_low(ctx), h_new(ctx), low_idx(ctx), chunk_keys(ctx), chunk_slots(ctx), pairs(ctx), prev_own(ctx), fold(ctx);
    IMT_TRY_CUDA(ctx, x.alloc(W * sizeof(uint64_t)));
    IMT_TRY_CUDA(ctx, upd.alloc(W * 3 * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, low_old.alloc(C * 3 * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, largest.alloc(C));
    IMT_TRY_CUDA(ctx, prev.alloc(W * L * sizeof(int)));
    IMT_TRY_CUDA(ctx, last.alloc(W * L));
    IMT_TRY_CUDA(ctx, ver.alloc(L * W * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, pairs.alloc(2 * W * sizeof(Fr)));
    if (out.low_siblings) IMT_TRY_CUDA(ctx, sib_low.alloc(C * depth * sizeof(Fr)));
    if (out.new_siblings) IMT_TRY_CUDA(ctx, sib_new.alloc(C * depth * sizeof(Fr)));
    if (out.fold_nodes && depth) {
        IMT_TRY_CUDA(ctx, prev_own.alloc(W * L * sizeof(int)));
        IMT_TRY_CUDA(ctx, fold.alloc(C * 4 * depth * sizeof(Fr)));
    }
    if (out.old_roots) IMT_TRY_CUDA(ctx, r_old.alloc(C * sizeof(Fr)));
    if (out.new_roots) IMT_TRY_CUDA(ctx, r_new.alloc(C * sizeof(Fr)));
    if (out.low_helpers) IMT_TRY_CUDA(ctx, h_low.alloc(C * depth));
    if (out.new_helpers) IMT_TRY_CUDA(ctx, h_new.alloc(C * depth));
    if (out.low_idx) IMT_TRY_CUDA(ctx, low_idx.alloc(C * sizeof(uint64_t)));
    IMT_TRY_CUDA(ctx, chunk_keys.alloc(C * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, chunk_slots.alloc(C * sizeof(uint32_t)));
    if (!t->d_alt_keys) IMT_TRY_CUDA(ctx, tree_malloc(ctx, (void**)&t->d_alt_keys, t->index_capacity * sizeof(Fr)));
    if (!t->d_alt_slots) IMT_TRY_CUDA(ctx, tree_malloc(ctx, (void**)&t->d_alt_slots, t->index_capacity * sizeof(uint32_t)));

    lap("alloc");
    for (size_t off = 0; off < b; off += C) {
        const size_t cb = std::min(C, b - off);
        const unsigned writes = (unsigned)(2 * cb);
        const uint64_t first = first_idx + off;
        const uint4* cvals = vals.as<uint4>() + 2 * off;
        k_ins_resolve<<<grid_for(cb * 32, 128), 128, 0, ctx->stream>>>((const uint4*)t->d_sorted_keys, t->d_sorted_slots, t->occupied, cvals, cb, first,
                                                                 ctx->fmt, x.as<uint64_t>(), upd.as<uint4>(), low_old.as<uint4>(),
                                                                 largest.as<uint8_t>());
        ++ctx->launches;
        IMT_TRY_CUDA(ctx, cudaGetLastError());
        lap("resolve");
        // version 0 of every write: the hash of its leaf preimage (IMT:662-671); then all levels
        IMT_TRY(launch_hash(ctx, 3, upd.p, ver.p, writes, ctx->fmt, kFmtMontgomery, ctx->stream));
        IMT_TRY(versioned_levels(ctx, t->d_levels, t->n, depth, x.as<uint64_t>(), writes, ver.as<Fr>(), prev.as<int>(), last.as<uint8_t>(),
                                 pairs.as<Fr>(), out.low_siblings ? sib_low.as<uint4>() : nullptr,
                                 out.new_siblings ? sib_new.as<uint4>() : nullptr, nullptr, fold.p ? prev_own.as<int>() : nullptr,
                                 fold.p ? fold.as<uint4>() : nullptr));
        k_ins_outputs<<<grid_for(cb, 256), 256, 0, ctx->stream>>>(ver.as<uint4>(), (const uint4*)t->d_levels, t->n, x.as<uint64_t>(), (unsigned)cb,
                                                                  depth, ctx->fmt, out.old_roots ? r_old.as<uint4>() : nullptr,
                                                                  out.new_roots ? r_new.as<uint4>() : nullptr,
                                                                  out.low_helpers ? h_low.as<uint8_t>() : nullptr,
                                                                  out.new_helpers ? h_new.as<uint8_t>() : nullptr,
                                                                  out.low_idx ? low_idx.as<uint64_t>() : nullptr);
        k_ins_commit<<<grid_for((size_t)writes * L, 256), 256, 0, ctx->stream>>>(ver.as<uint4>(), upd.as<uint4>(), x.as<uint64_t>(),
                                                                                last.as<uint8_t>(), writes, depth, t->n, (uint4*)t->d_levels,
                                                                                (uint4*)t->d_pre);
        ctx->launches += 2;
        IMT_TRY_CUDA(ctx, cudaGetLastError());
        lap("levels+commit");
        // ---- witnesses of this chunk back to the caller
        auto d2h = [&](void* host, size_t stride, const void* dev) -> cudaError_t {
            if (!host || stride == 0) return cudaSuccess;
            return cudaMemcpyAsync((char*)host + off * stride, dev, cb * stride, cudaMemcpyDeviceToHost, ctx->stream);
        };
        IMT_TRY_CUDA(ctx, d2h(out.old_roots, sizeof(Fr), r_old.p));
        IMT_TRY_CUDA(ctx, d2h(out.new_roots, sizeof(Fr), r_new.p));
        IMT_TRY_CUDA(ctx, d2h(out.low_idx, sizeof(uint64_t), low_idx.p));
        IMT_TRY_CUDA(ctx, d2h(out.low_leaves, 3 * sizeof(Fr), low_old.p));
        IMT_TRY_CUDA(ctx, d2h(out.low_siblings, depth * sizeof(Fr), sib_low.p));
        IMT_TRY_CUDA(ctx, d2h(out.new_siblings, depth * sizeof(Fr), sib_new.p));
        IMT_TRY_CUDA(ctx, d2h(out.fold_nodes, 4 * (size_t)depth * sizeof(Fr), fold.p));
        IMT_TRY_CUDA(ctx, d2h(out.low_helpers, depth, h_low.p));
        IMT_TRY_CUDA(ctx, d2h(out.new_helpers, depth, h_new.p));
        IMT_TRY_CUDA(ctx, d2h(out.is_largest, 1, largest.p));
        if (out.new_leaves) {  // upd[2k+1], strided
            IMT_TRY_CUDA(ctx, cudaMemcpy2DAsync((char*)out.new_leaves + off * 3 * sizeof(Fr), 3 * sizeof(Fr), (const char*)upd.p + 3 * sizeof(Fr),
                                                6 * sizeof(Fr), 3 * sizeof(Fr), cb, cudaMemcpyDeviceToHost, ctx->stream));
        }
        lap("d2h");
        // ---- keep the sorted index current: merge this chunk's keys into it
        IMT_TRY_CUDA(ctx, cudaMemcpyAsync(chunk_keys.p, cvals, cb * sizeof(Fr), cudaMemcpyDeviceToDevice, ctx->stream));
        k_iota_slots<<<grid_for(cb, 256), 256, 0, ctx->stream>>>(chunk_slots.as<uint32_t>(), cb, first);
        ++ctx->launches;
        IMT_TRY(sort_pairs(ctx, chunk_keys.as<Fr>(), chunk_slots.as<uint32_t>(), cb));
        k_merge_rank<<<grid_for(t->occupied + cb, 256), 256, 0, ctx->stream>>>((const uint4*)t->d_sorted_keys, t->d_sorted_slots, t->occupied,
                                                                              chunk_keys.as<uint4>(), chunk_slots.as<uint32_t>(), cb,
                                                                              (uint4*)t->d_alt_keys, t->d_alt_slots);
        ++ctx->launches;
        IMT_TRY_CUDA(ctx, cudaGetLastError());
        IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        std::swap(t->d_sorted_keys, t->d_alt_keys);  // the merged arrays become the index; the old ones the next merge target
        std::swap(t->d_sorted_slots, t->d_alt_slots);
        t->prefix_valid = false;
        t->occupied += cb;
        t->head_next_zero = false;
        lap("merge");
    }
    IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    return IMT_OK;
}

// ------------------------------------------------------------------------------------------------- sharded lookups
extern "C" imt_status imt_tree_set_shard(imt_tree* t, unsigned rank, unsigned world) {
    if (!t) return IMT_ERR_INVALID_ARG;
    if (world == 0 || (world & (world - 1)) || rank >= world) return fail(t->ctx, IMT_ERR_INVALID_ARG, "bad rank/world");
    if (t->rank != rank || t->world != world) {
        t->cap_valid = false;
        invalidate_index(t);
    }
    t->rank = rank;
    t->world = world;
    return IMT_OK;
}

extern "C" imt_status imt_tree_head_next_zero(imt_tree* t, int* flag) {
    if (!t || !flag) return IMT_ERR_INVALID_ARG;
    IMT_TRY(ensure_index(t));
    *flag = t->head_next_zero ? 1 : 0;
    return IMT_OK;
}

extern "C" imt_status imt_low_leaf_candidates(imt_tree* t, const void* values, size_t q, void* cand_keys, uint64_t* cand_slots, uint8_t* flags) {
    if (!t) return IMT_ERR_INVALID_ARG;
    imt_ctx* ctx = t->ctx;
    if (q && (!values || !cand_keys || !cand_slots || !flags)) return fail(ctx, IMT_ERR_INVALID_ARG, "null buffer");
    IMT_TRY(ensure_index(t));
    if (q == 0) return IMT_OK;
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf dv(ctx), dk(ctx), ds(ctx), df(ctx);
    IMT_TRY_CUDA(ctx, dv.alloc(q * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, dk.alloc(q * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, ds.alloc(q * sizeof(uint64_t)));
    IMT_TRY_CUDA(ctx, df.alloc(q));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(dv.p, values, q * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY(clear_err(ctx));
    k_low_leaf_candidates<<<grid_for(q, 256), 256, 0, ctx->stream>>>((const uint4*)t->d_sorted_keys, t->d_sorted_slots, t->occupied,
                                                                     (uint64_t)t->rank * t->n, (const uint4*)dv.p, q, ctx->fmt, dk.as<uint4>(),
                                                                     ds.as<uint64_t>(), df.as<uint8_t>(), ctx->d_err);
    ++ctx->launches;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    IMT_TRY(finish(ctx));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(cand_keys, dk.p, q * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(cand_slots, ds.p, q * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(flags, df.p, q, cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IMT_OK;
}

extern "C" imt_status imt_low_leaf_merge(imt_ctx* ctx, const void* values, const void* cand_keys, const uint64_t* cand_slots, const uint8_t* flags,
                                         unsigned world, size_t q, uint64_t occupied_total, uint64_t n_total, int head_next_zero,
                                         uint64_t* low_idx, uint8_t* matched) {
    if (!ctx) return IMT_ERR_INVALID_ARG;
    if (q && (!values || !cand_keys || !cand_slots || !flags || !low_idx)) return fail(ctx, IMT_ERR_INVALID_ARG, "null buffer");
    if (q == 0 || world == 0) return IMT_OK;
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf dv(ctx), dz(ctx), dk(ctx), ds(ctx), df(ctx), dl(ctx), dm(ctx);
    IMT_TRY_CUDA(ctx, dv.alloc(q * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, dz.alloc(q));
    IMT_TRY_CUDA(ctx, dk.alloc(world * q * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, ds.alloc(world * q * sizeof(uint64_t)));
    IMT_TRY_CUDA(ctx, df.alloc(world * q));
    IMT_TRY_CUDA(ctx, dl.alloc(q * sizeof(uint64_t)));
    IMT_TRY_CUDA(ctx, dm.alloc(q));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(dv.p, values, q * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(dk.p, cand_keys, world * q * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(ds.p, cand_slots, world * q * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(df.p, flags, world * q, cudaMemcpyHostToDevice, ctx->stream));
    k_is_zero<<<grid_for(q, 256), 256, 0, ctx->stream>>>((const uint4*)dv.p, q, dz.as<uint8_t>());
    k_low_leaf_merge<<<grid_for(q, 256), 256, 0, ctx->stream>>>((const uint4*)dk.p, ds.as<uint64_t>(), df.as<uint8_t>(), world, q, occupied_total,
                                                                n_total, head_next_zero ? 1u : 0u, dz.as<uint8_t>(), dl.as<uint64_t>(),
                                                                dm.as<uint8_t>());
    ctx->launches += 2;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(low_idx, dl.p, q * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (matched) IMT_TRY_CUDA(ctx, cudaMemcpyAsync(matched, dm.p, q, cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IMT_OK;
}

extern "C" imt_status imt_tree_leaves(imt_tree* t, const uint64_t* indices, size_t q, void* leaves, uint8_t* is_largest) {
    if (!t) return IMT_ERR_INVALID_ARG;
    imt_ctx* ctx = t->ctx;
    if (q && !indices) return fail(ctx, IMT_ERR_INVALID_ARG, "null buffer");
    if (!t->d_pre) return fail(ctx, IMT_ERR_INVALID_ARG, "tree was not built from leaves");
    if (q == 0) return IMT_OK;
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf di(ctx), dl(ctx), dg(ctx);
    IMT_TRY_CUDA(ctx, di.alloc(q * sizeof(uint64_t)));
    if (leaves) IMT_TRY_CUDA(ctx, dl.alloc(q * 3 * sizeof(Fr)));
    if (is_largest) IMT_TRY_CUDA(ctx, dg.alloc(q));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(di.p, indices, q * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY(clear_err(ctx));
    k_gather_leaves<<<grid_for(q, 256), 256, 0, ctx->stream>>>((const uint4*)t->d_pre, t->n, (uint64_t)t->rank * t->n * (t->world > 1 ? 1 : 0),
                                                               di.as<uint64_t>(), q, leaves ? dl.as<uint4>() : nullptr,
                                                               is_largest ? dg.as<uint8_t>() : nullptr, ctx->d_err);
    ++ctx->launches;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    IMT_TRY(finish(ctx));
    if (leaves) IMT_TRY_CUDA(ctx, cudaMemcpyAsync(leaves, dl.p, q * 3 * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    if (is_largest) IMT_TRY_CUDA(ctx, cudaMemcpyAsync(is_largest, dg.p, q, cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IMT_OK;
}

extern "C" imt_status imt_non_inclusion_limbs(imt_ctx* ctx, const void* low_leaves, const void* new_vals, size_t b, void* limbs, uint8_t* flags) {
    if (!ctx) return IMT_ERR_INVALID_ARG;
    if (b && (!low_leaves || !new_vals || !limbs)) return fail(ctx, IMT_ERR_INVALID_ARG, "null buffer");
    if (b == 0) return IMT_OK;
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf dl(ctx), dv(ctx), dout(ctx), df(ctx);
    IMT_TRY_CUDA(ctx, dl.alloc(b * 3 * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, dv.alloc(b * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, dout.alloc(b * 6 * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, df.alloc(b * 3));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(dl.p, low_leaves, b * 3 * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(dv.p, new_vals, b * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY(clear_err(ctx));
    k_limb_witness<<<grid_for(b, 256), 256, 0, ctx->stream>>>(dl.as<uint4>(), dv.as<uint4>(), 1, b, ctx->fmt, dout.as<uint4>(), df.as<uint8_t>(),
                                                              ctx->d_err);
    ++ctx->launches;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    IMT_TRY(finish(ctx));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(limbs, dout.p, b * 6 * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    if (flags) IMT_TRY_CUDA(ctx, cudaMemcpyAsync(flags, df.p, b * 3, cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IMT_OK;
}

// ------------------------------------------------------------------------------------------------- sharded inserts
// An insert batch over a subtree-sharded tree (one process per GPU), at most kShardInsertRound inserts per round of calls:
//   1. every rank   imt_shard_insert_neighbors   neighbours of each value in the rank's own sorted index
//   2. all-gather of the five arrays; every rank  imt_shard_insert_plan  -> the replicated plan: write slots x[2b]
//      (GLOBAL), the preimages every write stores (upd), the low leaves before, is_largest
//   3. every rank   imt_shard_insert_apply       its own writes: leaf hashes + versioned update of its subtree, commit,
//                                                index merge; returns the subtree-root version and local path per own write
//   4. all-gather of those (each write has one owner); every rank  imt_shard_insert_cap  the versioned update of the
//      replicated cap over ALL writes -> the global root after every write + the cap part of every path
// The result is the reference's sequence of rebuilds, as for imt_insert_batch (sharding.py assembles the witnesses).
extern "C" imt_status imt_shard_insert_neighbors(imt_tree* t, const void* values, size_t b, void* pred_keys, uint64_t* pred_slots,
                                                 void* succ_keys, uint64_t* succ_slots, uint8_t* flags) {
    if (!t) return IMT_ERR_INVALID_ARG;
    imt_ctx* ctx = t->ctx;
    if (b && (!values || !pred_keys || !pred_slots || !succ_keys || !succ_slots || !flags)) return fail(ctx, IMT_ERR_INVALID_ARG, "null buffer");
    IMT_TRY(ensure_index(t));
    if (b == 0) return IMT_OK;
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf dv(ctx), dpk(ctx), dps(ctx), dsk(ctx), dss(ctx), df(ctx);
    IMT_TRY_CUDA(ctx, dv.alloc(b * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, dpk.alloc(b * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, dsk.alloc(b * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, dps.alloc(b * sizeof(uint64_t)));
    IMT_TRY_CUDA(ctx, dss.alloc(b * sizeof(uint64_t)));
    IMT_TRY_CUDA(ctx, df.alloc(b));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(dv.p, values, b * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY(clear_err(ctx));
    k_ins_neighbors<<<grid_for(b, 256), 256, 0, ctx->stream>>>((const uint4*)t->d_sorted_keys, t->d_sorted_slots, t->occupied,
                                                               (uint64_t)t->rank * t->n, (const uint4*)dv.p, b, ctx->fmt, dpk.as<uint4>(),
                                                               dps.as<uint64_t>(), dsk.as<uint4>(), dss.as<uint64_t>(), df.as<uint8_t>(), ctx->d_err);
    ++ctx->launches;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    IMT_TRY(finish(ctx));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(pred_keys, dpk.p, b * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(succ_keys, dsk.p, b * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(pred_slots, dps.p, b * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(succ_slots, dss.p, b * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(flags, df.p, b, cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IMT_OK;
}

extern "C" imt_status imt_shard_insert_plan(imt_ctx* ctx, const void* values, size_t b, uint64_t first_idx, unsigned world, const void* pred_keys,
                                            const uint64_t* pred_slots, const void* succ_keys, const uint64_t* succ_slots, const uint8_t* flags,
                                            uint64_t* x, void* upd, void* low_old, uint8_t* is_largest) {
    if (!ctx) return IMT_ERR_INVALID_ARG;
    if (b && (!values || !pred_keys || !pred_slots || !succ_keys || !succ_slots || !flags || !x || !upd || !low_old || !is_largest))
        return fail(ctx, IMT_ERR_INVALID_ARG, "null buffer");
    if (b > kShardInsertRound || world == 0) return fail(ctx, IMT_ERR_INVALID_ARG, "at most 4096 inserts per round of sharded insert calls");
    if (b == 0) return IMT_OK;
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t g = (size_t)world * b;
    DevBuf staged(ctx), vals(ctx), dpk(ctx), dps(ctx), dsk(ctx), dss(ctx), df(ctx), dx(ctx), dupd(ctx), dlow(ctx), dlg(ctx);
    IMT_TRY_CUDA(ctx, staged.alloc(b * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, vals.alloc(b * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, dpk.alloc(g * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, dsk.alloc(g * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, dps.alloc(g * sizeof(uint64_t)));
    IMT_TRY_CUDA(ctx, dss.alloc(g * sizeof(uint64_t)));
    IMT_TRY_CUDA(ctx, df.alloc(g));
    IMT_TRY_CUDA(ctx, dx.alloc(2 * b * sizeof(uint64_t)));
    IMT_TRY_CUDA(ctx, dupd.alloc(2 * b * 3 * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, dlow.alloc(b * 3 * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, dlg.alloc(b));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(staged.p, values, b * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(dpk.p, pred_keys, g * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(dsk.p, succ_keys, g * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(dps.p, pred_slots, g * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(dss.p, succ_slots, g * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(df.p, flags, g, cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY(clear_err(ctx));
    IMT_TRY(launch_convert(ctx, staged.p, vals.p, b, ctx->fmt, kFmtCanonical));
    k_ins_plan<<<grid_for(b * 32, 128), 128, 0, ctx->stream>>>(vals.as<uint4>(), b, first_idx, ctx->fmt, world, dpk.as<uint4>(), dps.as<uint64_t>(),
                                                         dsk.as<uint4>(), dss.as<uint64_t>(), df.as<uint8_t>(), dx.as<uint64_t>(), dupd.as<uint4>(),
                                                         dlow.as<uint4>(), dlg.as<uint8_t>(), ctx->d_err);
    ++ctx->launches;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    IMT_TRY(finish(ctx));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(x, dx.p, 2 * b * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(upd, dupd.p, 2 * b * 3 * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(low_old, dlow.p, b * 3 * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(is_largest, dlg.p, b, cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IMT_OK;
}

extern "C" imt_status imt_shard_insert_apply(imt_tree* t, const uint64_t* x, const void* upd, size_t b, void* sub_roots, void* sib_local) {
    if (!t) return IMT_ERR_INVALID_ARG;
    imt_ctx* ctx = t->ctx;
    if (b && (!x || !upd || !sub_roots || !sib_local)) return fail(ctx, IMT_ERR_INVALID_ARG, "null buffer");
    if (b > kShardInsertRound) return fail(ctx, IMT_ERR_INVALID_ARG, "at most 4096 inserts per round of sharded insert calls");
    if (!t->d_pre) return fail(ctx, IMT_ERR_INVALID_ARG, "tree was not built from leaves");
    IMT_TRY(ensure_index(t));
    const unsigned depth = t->depth;
    const size_t writes_all = 2 * b;
    std::memset(sub_roots, 0, writes_all * sizeof(Fr));
    if (depth) std::memset(sib_local, 0, writes_all * depth * sizeof(Fr));
    // ---- this rank's writes, in time order
    const uint64_t base = (uint64_t)t->rank * t->n;
    std::vector<uint64_t> xl;
    std::vector<uint32_t> src;
    for (size_t w = 0; w < writes_all; ++w) {
        if (x[w] >= base && x[w] - base < t->n) xl.push_back(x[w] - base), src.push_back((uint32_t)w);
    }
    const unsigned writes = (unsigned)xl.size();
    if (writes == 0) return IMT_OK;
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<Fr> own_upd((size_t)writes * 3), new_keys;
    std::vector<uint32_t> new_slots;
    const Fr* upd_fe = static_cast<const Fr*>(upd);
    for (unsigned i = 0; i < writes; ++i) {
        for (int k = 0; k < 3; ++k) own_upd[3 * (size_t)i + k] = upd_fe[3 * (size_t)src[i] + k];
        if (src[i] & 1) new_keys.push_back(upd_fe[3 * (size_t)src[i]]), new_slots.push_back((uint32_t)xl[i]);  // odd writes are the new leaves
    }
    const size_t L = depth + 1;
    DevBuf dx(ctx), dupd(ctx), prev(ctx), last(ctx), ver(ctx), pairs(ctx), sib(ctx), top(ctx);
    IMT_TRY_CUDA(ctx, dx.alloc(writes * sizeof(uint64_t)));
    IMT_TRY_CUDA(ctx, dupd.alloc((size_t)writes * 3 * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, prev.alloc(writes * L * sizeof(int)));
    IMT_TRY_CUDA(ctx, last.alloc(writes * L));
    IMT_TRY_CUDA(ctx, ver.alloc(L * writes * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, pairs.alloc(2 * (size_t)writes * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, sib.alloc((size_t)writes * depth * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, top.alloc(writes * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(dx.p, xl.data(), writes * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(dupd.p, own_upd.data(), (size_t)writes * 3 * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY(clear_err(ctx));
    IMT_TRY(launch_hash(ctx, 3, dupd.p, ver.p, writes, ctx->fmt, kFmtMontgomery, ctx->stream));
    IMT_TRY(versioned_levels(ctx, t->d_levels, t->n, depth, dx.as<uint64_t>(), writes, ver.as<Fr>(), prev.as<int>(), last.as<uint8_t>(),
                             pairs.as<Fr>(), nullptr, nullptr, sib.as<uint4>()));
    k_export_versions<<<grid_for(writes, 256), 256, 0, ctx->stream>>>(ver.as<uint4>() + 2 * ((size_t)depth * writes), writes, ctx->fmt, top.as<uint4>());
    k_ins_commit<<<grid_for((size_t)writes * L, 256), 256, 0, ctx->stream>>>(ver.as<uint4>(), dupd.as<uint4>(), dx.as<uint64_t>(), last.as<uint8_t>(),
                                                                            writes, depth, t->n, (uint4*)t->d_levels, (uint4*)t->d_pre);
    ctx->launches += 2;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    std::vector<Fr> h_top(writes), h_sib((size_t)writes * depth);
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(h_top.data(), top.p, writes * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    if (depth) IMT_TRY_CUDA(ctx, cudaMemcpyAsync(h_sib.data(), sib.p, (size_t)writes * depth * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY(finish(ctx));
    Fr* out_roots = static_cast<Fr*>(sub_roots);
    Fr* out_sib = static_cast<Fr*>(sib_local);
    for (unsigned i = 0; i < writes; ++i) {
        out_roots[src[i]] = h_top[i];
        for (unsigned l = 0; l < depth; ++l) out_sib[(size_t)src[i] * depth + l] = h_sib[(size_t)i * depth + l];
    }
    // ---- the new leaves of this rank join its sorted index
    const size_t nk = new_keys.size();
    if (nk) {
        DevBuf staged(ctx), keys(ctx), slots(ctx);
        IMT_TRY_CUDA(ctx, staged.alloc(nk * sizeof(Fr)));
        IMT_TRY_CUDA(ctx, keys.alloc(nk * sizeof(Fr)));
        IMT_TRY_CUDA(ctx, slots.alloc(nk * sizeof(uint32_t)));
        if (!t->d_alt_keys) IMT_TRY_CUDA(ctx, tree_malloc(ctx, (void**)&t->d_alt_keys, t->index_capacity * sizeof(Fr)));
        if (!t->d_alt_slots) IMT_TRY_CUDA(ctx, tree_malloc(ctx, (void**)&t->d_alt_slots, t->index_capacity * sizeof(uint32_t)));
        IMT_TRY_CUDA(ctx, cudaMemcpyAsync(staged.p, new_keys.data(), nk * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
        IMT_TRY_CUDA(ctx, cudaMemcpyAsync(slots.p, new_slots.data(), nk * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        IMT_TRY(launch_convert(ctx, staged.p, keys.p, nk, ctx->fmt, kFmtCanonical));
        IMT_TRY(sort_pairs(ctx, keys.as<Fr>(), slots.as<uint32_t>(), nk));
        k_merge_rank<<<grid_for(t->occupied + nk, 256), 256, 0, ctx->stream>>>((const uint4*)t->d_sorted_keys, t->d_sorted_slots, t->occupied,
                                                                              keys.as<uint4>(), slots.as<uint32_t>(), nk, (uint4*)t->d_alt_keys,
                                                                              t->d_alt_slots);
        ++ctx->launches;
        IMT_TRY_CUDA(ctx, cudaGetLastError());
        IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        std::swap(t->d_sorted_keys, t->d_alt_keys);
        std::swap(t->d_sorted_slots, t->d_alt_slots);
        t->prefix_valid = false;
        t->occupied += nk;
    }
    if (t->rank == 0) t->head_next_zero = false;
    return IMT_OK;
}

extern "C" imt_status imt_shard_insert_cap(imt_tree* t, const uint64_t* x, const void* sub_roots, size_t b, void* roots, void* sib_cap) {
    if (!t) return IMT_ERR_INVALID_ARG;
    imt_ctx* ctx = t->ctx;
    if (b && (!x || !sub_roots || !roots)) return fail(ctx, IMT_ERR_INVALID_ARG, "null buffer");
    if (b > kShardInsertRound) return fail(ctx, IMT_ERR_INVALID_ARG, "at most 4096 inserts per round of sharded insert calls");
    if (!t->cap_valid) return fail(ctx, IMT_ERR_INVALID_ARG, "no cap attached: exchange the subtree roots first");
    if (b == 0) return IMT_OK;
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    const unsigned writes = (unsigned)(2 * b), depth = t->cap_depth;
    const size_t L = depth + 1;
    std::vector<uint64_t> owner(writes);
    for (unsigned w = 0; w < writes; ++w) {
        owner[w] = x[w] / t->n;
        if (owner[w] >= t->world) return fail(ctx, IMT_ERR_INDEX_OOB, "index out of bounds");
    }
    DevBuf dx(ctx), staged(ctx), prev(ctx), last(ctx), ver(ctx), pairs(ctx), sib(ctx), top(ctx);
    IMT_TRY_CUDA(ctx, dx.alloc(writes * sizeof(uint64_t)));
    IMT_TRY_CUDA(ctx, staged.alloc(writes * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, prev.alloc(writes * L * sizeof(int)));
    IMT_TRY_CUDA(ctx, last.alloc(writes * L));
    IMT_TRY_CUDA(ctx, ver.alloc(L * writes * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, pairs.alloc(2 * (size_t)writes * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, sib.alloc((size_t)writes * (depth ? depth : 1) * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, top.alloc(writes * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(dx.p, owner.data(), writes * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(staged.p, sub_roots, writes * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY(clear_err(ctx));
    IMT_TRY(launch_convert(ctx, staged.p, ver.p, writes, ctx->fmt, kFmtMontgomery));  // version 0 of write t: its subtree's root after it
    IMT_TRY(versioned_levels(ctx, t->d_cap, t->world, depth, dx.as<uint64_t>(), writes, ver.as<Fr>(), prev.as<int>(), last.as<uint8_t>(),
                             pairs.as<Fr>(), nullptr, nullptr, sib_cap ? sib.as<uint4>() : nullptr));
    k_export_versions<<<grid_for(writes, 256), 256, 0, ctx->stream>>>(ver.as<uint4>() + 2 * ((size_t)depth * writes), writes, ctx->fmt, top.as<uint4>());
    k_ins_commit<<<grid_for((size_t)writes * L, 256), 256, 0, ctx->stream>>>(ver.as<uint4>(), nullptr, dx.as<uint64_t>(), last.as<uint8_t>(), writes,
                                                                            depth, t->world, (uint4*)t->d_cap, nullptr);
    ctx->launches += 2;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    IMT_TRY(finish(ctx));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(roots, top.p, writes * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    if (sib_cap && depth) IMT_TRY_CUDA(ctx, cudaMemcpyAsync(sib_cap, sib.p, (size_t)writes * depth * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IMT_OK;
}

// ------------------------------------------------------------------------------------------------- sharded calls (NCCL inside)
// The same sharded algorithms the piecewise calls above expose to a caller-driven exchange, with the exchange done here through
// the group's transport (imt_comm.cu). `Parts` = the shards of one tree this process drives: one in process-per-GPU mode, all
// of them in single-process mode. Per-rank phases run on one host thread per local shard (a context is single-threaded, two
// contexts are independent); collectives are issued by the calling thread between the phases.
#include <thread>

namespace {

struct Parts {
    imt_group* g = nullptr;
    std::vector<imt_tree*> t;  // local shards, ascending rank
    size_t n_total() const { return t[0]->n * (size_t)group_world(g); }
};

imt_status parts_of_tree(imt_tree* tree, Parts* p) {
    if (!tree) return IMT_ERR_INVALID_ARG;
    imt_ctx* ctx = tree->ctx;
    if (!ctx->group) return fail(ctx, IMT_ERR_INVALID_ARG, "the context has no communicator: call imt_comm_create first");
    if (group_local_count(ctx->group) != 1) return fail(ctx, IMT_ERR_INVALID_ARG, "this tree is a shard of an imt_mtree: use the imt_mtree_* calls");
    if (!tree->cap_valid) return fail(ctx, IMT_ERR_INVALID_ARG, "no cap attached: exchange the subtree roots first");
    p->g = ctx->group;
    p->t = {tree};
    return IMT_OK;
}
imt_status parts_of_mtree(imt_mtree* mt, Parts* p) {
    IMT_TRY(mtree_parts(mt, &p->g, &p->t));
    for (imt_tree* t : p->t)
        if (!t->cap_valid) return fail(t->ctx, IMT_ERR_INVALID_ARG, "no cap attached");
    return IMT_OK;
}

// fn(slot) on every local shard, concurrently when there are several; returns the first failure
template <class Fn>
imt_status for_each_part(const Parts& p, Fn fn) {
    const size_t n = p.t.size();
    if (n == 1) return fn((size_t)0);
    std::vector<imt_status> st(n, IMT_OK);
    std::vector<std::thread> th;
    th.reserve(n);
    for (size_t i = 0; i < n; ++i) th.emplace_back([&, i] { st[i] = fn(i); });
    for (auto& x : th) x.join();
    for (size_t i = 0; i < n; ++i)
        if (st[i] != IMT_OK) {
            if (i) p.t[0]->ctx->last_error = p.t[i]->ctx->last_error;
            return st[i];
        }
    return IMT_OK;
}

// ---- low-leaf lookups: per-rank candidates packed as [header 32 B][keys q x 32][slots q x 8][flags q], one all-gather, merge
constexpr size_t kCandHeader = 32;
size_t cand_bytes(size_t q) { return (kCandHeader + q * 41 + 15) & ~(size_t)15; }
__global__ void k_cand_header(unsigned long long* hdr, unsigned long long occupied, unsigned long long head_next_zero) {
    hdr[0] = occupied, hdr[1] = head_next_zero, hdr[2] = 0, hdr[3] = 0;
}
__global__ void __launch_bounds__(256) k_low_leaf_merge_packed(const uint8_t* __restrict__ gathered, size_t stride, unsigned world, size_t q,
                                                               uint64_t n_total, const uint4* __restrict__ values,
                                                               uint64_t* __restrict__ low_idx, uint8_t* __restrict__ matched,
                                                               unsigned long long* __restrict__ occupied_out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    uint64_t occupied = 0;
    for (unsigned r = 0; r < world; ++r) occupied += reinterpret_cast<const unsigned long long*>(gathered + (size_t)r * stride)[0];
    const bool head_next_zero = reinterpret_cast<const unsigned long long*>(gathered)[1] != 0;  // rank 0 owns slot 0
    if (i == 0 && occupied_out) *occupied_out = occupied;
    if (i >= q) return;
    uint32_t best[8], v[8];
    bool have = false, present = false;
    uint64_t slot = 0;
    for (unsigned r = 0; r < world; ++r) {
        const uint8_t* blk = gathered + (size_t)r * stride + kCandHeader;
        const uint8_t f = blk[q * 40 + i];
        present |= (f & 2) != 0;
        if (f & 1) {
            uint32_t k[8];
            load_fe(k, reinterpret_cast<const uint4*>(blk) + 2 * i);
            if (!have || cmp256(k, best) > 0) copy256(best, k), slot = reinterpret_cast<const uint64_t*>(blk + q * 32)[i], have = true;
        }
    }
    load_fe(v, values + 2 * i);
    uint64_t low = 0;
    uint8_t hit = 0;
    if (head_next_zero) hit = 1;
    else if (have && !present) low = slot, hit = 1;
    else if (!zero256(v) && occupied < n_total) low = occupied, hit = 1;
    low_idx[i] = low;
    if (matched) matched[i] = hit;
}

// Device buffers of one local shard for a batch of q replicated queries
struct PartBufs {
    DevBuf values, low, matched, pack, gathered, occupied;
    explicit PartBufs(imt_ctx* c) : values(c), low(c), matched(c), pack(c), gathered(c), occupied(c) {}
};

// every local shard: values already on the device -> d_low / d_matched (replicated). occupied_total (optional) on the host.
imt_status sharded_lookup_dev(const Parts& p, std::vector<PartBufs*>& b, size_t q, uint64_t* occupied_total) {
    const unsigned world = group_world(p.g);
    const size_t stride = cand_bytes(q);
    IMT_TRY(for_each_part(p, [&](size_t i) -> imt_status {
        imt_tree* t = p.t[i];
        imt_ctx* ctx = t->ctx;
        IMT_TRY(ensure_index(t));
        IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
        IMT_TRY_CUDA(ctx, b[i]->pack.alloc(stride));
        IMT_TRY_CUDA(ctx, b[i]->gathered.alloc(stride * world));
        IMT_TRY_CUDA(ctx, b[i]->low.alloc(q * sizeof(uint64_t)));
        IMT_TRY_CUDA(ctx, b[i]->matched.alloc(q));
        IMT_TRY_CUDA(ctx, b[i]->occupied.alloc(sizeof(unsigned long long)));
        IMT_TRY(clear_err(ctx));
        uint8_t* pk = b[i]->pack.as<uint8_t>();
        k_cand_header<<<1, 1, 0, ctx->stream>>>((unsigned long long*)pk, t->occupied, (t->rank == 0 && t->head_next_zero) ? 1ull : 0ull);
        if (q)
            k_low_leaf_candidates<<<grid_for(q, 256), 256, 0, ctx->stream>>>((const uint4*)t->d_sorted_keys, t->d_sorted_slots, t->occupied,
                                                                             (uint64_t)t->rank * t->n, b[i]->values.as<uint4>(), q, ctx->fmt,
                                                                             (uint4*)(pk + kCandHeader), (uint64_t*)(pk + kCandHeader + q * 32),
                                                                             pk + kCandHeader + q * 40, ctx->d_err);
        ctx->launches += 2;
        IMT_TRY_CUDA(ctx, cudaGetLastError());
        return IMT_OK;
    }));
    std::vector<const void*> send(p.t.size());
    std::vector<void*> recv(p.t.size());
    for (size_t i = 0; i < p.t.size(); ++i) send[i] = b[i]->pack.p, recv[i] = b[i]->gathered.p;
    IMT_TRY(group_all_gather(p.g, send, recv, stride));
    return for_each_part(p, [&](size_t i) -> imt_status {
        imt_tree* t = p.t[i];
        imt_ctx* ctx = t->ctx;
        IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
        k_low_leaf_merge_packed<<<grid_for(q ? q : 1, 256), 256, 0, ctx->stream>>>(b[i]->gathered.as<uint8_t>(), stride, world, q, p.n_total(),
                                                                                   b[i]->values.as<uint4>(), b[i]->low.as<uint64_t>(),
                                                                                   b[i]->matched.as<uint8_t>(), b[i]->occupied.as<unsigned long long>());
        ++ctx->launches;
        IMT_TRY_CUDA(ctx, cudaGetLastError());
        if (i == 0 && occupied_total) {
            unsigned long long occ = 0;
            IMT_TRY_CUDA(ctx, cudaMemcpyAsync(&occ, b[i]->occupied.p, sizeof(occ), cudaMemcpyDeviceToHost, ctx->stream));
            IMT_TRY(finish(ctx));
            *occupied_total = occ;
            return IMT_OK;
        }
        return finish(ctx);
    });
}

// every local shard: d_idx (replicated, GLOBAL) -> the owner's rows, zeros elsewhere -> sum over ranks: replicated paths / leaves
imt_status sharded_gather_dev(const Parts& p, const std::vector<const uint64_t*>& d_idx, size_t q, const std::vector<void*>& d_sib,
                              const std::vector<uint8_t*>& d_hel, const std::vector<void*>& d_leaves, const std::vector<uint8_t*>& d_largest) {
    const unsigned depth = p.t[0]->depth + p.t[0]->cap_depth;
    const bool paths = !d_sib.empty() && depth, hel = !d_hel.empty() && depth, leaves = !d_leaves.empty(), largest = !d_largest.empty();
    IMT_TRY(for_each_part(p, [&](size_t i) -> imt_status {
        imt_tree* t = p.t[i];
        imt_ctx* ctx = t->ctx;
        IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
        IMT_TRY(clear_err(ctx));
        if (paths) IMT_TRY(launch_gather_proofs(t, d_idx[i], q, d_sib[i], hel ? d_hel[i] : nullptr, nullptr, true));
        if (leaves || largest) {
            if (!t->d_pre) return fail(ctx, IMT_ERR_INVALID_ARG, "tree was not built from leaves");
            k_gather_leaves<<<grid_for(q, 256), 256, 0, ctx->stream>>>((const uint4*)t->d_pre, t->n, (uint64_t)t->rank * t->n, d_idx[i], q,
                                                                       leaves ? (uint4*)d_leaves[i] : nullptr, largest ? d_largest[i] : nullptr,
                                                                       ctx->d_err, p.n_total());
            ++ctx->launches;
            IMT_TRY_CUDA(ctx, cudaGetLastError());
        }
        return finish(ctx);  // index errors surface before the collective (every rank sees the same indices: all fail alike)
    }));
    auto reduce = [&](auto& bufs, size_t count, bool bytes8) -> imt_status {
        std::vector<void*> v(bufs.size());
        for (size_t i = 0; i < bufs.size(); ++i) v[i] = (void*)bufs[i];
        return group_all_reduce_sum(p.g, v, count, bytes8);
    };
    if (paths) IMT_TRY(reduce(d_sib, q * depth * 4, false));
    if (paths && hel) IMT_TRY(reduce(d_hel, q * depth, true));
    if (leaves) IMT_TRY(reduce(d_leaves, q * 12, false));
    if (largest) IMT_TRY(reduce(d_largest, q, true));
    return IMT_OK;
}

imt_status sync_all(const Parts& p) {
    for (imt_tree* t : p.t) {
        IMT_TRY_CUDA(t->ctx, cudaSetDevice(t->ctx->device));
        IMT_TRY_CUDA(t->ctx, cudaStreamSynchronize(t->ctx->stream));
    }
    return IMT_OK;
}

// ---- host-facing: lookups / non-inclusion witnesses (replicated inputs; outputs read from the first local shard)
imt_status sharded_non_inclusion(const Parts& p, const void* values, size_t q, uint64_t* low_idx, uint8_t* matched, void* low_leaves,
                                 void* siblings, uint8_t* helpers, uint8_t* is_largest, uint64_t* occupied_total) {
    imt_ctx* c0 = p.t[0]->ctx;
    if (q && !values) return fail(c0, IMT_ERR_INVALID_ARG, "null buffer");
    const unsigned depth = p.t[0]->depth + p.t[0]->cap_depth;
    const size_t nl = p.t.size();
    std::vector<PartBufs*> b(nl, nullptr);
    std::vector<DevBuf*> dlv(nl, nullptr), dsib(nl, nullptr), dhel(nl, nullptr), dlg(nl, nullptr);
    auto cleanup = [&] {
        for (size_t i = 0; i < nl; ++i) {
            cudaSetDevice(p.t[i]->ctx->device);
            delete b[i], delete dlv[i], delete dsib[i], delete dhel[i], delete dlg[i];
        }
    };
    imt_status st = IMT_OK;
    for (size_t i = 0; i < nl && st == IMT_OK; ++i) {
        imt_ctx* ctx = p.t[i]->ctx;
        cudaSetDevice(ctx->device);
        b[i] = new PartBufs(ctx);
        dlv[i] = new DevBuf(ctx), dsib[i] = new DevBuf(ctx), dhel[i] = new DevBuf(ctx), dlg[i] = new DevBuf(ctx);
        cudaError_t e = b[i]->values.alloc(q * sizeof(Fr));
        if (e == cudaSuccess && q) e = cudaMemcpyAsync(b[i]->values.p, values, q * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess && low_leaves) e = dlv[i]->alloc(q * 3 * sizeof(Fr));
        if (e == cudaSuccess && (siblings || helpers)) e = dsib[i]->alloc(q * (size_t)depth * sizeof(Fr));
        if (e == cudaSuccess && helpers) e = dhel[i]->alloc(q * (size_t)depth);
        if (e == cudaSuccess && is_largest) e = dlg[i]->alloc(q);
        if (e != cudaSuccess) st = fail(ctx, IMT_ERR_CUDA, cudaGetErrorString(e));
    }
    if (st == IMT_OK) st = sharded_lookup_dev(p, b, q, occupied_total);
    if (st == IMT_OK && q && (low_leaves || siblings || helpers || is_largest)) {
        std::vector<const uint64_t*> idx(nl);
        std::vector<void*> vs, vl;
        std::vector<uint8_t*> vh, vg;
        for (size_t i = 0; i < nl; ++i) {
            idx[i] = b[i]->low.as<uint64_t>();
            if (siblings || helpers) vs.push_back(dsib[i]->p);
            if (helpers) vh.push_back(dhel[i]->as<uint8_t>());
            if (low_leaves) vl.push_back(dlv[i]->p);
            if (is_largest) vg.push_back(dlg[i]->as<uint8_t>());
        }
        st = sharded_gather_dev(p, idx, q, vs, vh, vl, vg);
    }
    if (st == IMT_OK && q) {
        imt_ctx* ctx = c0;
        cudaSetDevice(ctx->device);
        cudaError_t e = cudaSuccess;
        auto d2h = [&](void* host, const void* dev, size_t bytes) {
            if (host && bytes && e == cudaSuccess) e = cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        };
        d2h(low_idx, b[0]->low.p, q * sizeof(uint64_t));
        d2h(matched, b[0]->matched.p, q);
        d2h(low_leaves, dlv[0]->p, q * 3 * sizeof(Fr));
        d2h(siblings, dsib[0]->p, q * (size_t)depth * sizeof(Fr));
        d2h(helpers, dhel[0]->p, q * (size_t)depth);
        d2h(is_largest, dlg[0]->p, q);
        if (e != cudaSuccess) st = fail(ctx, IMT_ERR_CUDA, cudaGetErrorString(e));
    }
    const imt_status s2 = sync_all(p);
    cleanup();
    return st != IMT_OK ? st : s2;
}

imt_status sharded_gather_host(const Parts& p, const uint64_t* indices, size_t q, void* siblings, uint8_t* helpers, void* leaves,
                               uint8_t* is_largest) {
    imt_ctx* c0 = p.t[0]->ctx;
    if (q && !indices) return fail(c0, IMT_ERR_INVALID_ARG, "null buffer");
    if (q == 0) return IMT_OK;
    const unsigned depth = p.t[0]->depth + p.t[0]->cap_depth;
    const size_t nl = p.t.size();
    std::vector<DevBuf*> di(nl, nullptr), dlv(nl, nullptr), dsib(nl, nullptr), dhel(nl, nullptr), dlg(nl, nullptr);
    imt_status st = IMT_OK;
    for (size_t i = 0; i < nl && st == IMT_OK; ++i) {
        imt_ctx* ctx = p.t[i]->ctx;
        cudaSetDevice(ctx->device);
        di[i] = new DevBuf(ctx), dlv[i] = new DevBuf(ctx), dsib[i] = new DevBuf(ctx), dhel[i] = new DevBuf(ctx), dlg[i] = new DevBuf(ctx);
        cudaError_t e = di[i]->alloc(q * sizeof(uint64_t));
        if (e == cudaSuccess) e = cudaMemcpyAsync(di[i]->p, indices, q * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess && leaves) e = dlv[i]->alloc(q * 3 * sizeof(Fr));
        if (e == cudaSuccess && (siblings || helpers)) e = dsib[i]->alloc(q * (size_t)depth * sizeof(Fr));
        if (e == cudaSuccess && helpers) e = dhel[i]->alloc(q * (size_t)depth);
        if (e == cudaSuccess && is_largest) e = dlg[i]->alloc(q);
        if (e != cudaSuccess) st = fail(ctx, IMT_ERR_CUDA, cudaGetErrorString(e));
    }
    if (st == IMT_OK) {
        std::vector<const uint64_t*> idx(nl);
        std::vector<void*> vs, vl;
        std::vector<uint8_t*> vh, vg;
        for (size_t i = 0; i < nl; ++i) {
            idx[i] = di[i]->as<uint64_t>();
            if (siblings || helpers) vs.push_back(dsib[i]->p);
            if (helpers) vh.push_back(dhel[i]->as<uint8_t>());
            if (leaves) vl.push_back(dlv[i]->p);
            if (is_largest) vg.push_back(dlg[i]->as<uint8_t>());
        }
        st = sharded_gather_dev(p, idx, q, vs, vh, vl, vg);
    }
    if (st == IMT_OK) {
        cudaSetDevice(c0->device);
        cudaError_t e = cudaSuccess;
        auto d2h = [&](void* host, const void* dev, size_t bytes) {
            if (host && bytes && e == cudaSuccess) e = cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c0->stream);
        };
        d2h(siblings, dsib[0]->p, q * (size_t)depth * sizeof(Fr));
        d2h(helpers, dhel[0]->p, q * (size_t)depth);
        d2h(leaves, dlv[0]->p, q * 3 * sizeof(Fr));
        d2h(is_largest, dlg[0]->p, q);
        if (e != cudaSuccess) st = fail(c0, IMT_ERR_CUDA, cudaGetErrorString(e));
    }
    const imt_status s2 = sync_all(p);
    for (size_t i = 0; i < nl; ++i) {
        cudaSetDevice(p.t[i]->ctx->device);
        delete di[i], delete dlv[i], delete dsib[i], delete dhel[i], delete dlg[i];
    }
    return st != IMT_OK ? st : s2;
}

// ---- host-staged collectives for the (small) plan arrays of an insert round
imt_status host_all_gather(const Parts& p, const std::vector<const void*>& h_send, size_t bytes, std::vector<uint8_t>* h_recv) {
    const size_t nl = p.t.size();
    const unsigned world = group_world(p.g);
    std::vector<DevBuf*> s(nl, nullptr), r(nl, nullptr);
    std::vector<const void*> send(nl);
    std::vector<void*> recv(nl);
    imt_status st = IMT_OK;
    for (size_t i = 0; i < nl && st == IMT_OK; ++i) {
        imt_ctx* ctx = p.t[i]->ctx;
        cudaSetDevice(ctx->device);
        s[i] = new DevBuf(ctx), r[i] = new DevBuf(ctx);
        cudaError_t e = s[i]->alloc(bytes);
        if (e == cudaSuccess) e = r[i]->alloc(bytes * world);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s[i]->p, h_send[i], bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) st = fail(ctx, IMT_ERR_CUDA, cudaGetErrorString(e));
        send[i] = s[i]->p, recv[i] = r[i]->p;
    }
    if (st == IMT_OK) st = group_all_gather(p.g, send, recv, bytes);
    if (st == IMT_OK) {
        imt_ctx* ctx = p.t[0]->ctx;
        cudaSetDevice(ctx->device);
        h_recv->resize(bytes * world);
        if (cudaMemcpyAsync(h_recv->data(), r[0]->p, bytes * world, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
            st = fail(ctx, IMT_ERR_CUDA, "all-gather read back failed");
    }
    const imt_status s2 = sync_all(p);
    for (size_t i = 0; i < nl; ++i) {
        cudaSetDevice(p.t[i]->ctx->device);
        delete s[i], delete r[i];
    }
    return st != IMT_OK ? st : s2;
}
// in place on the host arrays of every local shard: element-wise sum over ranks (one owner per element); result in h[0]
imt_status host_all_reduce_u64(const Parts& p, const std::vector<void*>& h, size_t count) {
    const size_t nl = p.t.size(), bytes = count * 8;
    std::vector<DevBuf*> d(nl, nullptr);
    std::vector<void*> v(nl);
    imt_status st = IMT_OK;
    for (size_t i = 0; i < nl && st == IMT_OK; ++i) {
        imt_ctx* ctx = p.t[i]->ctx;
        cudaSetDevice(ctx->device);
        d[i] = new DevBuf(ctx);
        cudaError_t e = d[i]->alloc(bytes);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d[i]->p, h[i], bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) st = fail(ctx, IMT_ERR_CUDA, cudaGetErrorString(e));
        v[i] = d[i]->p;
    }
    if (st == IMT_OK) st = group_all_reduce_sum(p.g, v, count, false);
    if (st == IMT_OK) {
        imt_ctx* ctx = p.t[0]->ctx;
        cudaSetDevice(ctx->device);
        if (cudaMemcpyAsync(h[0], d[0]->p, bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) st = fail(ctx, IMT_ERR_CUDA, "all-reduce read back failed");
    }
    const imt_status s2 = sync_all(p);
    for (size_t i = 0; i < nl; ++i) {
        cudaSetDevice(p.t[i]->ctx->device);
        delete d[i];
    }
    return st != IMT_OK ? st : s2;
}

// ---- inserts: the rounds of imt_shard_insert_* with the exchanges done here
imt_status sharded_insert(const Parts& p, const void* new_vals, size_t b_total, uint64_t first_idx, imt_insert_witness* w) {
    imt_ctx* c0 = p.t[0]->ctx;
    if (b_total && !new_vals) return fail(c0, IMT_ERR_INVALID_ARG, "null buffer");
    for (imt_tree* t : p.t)
        if (!t->d_pre) return fail(c0, IMT_ERR_INVALID_ARG, "tree was not built from leaves");
    const unsigned world = group_world(p.g), d_local = p.t[0]->depth, d_cap = p.t[0]->cap_depth, depth = d_local + d_cap;
    const size_t nl = p.t.size(), n_total = p.n_total();
    // total occupancy: the lookup machinery with zero queries carries the per-rank counts
    uint64_t occupied = 0;
    {
        std::vector<PartBufs*> pb(nl, nullptr);
        imt_status st = IMT_OK;
        for (size_t i = 0; i < nl && st == IMT_OK; ++i) {
            cudaSetDevice(p.t[i]->ctx->device);
            pb[i] = new PartBufs(p.t[i]->ctx);
            if (pb[i]->values.alloc(16) != cudaSuccess) st = fail(p.t[i]->ctx, IMT_ERR_CUDA, "alloc");
        }
        if (st == IMT_OK) st = sharded_lookup_dev(p, pb, 0, &occupied);
        const imt_status s2 = sync_all(p);
        for (size_t i = 0; i < nl; ++i) {
            cudaSetDevice(p.t[i]->ctx->device);
            delete pb[i];
        }
        IMT_TRY(st);
        IMT_TRY(s2);
    }
    if (first_idx != occupied) return fail(c0, IMT_ERR_INVALID_ARG, "first_idx must be the next free slot (the total number of occupied slots)");
    if (b_total > n_total - occupied) return fail(c0, IMT_ERR_TREE_FULL, imt_status_string(IMT_ERR_TREE_FULL));
    const imt_insert_witness none = {};
    const imt_insert_witness out = w ? *w : none;
    const Fr* vals_all = static_cast<const Fr*>(new_vals);
    for (size_t off = 0; off < b_total; off += kShardInsertRound) {
        const size_t b = std::min(kShardInsertRound, b_total - off), W = 2 * b;
        const Fr* vals = vals_all + off;
        Fr root_before;
        IMT_TRY(imt_tree_root(p.t[0], &root_before));
        // 1. neighbours of every value in each rank's own index, packed [pred keys][succ keys][pred slots][succ slots][flags]
        const size_t nb = b * (32 + 32 + 8 + 8 + 1);
        std::vector<std::vector<uint8_t>> packs(nl, std::vector<uint8_t>(nb));
        IMT_TRY(for_each_part(p, [&](size_t i) -> imt_status {
            uint8_t* pk = packs[i].data();
            return imt_shard_insert_neighbors(p.t[i], vals, b, pk, (uint64_t*)(pk + b * 64), pk + b * 32, (uint64_t*)(pk + b * 72), pk + b * 80);
        }));
        std::vector<const void*> hs(nl);
        for (size_t i = 0; i < nl; ++i) hs[i] = packs[i].data();
        std::vector<uint8_t> gathered;
        IMT_TRY(host_all_gather(p, hs, nb, &gathered));
        // 2. the replicated plan (identical on every rank: computed once per process)
        std::vector<Fr> pk_all((size_t)world * b), sk_all((size_t)world * b);
        std::vector<uint64_t> ps_all((size_t)world * b), ss_all((size_t)world * b);
        std::vector<uint8_t> fl_all((size_t)world * b);
        for (unsigned r = 0; r < world; ++r) {
            const uint8_t* blk = gathered.data() + (size_t)r * nb;
            std::memcpy(&pk_all[(size_t)r * b], blk, b * 32);
            std::memcpy(&sk_all[(size_t)r * b], blk + b * 32, b * 32);
            std::memcpy(&ps_all[(size_t)r * b], blk + b * 64, b * 8);
            std::memcpy(&ss_all[(size_t)r * b], blk + b * 72, b * 8);
            std::memcpy(&fl_all[(size_t)r * b], blk + b * 80, b);
        }
        std::vector<uint64_t> x(W);
        std::vector<Fr> upd(W * 3), low_old(b * 3);
        std::vector<uint8_t> largest(b);
        IMT_TRY(imt_shard_insert_plan(c0, vals, b, first_idx + off, world, pk_all.data(), ps_all.data(), sk_all.data(), ss_all.data(), fl_all.data(),
                                      x.data(), upd.data(), low_old.data(), largest.data()));
        // 3. every rank applies its own writes to its subtree
        std::vector<std::vector<Fr>> sub(nl, std::vector<Fr>(W)), sibl(nl, std::vector<Fr>(W * (size_t)(d_local ? d_local : 1)));
        IMT_TRY(for_each_part(p, [&](size_t i) -> imt_status {
            return imt_shard_insert_apply(p.t[i], x.data(), upd.data(), b, sub[i].data(), sibl[i].data());
        }));
        // 4. each write has one owner: a sum over ranks assembles the subtree-root versions and the local paths
        {
            std::vector<void*> v(nl);
            for (size_t i = 0; i < nl; ++i) v[i] = sub[i].data();
            IMT_TRY(host_all_reduce_u64(p, v, W * 4));
            if (d_local) {
                for (size_t i = 0; i < nl; ++i) v[i] = sibl[i].data();
                IMT_TRY(host_all_reduce_u64(p, v, W * (size_t)d_local * 4));
            }
        }
        // 5. all writes on the replicated cap
        std::vector<std::vector<Fr>> roots(nl, std::vector<Fr>(W)), sibc(nl, std::vector<Fr>(W * (size_t)(d_cap ? d_cap : 1)));
        IMT_TRY(for_each_part(p, [&](size_t i) -> imt_status {
            return imt_shard_insert_cap(p.t[i], x.data(), sub[0].data(), b, roots[i].data(), d_cap ? sibc[i].data() : nullptr);
        }));
        // 6. witnesses of this round (IMT:444-489), from the first local shard's replicated copies
        const Fr* rt = roots[0].data();
        for (size_t k = 0; k < b; ++k) {
            const size_t o = off + k;
            if (out.old_roots) static_cast<Fr*>(out.old_roots)[o] = k ? rt[2 * k - 1] : root_before;
            if (out.new_roots) static_cast<Fr*>(out.new_roots)[o] = rt[2 * k + 1];
            if (out.low_idx) out.low_idx[o] = x[2 * k];
            if (out.is_largest) out.is_largest[o] = largest[k];
            if (out.low_leaves) std::memcpy(static_cast<Fr*>(out.low_leaves) + 3 * o, &low_old[3 * k], 3 * sizeof(Fr));
            if (out.new_leaves) std::memcpy(static_cast<Fr*>(out.new_leaves) + 3 * o, &upd[3 * (2 * k + 1)], 3 * sizeof(Fr));
            for (int side = 0; side < 2; ++side) {
                const size_t t = 2 * k + side;
                Fr* sib_out = static_cast<Fr*>(side ? out.new_siblings : out.low_siblings);
                uint8_t* hel_out = side ? out.new_helpers : out.low_helpers;
                if (sib_out) {
                    if (d_local) std::memcpy(sib_out + o * depth, &sibl[0][t * d_local], d_local * sizeof(Fr));
                    if (d_cap) std::memcpy(sib_out + o * depth + d_local, &sibc[0][t * d_cap], d_cap * sizeof(Fr));
                }
                if (hel_out)
                    for (unsigned l = 0; l < depth; ++l) hel_out[o * depth + l] = ((x[t] >> l) & 1) == 0;  // 1 = current node is LEFT (utils.rs:70)
            }
        }
    }
    return IMT_OK;
}

// owner-sharded witness traces: the positions (into `indices`) of the queries whose leaves shard t owns
std::vector<uint64_t> owned_positions(const imt_tree* t, const uint64_t* indices, size_t q, uint64_t n_total, bool* oob) {
    std::vector<uint64_t> pos;
    const uint64_t base = (uint64_t)t->rank * t->n;
    for (size_t i = 0; i < q; ++i) {
        if (indices[i] >= n_total) *oob = true;
        else if (indices[i] >= base && indices[i] - base < t->n) pos.push_back(i);
    }
    return pos;
}

}  // namespace

// ---- one process per GPU
extern "C" imt_status imt_sharded_get_proofs(imt_tree* tree, const uint64_t* indices, size_t q, void* siblings, uint8_t* helpers) {
    Parts p;
    IMT_TRY(parts_of_tree(tree, &p));
    if (q && !siblings) return fail(tree->ctx, IMT_ERR_INVALID_ARG, "null buffer");
    return sharded_gather_host(p, indices, q, siblings, helpers, nullptr, nullptr);
}
extern "C" imt_status imt_sharded_leaves(imt_tree* tree, const uint64_t* indices, size_t q, void* leaves, uint8_t* is_largest) {
    Parts p;
    IMT_TRY(parts_of_tree(tree, &p));
    return sharded_gather_host(p, indices, q, nullptr, nullptr, leaves, is_largest);
}
extern "C" imt_status imt_sharded_low_leaf_lookup(imt_tree* tree, const void* values, size_t q, uint64_t* low_idx, uint8_t* matched) {
    Parts p;
    IMT_TRY(parts_of_tree(tree, &p));
    if (q && !low_idx) return fail(tree->ctx, IMT_ERR_INVALID_ARG, "null buffer");
    return sharded_non_inclusion(p, values, q, low_idx, matched, nullptr, nullptr, nullptr, nullptr, nullptr);
}
extern "C" imt_status imt_sharded_occupied(imt_tree* tree, uint64_t* occupied_total) {
    Parts p;
    IMT_TRY(parts_of_tree(tree, &p));
    if (!occupied_total) return fail(tree->ctx, IMT_ERR_INVALID_ARG, "null buffer");
    return sharded_non_inclusion(p, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, occupied_total);
}
extern "C" imt_status imt_sharded_non_inclusion_paths(imt_tree* tree, const void* values, size_t q, uint64_t* low_idx, uint8_t* matched,
                                                      void* low_leaves, void* siblings, uint8_t* helpers, uint8_t* is_largest) {
    Parts p;
    IMT_TRY(parts_of_tree(tree, &p));
    return sharded_non_inclusion(p, values, q, low_idx, matched, low_leaves, siblings, helpers, is_largest, nullptr);
}
extern "C" imt_status imt_sharded_insert_batch(imt_tree* tree, const void* new_vals, size_t b, uint64_t first_idx, imt_insert_witness* w) {
    Parts p;
    IMT_TRY(parts_of_tree(tree, &p));
    return sharded_insert(p, new_vals, b, first_idx, w);
}
extern "C" imt_status imt_sharded_trace_proofs(imt_tree* tree, const uint64_t* indices, size_t q, uint64_t* positions, size_t* n_mine, void* states) {
    if (!tree) return IMT_ERR_INVALID_ARG;
    imt_ctx* ctx = tree->ctx;
    if (!n_mine || (q && (!indices || !positions))) return fail(ctx, IMT_ERR_INVALID_ARG, "null buffer");
    *n_mine = 0;
    if (!tree->cap_valid && tree->world > 1) return fail(ctx, IMT_ERR_INVALID_ARG, "no cap attached: exchange the subtree roots first");
    bool oob = false;
    const std::vector<uint64_t> pos = owned_positions(tree, indices, q, (uint64_t)tree->n * tree->world, &oob);
    if (oob) return fail(ctx, IMT_ERR_INDEX_OOB, "index out of bounds");
    std::vector<uint64_t> mine(pos.size());
    for (size_t j = 0; j < pos.size(); ++j) mine[j] = indices[pos[j]], positions[j] = pos[j];
    *n_mine = pos.size();
    if (pos.empty()) return IMT_OK;
    if (!states) return fail(ctx, IMT_ERR_INVALID_ARG, "null buffer");
    return imt_tree_trace_proofs(tree, mine.data(), mine.size(), states);
}

// ---- one process, N GPUs
extern "C" imt_status imt_mtree_get_proofs(imt_mtree* mt, const uint64_t* indices, size_t q, void* siblings, uint8_t* helpers) {
    Parts p;
    IMT_TRY(parts_of_mtree(mt, &p));
    if (q && !siblings) return fail(p.t[0]->ctx, IMT_ERR_INVALID_ARG, "null buffer");
    return sharded_gather_host(p, indices, q, siblings, helpers, nullptr, nullptr);
}
extern "C" imt_status imt_mtree_leaves(imt_mtree* mt, const uint64_t* indices, size_t q, void* leaves, uint8_t* is_largest) {
    Parts p;
    IMT_TRY(parts_of_mtree(mt, &p));
    return sharded_gather_host(p, indices, q, nullptr, nullptr, leaves, is_largest);
}
extern "C" imt_status imt_mtree_low_leaf_lookup(imt_mtree* mt, const void* values, size_t q, uint64_t* low_idx, uint8_t* matched) {
    Parts p;
    IMT_TRY(parts_of_mtree(mt, &p));
    if (q && !low_idx) return fail(p.t[0]->ctx, IMT_ERR_INVALID_ARG, "null buffer");
    return sharded_non_inclusion(p, values, q, low_idx, matched, nullptr, nullptr, nullptr, nullptr, nullptr);
}
extern "C" imt_status imt_mtree_occupied(imt_mtree* mt, uint64_t* occupied_total) {
    Parts p;
    IMT_TRY(parts_of_mtree(mt, &p));
    if (!occupied_total) return fail(p.t[0]->ctx, IMT_ERR_INVALID_ARG, "null buffer");
    return sharded_non_inclusion(p, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, occupied_total);
}
extern "C" imt_status imt_mtree_non_inclusion_paths(imt_mtree* mt, const void* values, size_t q, uint64_t* low_idx, uint8_t* matched,
                                                    void* low_leaves, void* siblings, uint8_t* helpers, uint8_t* is_largest) {
    Parts p;
    IMT_TRY(parts_of_mtree(mt, &p));
    return sharded_non_inclusion(p, values, q, low_idx, matched, low_leaves, siblings, helpers, is_largest, nullptr);
}
extern "C" imt_status imt_mtree_insert_batch(imt_mtree* mt, const void* new_vals, size_t b, uint64_t first_idx, imt_insert_witness* w) {
    Parts p;
    IMT_TRY(parts_of_mtree(mt, &p));
    return sharded_insert(p, new_vals, b, first_idx, w);
}
// Every device traces the queries it owns — each traced hash reads stored nodes of that device's subtree or of the replicated
// cap — and drains them over its own PCIe link straight to their place in the caller's array (one copy per query: a depth-24
// path is 304 KB of trace, large enough for the link).
extern "C" imt_status imt_mtree_trace_proofs(imt_mtree* mt, const uint64_t* indices, size_t q, void* states) {
    Parts p;
    IMT_TRY(parts_of_mtree(mt, &p));
    imt_ctx* c0 = p.t[0]->ctx;
    if (q && (!indices || !states)) return fail(c0, IMT_ERR_INVALID_ARG, "null buffer");
    if (q == 0) return IMT_OK;
    const uint64_t n_total = p.n_total();
    const unsigned depth = p.t[0]->depth + p.t[0]->cap_depth;
    const size_t per_query = (size_t)depth * trace_fe_per_hash(c0, 2) * sizeof(Fr);
    for (size_t i = 0; i < q; ++i)
        if (indices[i] >= n_total) return fail(c0, IMT_ERR_INDEX_OOB, "index out of bounds");
    return for_each_part(p, [&](size_t s) -> imt_status {
        imt_tree* t = p.t[s];
        imt_ctx* ctx = t->ctx;
        bool oob = false;
        const std::vector<uint64_t> pos = owned_positions(t, indices, q, n_total, &oob);
        if (pos.empty()) return IMT_OK;
        IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
        std::vector<uint64_t> mine(pos.size());
        for (size_t j = 0; j < pos.size(); ++j) mine[j] = indices[pos[j]];
        size_t chunk = 4096;
        while (chunk > 64 && chunk * per_query > ((size_t)2 << 30)) chunk >>= 1;
        chunk = std::min(chunk, pos.size());
        DevBuf di(ctx), buf0(ctx), buf1(ctx);
        IMT_TRY_CUDA(ctx, di.alloc(mine.size() * sizeof(uint64_t)));
        IMT_TRY_CUDA(ctx, buf0.alloc(chunk * per_query));
        if (pos.size() > chunk) IMT_TRY_CUDA(ctx, buf1.alloc(chunk * per_query));
        IMT_TRY_CUDA(ctx, cudaMemcpyAsync(di.p, mine.data(), mine.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        IMT_TRY(clear_err(ctx));
        void* bufs[2] = {buf0.p, buf1.p};
        Event produced[2], drained[2];
        for (int k = 0; k < 2; ++k) {
            IMT_TRY_CUDA(ctx, produced[k].create());
            IMT_TRY_CUDA(ctx, drained[k].create());
        }
        size_t c = 0;
        for (size_t off = 0; off < pos.size(); off += chunk, ++c) {
            const size_t cq = std::min(chunk, pos.size() - off);
            const int bsel = (int)(c & 1);
            if (c >= 2) IMT_TRY_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, drained[bsel], 0));
            IMT_TRY(launch_tree_trace(t, di.as<uint64_t>() + off, cq, bufs[bsel]));
            IMT_TRY_CUDA(ctx, cudaEventRecord(produced[bsel], ctx->stream));
            IMT_TRY_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, produced[bsel], 0));
            for (size_t j = 0; j < cq; ++j)
                IMT_TRY_CUDA(ctx, cudaMemcpyAsync((char*)states + pos[off + j] * per_query, (const char*)bufs[bsel] + j * per_query, per_query,
                                                  cudaMemcpyDeviceToHost, ctx->copy_stream));
            IMT_TRY_CUDA(ctx, cudaEventRecord(drained[bsel], ctx->copy_stream));
        }
        IMT_TRY_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
        return finish(ctx);
    });
}

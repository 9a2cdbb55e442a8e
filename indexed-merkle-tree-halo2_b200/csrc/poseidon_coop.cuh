// Latency-oriented Poseidon for the top of the tree: THREE lanes cooperate on one hash.
//
// A level with fewer nodes than the GPU has warp slots cannot be sped up by more threads: with one thread per hash
// (poseidon.cuh) such a level costs one full hash latency, ~0.57 ms of dependent IMADs, however few nodes it has, and
// a depth-24 build ends with ~15 of them. Here each of the three state lanes of a hash lives in its own SIMT lane
// (a quad per hash, the 4th lane idles), every lane runs the SAME instruction stream on its own element, and values
// cross lanes with warp shuffles:
//   full round     lane i: u_i = s_i^5 + c_i            | exchange u | lane i: s_i' = M[i] . u
//   partial round  lane 0: x^2, x^4                     (lanes 1, 2 idle through two multiplications)
//                  lane 0: u = x^4 x + c   lane 1: P1 = s1 row1   lane 2: P2 = s2 row2      (one multiply-reduce slot)
//                  exchange u, P1, P2
//                  lane 0: s0' = u row0 + (P1 + P2)   lane i: s_i' = u col_i + s_i          (one multiply-add slot)
// i.e. 4 dependent multiply-reduce slots per partial round instead of 8, 4 instead of 18 per full round: ~1.7x lower
// latency per hash. Results are the same field elements as poseidon.cuh (canonical on output), bit for bit.
// Round constants come from a copy of the parameters in global memory (the address depends on the lane).
#pragma once
#include "kernels_common.cuh"

namespace imt {

__device__ __forceinline__ void ld_fe(uint32_t* x, const Fr* p) { load_fe(x, reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void shfl_fe(uint32_t* d, const uint32_t* s, int src_lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = __shfl_sync(0xffffffffu, s[i], src_lane);
}

// (the extended S-box trace is produced by the one-thread-per-hash kernels only)
struct NoCoopTrace {
    __device__ __forceinline__ void emit(const uint32_t*, int) {}
};
// Witness-trace sink of the cooperative kernels: lane r writes element r of every traced state (3 x 32 contiguous bytes
// per state from the three lanes of a quad).
struct CoopTraceSink {
    uint4* dst;  // next state of this hash (3 FE each)
    int fmt;
    bool on;     // false for the padding quads of the last warp: they run the same code (shuffles are collective) but store nothing
    __device__ __forceinline__ void emit(const uint32_t* x, int r) {
        if (on && r < 3) {
            uint32_t t[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = x[i];
            if (fmt == kFmtCanonical) from_mont(t, t);
            else canonicalize(t);
            store_fe(dst + 2 * r, t);
        }
        dst += 6;
    }
};

// x: this lane's state element (role r = 0, 1, 2; r = 3 mirrors lane 2 and is ignored). base = first lane of the quad.
template <class Sink>
__device__ __forceinline__ void permute_coop(uint32_t* x, const PoseidonParams* __restrict__ G, int r, int base, Sink& sink) {
    const int rr = r < 3 ? r : 2;
    const bool lead = r == 0;
    {
        uint32_t c[8];
        ld_fe(c, &G->pre[rr]);
        add_semi(x, x, c);
    }
    sink.emit(x, r);
#pragma unroll 1
    for (int round = 0; round < kRF + kRP; ++round) {
        const bool full = round < kHalfF || round >= kHalfF + kRP;
        if (full) {
            const int fr = round < kHalfF ? round : round - kRP;
            uint32_t c[8], u0[8], u1[8], u2[8], m0[8], m1[8], m2[8];
            ld_fe(c, &G->full[fr][rr]);
            sbox_add(x, x, c);
            shfl_fe(u0, x, base);
            shfl_fe(u1, x, base + 1);
            shfl_fe(u2, x, base + 2);
            const Fr(*m)[3] = (round == kHalfF - 1) ? G->pre_sparse : G->mds;
            ld_fe(m0, &m[rr][0]);
            ld_fe(m1, &m[rr][1]);
            ld_fe(m2, &m[rr][2]);
            dot3(x, u0, u1, u2, m0, m1, m2);
            sink.emit(x, r);
        } else {
            const PartialRound* pr = &G->partial[round - kHalfF];
            uint32_t x2[8], x4[8], a[8], b[8], c[8], t[8];
            mont_sqr(x2, x);
            mont_sqr(x4, x2);
            ld_fe(b, lead ? &pr->c : &pr->row[rr]);  // lead: the round constant (added), others: their row entry (multiplied)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = lead ? x4[i] : x[i];
                c[i] = lead ? b[i] : 0u;
                b[i] = lead ? x[i] : b[i];
            }
            mul_add(t, a, b, c);  // lead: u = x^4 x + c ; lane i: P_i = s_i row_i
            uint32_t u[8], p1[8], p2[8], y[8], k[8];
            shfl_fe(u, t, base);
            shfl_fe(p1, t, base + 1);
            shfl_fe(p2, t, base + 2);
            add_semi(y, p1, p2);
            ld_fe(k, lead ? &pr->row[0] : &pr->col[rr - 1]);
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = lead ? y[i] : x[i];
            mul_add(x, u, k, y);  // lead: s0' = u row0 + P1 + P2 ; lane i: s_i' = u col_i + s_i
            sink.emit(x, r);
        }
    }
}

// out[h] = H(in[ARITY*h .. ARITY*h + ARITY)) for a SMALL batch (one tree level near the root, the leaves of an insert
// batch, a handful of hashes from the reference-style API): 4 threads per hash. Formats as in k_hash.
template <int ARITY>
__global__ void __launch_bounds__(128) k_hash_coop(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n, int in_fmt, int out_fmt,
                                                   const PoseidonParams* __restrict__ G, uint32_t* __restrict__ err) {
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t h = tid >> 2;
    const int r = (int)(tid & 3);
    const int base = (int)(threadIdx.x & 31) & ~3;
    const size_t hc = h < n ? h : n - 1;  // lanes past the end recompute the last hash: every lane must reach the shuffles
    uint32_t x[8], second[8];             // this lane's state element; what it absorbs before the second permutation
    bool ok = true;
    if (r == 0) {
        ld_fe(x, &G->cap);
    } else {
        load_fe(x, in + 2 * (ARITY * hc + (r == 1 ? 0 : 1)));
        ok &= ingest(x, in_fmt);
    }
    // second absorb: the remaining input (ARITY 3) followed by the padding 1
    if (ARITY == 3 && r == 1) {
        load_fe(second, in + 2 * (ARITY * hc + 2));
        ok &= ingest(second, in_fmt);
    } else {
        ld_fe(second, &G->one);
        const bool pad_here = ARITY == 3 ? r == 2 : r == 1;
#pragma unroll
        for (int i = 0; i < 8; ++i) second[i] = pad_here ? second[i] : 0u;
    }
    if (!ok) atomicOr(err, kErrNonCanonical);
    NoCoopTrace nt;
    permute_coop(x, G, r, base, nt);
    add_semi(x, x, second);
    permute_coop(x, G, r, base, nt);
    if (r == 1 && h < n) {
        canonicalize(x);
        egress(x, out_fmt);
        store_fe(out + 2 * h, x);
    }
}

// Batched verify_proof / compute_merkle_root for a SMALL batch of paths: a quad folds one path, level after level.
// Same contract as k_fold_paths (kernels.cuh), including the optional witness trace.
__global__ void __launch_bounds__(128) k_fold_paths_coop(const uint4* __restrict__ leaves, const uint64_t* __restrict__ idx,
                                                         const uint4* __restrict__ siblings, const uint4* __restrict__ roots, size_t q,
                                                         unsigned depth, int fmt, uint8_t* __restrict__ ok_out, uint4* __restrict__ roots_out,
                                                         uint4* __restrict__ states, const PoseidonParams* __restrict__ G,
                                                         uint32_t* __restrict__ err) {
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t p = tid >> 2;
    const int r = (int)(tid & 3);
    const int base = (int)(threadIdx.x & 31) & ~3;
    const size_t pc = p < q ? p : q - 1;
    const bool active = p < q;
    uint32_t h[8], one[8];
    load_fe(h, leaves + 2 * pc);  // every lane of the quad carries the running digest
    bool ok = ingest(h, fmt);
    ld_fe(one, &G->one);
#pragma unroll
    for (int i = 0; i < 8; ++i) one[i] = (r == 1) ? one[i] : 0u;
    uint64_t index = idx[pc];
#pragma unroll 1
    for (unsigned l = 0; l < depth; ++l) {
        uint32_t s[8], x[8];
        load_fe(s, siblings + 2 * (pc * depth + l));
        ok &= ingest(s, fmt);
        const bool left = (index & 1) == 0;
        const bool take_h = (r == 1) == left;  // lane 1 holds the left input, lanes 2 (and 3) the right one
        if (r == 0) ld_fe(x, &G->cap);
        else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = take_h ? h[i] : s[i];
        }
        if (states) {  // warp-uniform: `states` is a kernel argument
            CoopTraceSink sink{states + (pc * depth + l) * (size_t)(kStatesPerHash * 3 * 2), fmt, active};
            permute_coop(x, G, r, base, sink);
            add_semi(x, x, one);
            permute_coop(x, G, r, base, sink);
        } else {
            NoCoopTrace nt;
            permute_coop(x, G, r, base, nt);
            add_semi(x, x, one);
            permute_coop(x, G, r, base, nt);
        }
        canonicalize(x);
        shfl_fe(h, x, base + 1);  // the digest is lane 1's element
        index >>= 1;
    }
    if (!ok) atomicOr(err, kErrNonCanonical);
    canonicalize(h);
    if (r != 1 || !active) return;
    if (ok_out) {
        uint32_t rt[8];
        load_fe(rt, roots + 2 * p);
        if (!ingest(rt, fmt)) atomicOr(err, kErrNonCanonical);
        canonicalize(rt);
        bool same = true;
#pragma unroll
        for (int k = 0; k < 8; ++k) same &= rt[k] == h[k];
        ok_out[p] = (uint8_t)same;
    }
    if (roots_out) {
        egress(h, fmt);
        store_fe(roots_out + 2 * p, h);
    }
}

}  // namespace imt

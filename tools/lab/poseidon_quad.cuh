// LAB ONLY (tools/lab/latency_lab.cu) — measured and NOT shipped, see DESIGN.md "Cooperative kernel": 258 us per level against
// 283 us for the shipped 3-lane kernel under the serial carry discipline, but only 227 us against 233 us once both run with free
// carry chains — and under 7 of the 32 chain-head masks this kernel (never the shipped one) returned WRONG digests whose value
// changed with the ptxas optimisation level, i.e. a code-generation hazard we could not pin down. Kept as evidence.
//
// Latency-oriented Poseidon, second generation: FOUR lanes per hash and THREE dependent multiply slots per partial round.
//
// poseidon_coop.cuh spends 4 dependent multiply-reduce slots per partial round (x^2, x^4, x^4 x + c, row . s). The last two
// can be fused: with u = x^4 x + c,
//     s0' = row0 u + row1 s1 + row2 s2 = (row0 x) x^4 + (row0 c + row1 s1 + row2 s2) = y x^4 + K,
// and y = row0 x does not depend on the squarings, so a helper lane computes it while lane 0 squares. The multiplicative
// depth of a partial round drops to 3 — the minimum for x^5 (two squarings and one product). u itself is still needed
// (s_i' = col_i u + s_i) but no longer on the critical path: the helper lane computes it beside lane 0's last product.
//
// A quad per hash; every lane runs the SAME instruction stream, one fused multiply-add-reduce per slot:
//   slot 1   lane 0: x2 = x x          lanes 1,2: s_i = col_i' u' + s_i   (u', col' of the PREVIOUS round)   lane 3: y = row0 x
//   slot 2   lane 0: x4 = x2 x2        lane 1: P1 = row1 s1 + row0 c      lane 2: P2 = row2 s2               lane 3: (keeps y)
//   exchange lane 0 <- y, P1, P2       lane 3 <- x4
//   slot 3   lane 0: x' = y x4 + P1 + P2                                                                     lane 3: u = x4 x + c
//   exchange lanes 1,2 <- u            lane 3 <- x'
// 57 x 3 + 1 slots per permutation instead of 57 x 4; lane 0's squarings run as general products (SIMT: one stream).
// Full rounds are those of poseidon_coop.cuh (lane i: s_i^5 + c_i, exchange, row i of the matrix). Results are the same
// field elements as poseidon.cuh, bit for bit after canonicalisation (the representatives in [0, 2p) differ in between).
#pragma once
#include "poseidon_coop.cuh"

#ifdef IMT_QUAD_CLEAR_AT_LOOP_TOP
#define IMT_QUAD_LOOP_CLEAR cc::clear()
#else
#define IMT_QUAD_LOOP_CLEAR (void)0
#endif
#ifndef IMT_QUAD_DBG
#define IMT_QUAD_DBG 0
#endif
#define IMT_DBG_CLEAR(bit) do { if (IMT_QUAD_DBG & (bit)) cc::clear(); } while (0)

namespace imt {

// what the fused schedule needs beyond PoseidonParams: kc[k] = row0_k * c_k (the constant part of K), and a zero element
struct QuadAux {
    Fr kc[kRP];
    Fr zero;
};

// t = a b + (c1 + c2) R, reduced; inputs semi-reduced, a b / R + c1 + c2 + p must stay below 4p; output semi-reduced
__device__ __forceinline__ void fma2(uint32_t* t, const uint32_t* a, const uint32_t* b, const uint32_t* c1, const uint32_t* c2) {
    Wide w;
    wide_zero(w);
    mul_wide(w, a, b);
    add_hi(w, c1);
    add_hi(w, c2);
    redc(t, w);
    cond_sub_2p(t);
}
__device__ __forceinline__ void fma1(uint32_t* t, const uint32_t* a, const uint32_t* b, const uint32_t* c1) {
    Wide w;
    wide_zero(w);
    mul_wide(w, a, b);
    add_hi(w, c1);
    redc(t, w);
    cond_sub_2p(t);
}

struct NoQuadTrace {
    __device__ __forceinline__ void emit(const uint32_t*, int, int) {}
};
// Witness-trace sink: lane r (< 3) writes element r of state `idx` of the current permutation (3 x 32 contiguous bytes per
// state); the lanes of a quad reach a given state at different slots, hence the explicit index.
struct QuadTraceSink {
    uint4* base;  // state 0 of the current permutation of this hash
    int fmt;
    bool on;
    __device__ __forceinline__ void emit(const uint32_t* x, int r, int idx) {
        if (on && r < 3) {
            uint32_t t[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = x[i];
            if (fmt == kFmtCanonical) from_mont(t, t);
            else canonicalize(t);
            store_fe(base + 6 * (size_t)idx + 2 * r, t);
        }
    }
    __device__ __forceinline__ void next_perm() { base += 6 * (size_t)kStatesPerPerm; }
};

// x: this lane's state element (role r = 0, 1, 2; lane 3 mirrors lane 2 outside the partial rounds). base = first lane of the quad.
template <class Sink>
__device__ __forceinline__ void permute_quad(uint32_t* x, const PoseidonParams* __restrict__ G, const QuadAux* __restrict__ A, int r, int base,
                                             Sink& sink) {
    const int rr = r < 3 ? r : 2;
    {
        uint32_t c[8];
        ld_fe(c, &G->pre[rr]);
        add_semi(x, x, c);
    }
    sink.emit(x, r, 0);
    // ---- first half of the full rounds
#pragma unroll 1
    for (int round = 0; round < kHalfF; ++round) {
        IMT_QUAD_LOOP_CLEAR;
        uint32_t c[8], u0[8], u1[8], u2[8], m0[8], m1[8], m2[8];
        ld_fe(c, &G->full[round][rr]);
        IMT_DBG_CLEAR(1);
        sbox_add(x, x, c);
        shfl_fe(u0, x, base);
        shfl_fe(u1, x, base + 1);
        shfl_fe(u2, x, base + 2);
        const Fr(*m)[3] = (round == kHalfF - 1) ? G->pre_sparse : G->mds;
        ld_fe(m0, &m[rr][0]);
        ld_fe(m1, &m[rr][1]);
        ld_fe(m2, &m[rr][2]);
        IMT_DBG_CLEAR(2);
        dot3(x, u0, u1, u2, m0, m1, m2);
        sink.emit(x, r, 1 + round);
    }
    // ---- partial rounds, fused schedule. Per lane:  X = lane 0, 3: s0   lanes 1, 2: u of the previous round (0 at first)
    //                                                 S = lanes 1, 2: s_i   lanes 0, 3: 0
    const bool lead = r == 0, helper = r == 3, side = r == 1 || r == 2;
    uint32_t X[8], S[8];
    {
        uint32_t s0[8];
        shfl_fe(s0, x, base);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            X[i] = side ? 0u : s0[i];
            S[i] = side ? x[i] : 0u;
        }
    }
#pragma unroll 1
    for (int k = 0; k < kRP; ++k) {
        IMT_QUAD_LOOP_CLEAR;
        const PartialRound* pr = &G->partial[k];
        const PartialRound* pp = &G->partial[k ? k - 1 : 0];
        uint32_t T1[8], T2[8], B[8], C[8];
        // slot 1: lane 0  x x        lanes 1, 2  u' col' + s_i       lane 3  x row0
        ld_fe(B, side ? &pp->col[r - 1] : &pr->row[0]);
#pragma unroll
        for (int i = 0; i < 8; ++i) B[i] = lead ? X[i] : B[i];
        IMT_DBG_CLEAR(4);
        fma1(T1, X, B, S);
        if (k) sink.emit(T1, side ? r : 3, kHalfF + k);  // s1, s2 of the state after partial round k - 1
        uint32_t Y[8];
        shfl_fe(Y, T1, base + 3);  // y for lane 0 (consumed in slot 3)
#pragma unroll
        for (int i = 0; i < 8; ++i) S[i] = side ? T1[i] : 0u;
        // slot 2: lane 0  x2 x2      lane 1  s1 row1 + row0 c        lane 2  s2 row2      lane 3  idle (the product is discarded)
        ld_fe(B, side ? &pr->row[r] : &G->one);
        ld_fe(C, r == 1 ? &A->kc[k] : &A->zero);
#pragma unroll
        for (int i = 0; i < 8; ++i) B[i] = lead ? T1[i] : B[i];
        IMT_DBG_CLEAR(8);
        fma1(T2, T1, B, C);
        IMT_DBG_CLEAR(32);
        cond_sub_p(T2);  // canonical: lane 0 adds two of these on top of a product
        // exchange: lane 0 <- P1, P2; lane 3 <- x4
        uint32_t V1[8], V2[8];
        shfl_fe(V1, T2, lead ? base + 1 : base);
        shfl_fe(V2, T2, base + 2);
        // slot 3: lane 0  y x4 + P1 + P2      lane 3  x4 x + c
        uint32_t T3[8], a3[8], b3[8], c3[8];
        ld_fe(c3, &pr->c);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a3[i] = lead ? Y[i] : V1[i];
            b3[i] = lead ? T2[i] : X[i];
            c3[i] = lead ? V1[i] : c3[i];
            V2[i] = lead ? V2[i] : 0u;
        }
        IMT_DBG_CLEAR(16);
        fma2(T3, a3, b3, c3, V2);
        sink.emit(T3, lead ? 0 : 3, kHalfF + 1 + k);  // s0 of the state after partial round k
        // exchange: lanes 1, 2 <- u (lane 3); lane 3 <- x' (lane 0)
        shfl_fe(X, T3, side ? base + 3 : base);
    }
    // ---- flush: the last round's u into s1, s2; then back to one element per lane
    {
        uint32_t B[8], T1[8];
        ld_fe(B, side ? &G->partial[kRP - 1].col[r - 1] : &G->one);
        fma1(T1, X, B, S);
        sink.emit(T1, side ? r : 3, kHalfF + kRP);
        uint32_t s2[8];
        shfl_fe(s2, T1, base + 2);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = lead ? X[i] : (helper ? s2[i] : T1[i]);
    }
    // ---- second half of the full rounds
#pragma unroll 1
    for (int round = kHalfF; round < kRF; ++round) {
        IMT_QUAD_LOOP_CLEAR;
        uint32_t c[8], u0[8], u1[8], u2[8], m0[8], m1[8], m2[8];
        ld_fe(c, &G->full[round][rr]);
        IMT_DBG_CLEAR(1);
        sbox_add(x, x, c);
        shfl_fe(u0, x, base);
        shfl_fe(u1, x, base + 1);
        shfl_fe(u2, x, base + 2);
        ld_fe(m0, &G->mds[rr][0]);
        ld_fe(m1, &G->mds[rr][1]);
        ld_fe(m2, &G->mds[rr][2]);
        IMT_DBG_CLEAR(2);
        dot3(x, u0, u1, u2, m0, m1, m2);
        sink.emit(x, r, 1 + kRP + round);
    }
}

// kc[k] = row0_k * c_k, once per context
__global__ void k_quad_aux(const PoseidonParams* __restrict__ G, QuadAux* __restrict__ A) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > kRP) return;
    uint32_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (k < kRP) {
        uint32_t a[8], b[8];
        ld_fe(a, &G->partial[k].row[0]);
        ld_fe(b, &G->partial[k].c);
        mont_mul(t, a, b);
        canonicalize(t);
        store_fe(reinterpret_cast<uint4*>(&A->kc[k]), t);
    } else {
        store_fe(reinterpret_cast<uint4*>(&A->zero), t);
    }
}

// out[h] = H(in[ARITY*h .. ARITY*h + ARITY)): same contract as k_hash_coop
template <int ARITY>
__global__ void __launch_bounds__(128) k_hash_quad(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n, int in_fmt, int out_fmt,
                                                   const PoseidonParams* __restrict__ G, const QuadAux* __restrict__ A, uint32_t* __restrict__ err) {
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t h = tid >> 2;
    const int r = (int)(tid & 3);
    const int base = (int)(threadIdx.x & 31) & ~3;
    const size_t hc = h < n ? h : n - 1;
    uint32_t x[8], second[8];
    bool ok = true;
    if (r == 0) {
        ld_fe(x, &G->cap);
    } else {
        load_fe(x, in + 2 * (ARITY * hc + (r == 1 ? 0 : 1)));
        ok &= ingest(x, in_fmt);
    }
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
    NoQuadTrace nt;
    permute_quad(x, G, A, r, base, nt);
    add_semi(x, x, second);
    permute_quad(x, G, A, r, base, nt);
    if (r == 1 && h < n) {
        canonicalize(x);
        egress(x, out_fmt);
        store_fe(out + 2 * h, x);
    }
}

}  // namespace imt

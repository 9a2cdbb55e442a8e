// Poseidon permutation + fixed-length sponge over BN254 Fr, T = 3, RATE = 2, R_F = 8, R_P = 57.
//
// Replaces what the reference gets from pse-poseidon (`Poseidon::<Fr,3,2>::new(8,57)`, `update`,
// `squeeze_and_reset`: call sites /root/reference/src/utils.rs:46-47, 96-100 and
// src/indexed_merkle_tree.rs:370-376, 407-415, 510-518, 663-669, 807-809). Round schedule = the "optimized"
// one pse-poseidon and halo2-base's in-circuit hasher (indexed_merkle_tree.rs:92, 194, 271, 299) share:
//   s += start[0]; 3 x { s = s^5 + start[i]; s = MDS s }; s = s^5 + start[4]; s = PRE_SPARSE s;
//   57 x { s0 = s0^5 + partial[k]; s = SPARSE_k s }; 3 x { s = s^5 + end[i]; s = MDS s }; s = s^5; s = MDS s
// All state values are Montgomery-form and semi-reduced ([0, 2p)); see fr.cuh.
#pragma once
#include "fr.cuh"

namespace imt {

constexpr int kT = 3;
constexpr int kRF = 8;
constexpr int kRP = 57;
constexpr int kHalfF = kRF / 2;
constexpr int kStatesPerPerm = 1 + kRF + kRP;     // 66: after the pre-add, then after every round's linear layer
constexpr int kStatesPerHash = 2 * kStatesPerPerm;  // 132: every fixed-length hash of <= 3 inputs is 2 permutations
constexpr int kSboxPerPerm = kRF * kT + kRP;        // 81 S-boxes per permutation: 3 per full round, 1 per partial round
constexpr int kSboxPerHash = 2 * kSboxPerPerm;      // 162, each traced as (x^2, x^4, x^5 + c)

struct PartialRound {
    Fr c;        // optimized partial-round constant (added to s0 after the S-box)
    Fr row[3];   // sparse matrix: s0' = row . s
    Fr col[2];   //                s_i' = col[i-1] * s0 + s_i
};

// Everything Poseidon::new(8,57) derives at construction, in Montgomery form, canonical (< p).
struct PoseidonParams {
    Fr full[kRF][3];      // full[r]: constants added after the S-box of full round r (r = 0..7); full[7] = 0.
    Fr pre[3];            // start[0]: added before the first S-box
    Fr mds[3][3];
    Fr pre_sparse[3][3];  // linear layer of full round 3 (the last of the first half)
    PartialRound partial[kRP];
    Fr cap;               // 2^64: initial state[0] of the sponge
    Fr one;               // 1: the padding element
};

// A trace sink sees the state after the pre-add and after every round's linear layer (emit) and, for the optional extended
// trace of SURVEY 8a row 9, the three product cells of every S-box the chip materialises (emit_sbox: x^2, x^4, x^5 + c).
struct NoTrace {
    IMT_HD void emit(const uint32_t (*)[8]) {}
    IMT_HD void emit_sbox(const uint32_t*, const uint32_t*, const uint32_t*) {}
};

// u = x^5 + c   (x semi-reduced, c canonical constant; u semi-reduced)
template <class Sink>
IMT_HD void sbox_add(uint32_t* u, const uint32_t* x, const uint32_t* c, Sink& sink) {
    uint32_t x2[8], x4[8];
    mont_sqr(x2, x);
    mont_sqr(x4, x2);
    Wide w;
    wide_zero(w);
    mul_wide(w, x4, x);
    add_hi(w, c);
    redc(u, w);       // < 4p^2/2^256 + p + p < 2.76 p
    cond_sub_2p(u);
    sink.emit_sbox(x2, x4, u);
}
IMT_HD void sbox_add(uint32_t* u, const uint32_t* x, const uint32_t* c) {
    NoTrace nt;
    sbox_add(u, x, c, nt);
}
IMT_HD void sbox(uint32_t* u, const uint32_t* x) {
    uint32_t x2[8], x4[8];
    mont_sqr(x2, x);
    mont_sqr(x4, x2);
    mont_mul(u, x4, x);  // semi-reduced
}

// r = m0*u0 + m1*u1 + m2*u2 with ONE reduction (m canonical constants, u semi-reduced)
IMT_HD void dot3(uint32_t* r, const uint32_t* u0, const uint32_t* u1, const uint32_t* u2, const uint32_t* m0,
                 const uint32_t* m1, const uint32_t* m2) {
    Wide w;
    wide_zero(w);
    mul_wide(w, u0, m0);
    mac_wide(w, u1, m1);
    mac_wide(w, u2, m2);
    redc(r, w);       // < 6p^2/2^256 + p < 2.14 p
    cond_sub_2p(r);
}
// r = m*u + s with one reduction
IMT_HD void mul_add(uint32_t* r, const uint32_t* u, const uint32_t* m, const uint32_t* s) {
    Wide w;
    wide_zero(w);
    mul_wide(w, u, m);
    add_hi(w, s);
    redc(r, w);       // < 2p^2/2^256 + 2p + p < 3.38 p
    cond_sub_2p(r);
}

template <class Sink>
IMT_HD void full_round(uint32_t (*s)[8], const Fr* c, const Fr (*m)[3], Sink& sink) {
    uint32_t u0[8], u1[8], u2[8];
    sbox_add(u0, s[0], c[0].l, sink);
    sbox_add(u1, s[1], c[1].l, sink);
    sbox_add(u2, s[2], c[2].l, sink);
    dot3(s[0], u0, u1, u2, m[0][0].l, m[0][1].l, m[0][2].l);
    dot3(s[1], u0, u1, u2, m[1][0].l, m[1][1].l, m[1][2].l);
    dot3(s[2], u0, u1, u2, m[2][0].l, m[2][1].l, m[2][2].l);
    sink.emit(s);
}

template <class Sink>
IMT_HD void partial_round(uint32_t (*s)[8], const PartialRound& pr, Sink& sink) {
    uint32_t u[8], n0[8];
    sbox_add(u, s[0], pr.c.l, sink);
    dot3(n0, u, s[1], s[2], pr.row[0].l, pr.row[1].l, pr.row[2].l);
    mul_add(s[1], u, pr.col[0].l, s[1]);
    mul_add(s[2], u, pr.col[1].l, s[2]);
#pragma unroll
    for (int i = 0; i < 8; ++i) s[0][i] = n0[i];
    sink.emit(s);
}

// (s0, s1, s2) <- (s1, s2, s0)
IMT_HD void rotate3(uint32_t (*s)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t t = s[0][i];
        s[0][i] = s[1][i];
        s[1][i] = s[2][i];
        s[2][i] = t;
    }
}

#ifdef IMT_PERMUTE_UNROLLED
// One permutation, in place: every S-box / dot product of a round is its own inlined copy (~150 KB of SASS).
template <class Sink>
IMT_HD void permute(uint32_t (*s)[8], const PoseidonParams& P, Sink& sink) {
    add_semi(s[0], s[0], P.pre[0].l);
    add_semi(s[1], s[1], P.pre[1].l);
    add_semi(s[2], s[2], P.pre[2].l);
    sink.emit(s);
#pragma unroll 1
    for (int r = 0; r < kHalfF; ++r) full_round(s, P.full[r], r == kHalfF - 1 ? P.pre_sparse : P.mds, sink);
#pragma unroll 1
    for (int k = 0; k < kRP; ++k) partial_round(s, P.partial[k], sink);
#pragma unroll 1
    for (int r = kHalfF; r < kRF; ++r) full_round(s, P.full[r], P.mds, sink);
}
#else
// One permutation, in place. `P` may live in __constant__, shared or host memory.
// Compact form: ONE copy each of the S-box, the 3-term dot product and the multiply-add serves all 65 rounds (the
// state is rotated through fixed registers instead of duplicating code per lane), so the whole permutation is
// ~25 KB of SASS and stays resident in the 32 KB L1.5 instruction cache; the fully inlined form above is ~150 KB and
// loses ~20 % of the multiply pipe to instruction-fetch stalls (profiles/). All branches are warp-uniform.
template <class Sink>
IMT_HD void permute(uint32_t (*s)[8], const PoseidonParams& P, Sink& sink) {
    add_semi(s[0], s[0], P.pre[0].l);
    add_semi(s[1], s[1], P.pre[1].l);
    add_semi(s[2], s[2], P.pre[2].l);
    sink.emit(s);
#pragma unroll 1
    for (int r = 0; r < kRF + kRP; ++r) {
        const bool full = r < kHalfF || r >= kHalfF + kRP;
        const int fr = r < kHalfF ? r : r - kRP;  // index into full[] (full rounds)
        const int pk = full ? 0 : r - kHalfF;     // index into partial[] (partial rounds)
        // ---- S-box layer (+ the constants folded behind it): all three lanes, or lane 0 only
        const int lanes = full ? 3 : 1;
#pragma unroll 1
        for (int j = 0; j < lanes; ++j) {
            const uint32_t* c = full ? P.full[fr][j].l : P.partial[pk].c.l;
            sbox_add(s[0], s[0], c, sink);
            if (full) rotate3(s);  // three rotations bring the lanes back in order
        }
        // ---- linear layer: dense 3x3 (full rounds) or sparse (row . s ; s_i + col_i * s0)
        const Fr(*m)[3] = (r == kHalfF - 1) ? P.pre_sparse : P.mds;
        uint32_t n[3][8] = {};
#pragma unroll 1
        for (int j = 0; j < 3; ++j) {
            uint32_t t[8];
            if (full || j == 0) {
                const Fr* row = full ? m[j] : P.partial[pk].row;
                dot3(t, s[0], s[1], s[2], row[0].l, row[1].l, row[2].l);
            } else {
                uint32_t sj[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) sj[i] = (j == 1) ? s[1][i] : s[2][i];
                mul_add(t, s[0], P.partial[pk].col[j - 1].l, sj);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {  // n <- (n1, n2, t): after three steps n = (t0, t1, t2)
                n[0][i] = n[1][i];
                n[1][i] = n[2][i];
                n[2][i] = t[i];
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s[0][i] = n[0][i], s[1][i] = n[1][i], s[2][i] = n[2][i];
        sink.emit(s);
    }
}
#endif

// H(in[0..ARITY)) = update(in) + squeeze_and_reset(), ARITY in {2, 3}: two permutations.
// Inputs: Montgomery form, semi-reduced. Output: Montgomery form, canonical.
// The two permutations share one copy of the round code (the loop is deliberately not unrolled).
template <int ARITY, class Sink>
IMT_HD void hash_fixed(uint32_t* out, const uint32_t (*in)[8], const PoseidonParams& P, Sink& sink) {
    uint32_t s[3][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        s[0][i] = P.cap.l[i];
        s[1][i] = in[0][i];
        s[2][i] = in[1][i];
    }
#pragma unroll 1
    for (int perm = 0; perm < 2; ++perm) {
        if (perm == 1) {  // second absorb: the remaining input (if any) followed by the padding 1
            if constexpr (ARITY == 3) {
                add_semi(s[1], s[1], in[2]);
                add_semi(s[2], s[2], P.one.l);
            } else {
                add_semi(s[1], s[1], P.one.l);
            }
        }
        permute(s, P, sink);
    }
    canonicalize(s[1]);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = s[1][i];
}

}  // namespace imt

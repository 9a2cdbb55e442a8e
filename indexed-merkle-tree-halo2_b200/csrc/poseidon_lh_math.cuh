// The algebra of the lead / helper latency kernel (poseidon_lh.cuh), free of CUDA intrinsics so that the CPU tests compile it in
// host mode (tests/host_shim.cpp, emulated carry flag): the per-round tables and a sequential statement of the recurrence the
// kernel distributes over warps. With u = x^5 + c every output of a partial round (poseidon.cuh partial_round) is affine in u:
//     x'   = row_0 u + row_1 s_1 + row_2 s_2 = x^4 (row_0 x) + K          K   = row_0 c + row_1 s_1 + row_2 s_2
//     s_i' = col_i u + s_i                   = x^4 (col_i x) + (col_i c + s_i)
//     K'   = row_0' c' + row_1' s_1' + row_2' s_2' = x^4 (rho x) + (kappa + row_1' s_1 + row_2' s_2)
//            rho = row_1' col_1 + row_2' col_2,  kappa = rho c + row_0' c'                        (' = the next round)
#pragma once
#include "poseidon.cuh"

namespace imt {

// per partial round k, by helper role: the multiplier of slot A and the constant part of the addend of slot C (Montgomery, canonical)
struct LhRound {
    Fr mul_a[8];  // 0: row_0   1: col_1   2: col_2   3: rho   4: row_1 of round k + 1   5: row_2 of round k + 1   6, 7: 0
    Fr add_c[4];  // 0: 0       1: col_1 c   2: col_2 c   3: kappa
};
struct LhAux {
    LhRound round[kRP];
    Fr kc0;  // row_0 c of the first partial round: the constant part of the first K
};

// tables of round `pr`; nx = the next partial round, null for the last one (its K' is never used: rho = kappa = 0)
IMT_HD void lh_make_round(LhRound* o, const PartialRound& pr, const PartialRound* nx) {
    uint32_t r0n[8], r1n[8], r2n[8], cn[8], a[8], b[8], t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r0n[i] = nx ? nx->row[0].l[i] : 0u;
        r1n[i] = nx ? nx->row[1].l[i] : 0u;
        r2n[i] = nx ? nx->row[2].l[i] : 0u;
        cn[i] = nx ? nx->c.l[i] : 0u;
        o->mul_a[0].l[i] = pr.row[0].l[i];
        o->mul_a[1].l[i] = pr.col[0].l[i];
        o->mul_a[2].l[i] = pr.col[1].l[i];
        o->mul_a[4].l[i] = r1n[i];
        o->mul_a[5].l[i] = r2n[i];
        o->mul_a[6].l[i] = o->mul_a[7].l[i] = o->add_c[0].l[i] = 0u;
    }
    mont_mul(a, r1n, pr.col[0].l);
    mont_mul(b, r2n, pr.col[1].l);
    add_semi(t, a, b);
    canonicalize(t);  // rho
#pragma unroll
    for (int i = 0; i < 8; ++i) o->mul_a[3].l[i] = t[i];
    mont_mul(a, t, pr.c.l);
    mont_mul(b, r0n, cn);
    add_semi(t, a, b);
    canonicalize(t);  // kappa
#pragma unroll
    for (int i = 0; i < 8; ++i) o->add_c[3].l[i] = t[i];
    mont_mul(a, pr.col[0].l, pr.c.l);
    canonicalize(a);
    mont_mul(b, pr.col[1].l, pr.c.l);
    canonicalize(b);
#pragma unroll
    for (int i = 0; i < 8; ++i) o->add_c[1].l[i] = a[i], o->add_c[2].l[i] = b[i];
}
IMT_HD void lh_make_kc0(Fr* o, const PartialRound& pr0) {
    uint32_t a[8];
    mont_mul(a, pr0.row[0].l, pr0.c.l);
    canonicalize(a);
#pragma unroll
    for (int i = 0; i < 8; ++i) o->l[i] = a[i];
}

// The 57 partial rounds of one permutation in the lead / helper formulation, one role after the other (what the kernel runs in
// parallel lanes and warps): s = (x, s_1, s_2) semi-reduced in, semi-reduced out; the same field elements as 57 x partial_round.
IMT_HD void lh_partial_rounds(uint32_t (*s)[8], const PoseidonParams& P, const LhAux& A) {
    uint32_t x[8], s1[8], s2[8], K[8], d1[8], d2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = s[0][i], s1[i] = s[1][i], s2[i] = s[2][i];
    mont_mul(d1, s1, P.partial[0].row[1].l);  // prologue: the first K
    mont_mul(d2, s2, P.partial[0].row[2].l);
    add_semi(K, A.kc0.l, d1);
    add_semi(K, K, d2);
    for (int k = 0; k < kRP; ++k) {
        const LhRound& r = A.round[k];
        uint32_t x2[8], x4[8], y0[8], w1[8], w2[8], z[8], add[8], n1[8], n2[8], nk[8];
        mont_sqr(x2, x);  // lead
        mont_sqr(x4, x2);
        mont_mul(y0, x, r.mul_a[0].l);  // helpers, slot A
        mont_mul(w1, x, r.mul_a[1].l);
        mont_mul(w2, x, r.mul_a[2].l);
        mont_mul(z, x, r.mul_a[3].l);
        mont_mul(d1, s1, r.mul_a[4].l);
        mont_mul(d2, s2, r.mul_a[5].l);
        add_semi(add, r.add_c[1].l, s1);  // helpers, slot C
        mul_add(n1, x4, w1, add);
        add_semi(add, r.add_c[2].l, s2);
        mul_add(n2, x4, w2, add);
        add_semi(add, r.add_c[3].l, d1);
        add_semi(add, add, d2);
        mul_add(nk, x4, z, add);
        mul_add(x, x4, y0, K);  // lead: x' = x^4 y_0 + K
#pragma unroll
        for (int i = 0; i < 8; ++i) s1[i] = n1[i], s2[i] = n2[i], K[i] = nk[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s[0][i] = x[i], s[1][i] = s1[i], s[2][i] = s2[i];
}

}  // namespace imt

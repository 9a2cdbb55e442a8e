// Poseidon over BN254 Fr for ANY width: `Poseidon::<Fr, T, RATE>::new(r_f, r_p)` with RATE = T - 1.
//
// The reference's native tree and chip entry points are generic over T and RATE (/root/reference/src/utils.rs:6, 19;
// src/indexed_merkle_tree.rs:65, 127, 231); only <3, 2>(8, 57) is instantiated (indexed_merkle_tree.rs:362-365) and that
// instance has its own tuned kernels (poseidon.cuh, poseidon_coop.cuh). This header is the same optimized round schedule
// with T a template parameter and r_f / r_p / the input length run-time values, reading the parameters from a dense Fr
// array in global memory (every lane reads the same address: one broadcast transaction, L1-resident).
// Same compact form as poseidon.cuh: ONE S-box, ONE T-term dot product and ONE multiply-add serve all rounds; the state
// is rotated through fixed registers, every branch is warp-uniform.
#pragma once
#include <cstddef>

#include "fr.cuh"

namespace imt {

constexpr unsigned kSpecMinT = 2;
constexpr unsigned kSpecMaxT = 5;   // T-term dot products are reduced once: needs 2 T p^2 / 2^256 + p < 4p, i.e. T <= 7
constexpr unsigned kSpecMaxRounds = 256;

// Position of every parameter inside the dense array (units: Fr). Shared by the host generator and the kernels.
using std::size_t;

struct SpecLayout {
    unsigned t, r_f, r_p;
    IMT_HD size_t cap() const { return 0; }   // 2^64: initial state[0] of the sponge
    IMT_HD size_t one() const { return 1; }   // the padding element
    IMT_HD size_t pre(unsigned i) const { return 2 + i; }                                  // added before the first S-box
    IMT_HD size_t full(unsigned r, unsigned i) const { return 2 + t + (size_t)r * t + i; }  // after the S-box of full round r
    IMT_HD size_t mds(unsigned i, unsigned j) const { return 2 + t + (size_t)r_f * t + (size_t)i * t + j; }
    IMT_HD size_t pre_sparse(unsigned i, unsigned j) const { return mds(0, 0) + (size_t)t * t + (size_t)i * t + j; }
    IMT_HD size_t partial(unsigned k) const { return pre_sparse(0, 0) + (size_t)t * t + (size_t)k * 2 * t; }
    IMT_HD size_t partial_c(unsigned k) const { return partial(k); }                        // constant of partial round k
    IMT_HD size_t partial_row(unsigned k, unsigned i) const { return partial(k) + 1 + i; }  // s0' = row . s
    IMT_HD size_t partial_col(unsigned k, unsigned i) const { return partial(k) + 1 + t + i; }  // s_{i+1}' = col[i] s0 + s_{i+1}
    IMT_HD size_t total() const { return partial(r_p); }
    IMT_HD unsigned states_per_perm() const { return 1 + r_f + r_p; }
};

// Everything below also compiles for the HOST (fr.cuh's emulated carry flag): tests/host_shim.cpp runs exactly this
// source on the CPU against the oracle (tests/test_host_field.py). Test-only; the library has no CPU compute path.
IMT_HD void spec_load_const(uint32_t* x, const Fr* p) {
#ifdef __CUDA_ARCH__
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    x[0] = a.x, x[1] = a.y, x[2] = a.z, x[3] = a.w;
    x[4] = b.x, x[5] = b.y, x[6] = b.z, x[7] = b.w;
#else
    for (int i = 0; i < 8; ++i) x[i] = p->l[i];
#endif
}

// u = x^5 + c
template <class Sink>
IMT_HD void spec_sbox_add(uint32_t* u, const uint32_t* x, const Fr* c, Sink& sink) {
    uint32_t x2[8], x4[8], cc_[8];
    spec_load_const(cc_, c);
    mont_sqr(x2, x);
    mont_sqr(x4, x2);
    Wide w;
    wide_zero(w);
    mul_wide(w, x4, x);
    add_hi(w, cc_);
    redc(u, w);
    cond_sub_2p(u);
    sink.emit_sbox(x2, x4, u);
}

// r = sum_k row[k] * s[k], one reduction (row canonical constants, s semi-reduced): < 2 T p^2 / 2^256 + p < 4p for T <= 7
template <int T>
IMT_HD void spec_dot(uint32_t* r, const uint32_t (*s)[8], const Fr* row) {
    Wide w;
    wide_zero(w);
    uint32_t m[8];
    spec_load_const(m, row);
    mul_wide(w, s[0], m);
#pragma unroll
    for (int k = 1; k < T; ++k) {
        spec_load_const(m, row + k);
        mac_wide(w, s[k], m);
    }
    redc(r, w);
    cond_sub_2p(r);
}

// r = m * u + a
IMT_HD void spec_mul_add(uint32_t* r, const uint32_t* u, const Fr* m, const uint32_t* a) {
    uint32_t mm[8];
    spec_load_const(mm, m);
    Wide w;
    wide_zero(w);
    mul_wide(w, u, mm);
    add_hi(w, a);
    redc(r, w);
    cond_sub_2p(r);
}

// (s0, ..., sT-1) <- (s1, ..., sT-1, s0)
template <int T>
IMT_HD void spec_rotate(uint32_t (*s)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t t0 = s[0][i];
#pragma unroll
        for (int k = 0; k + 1 < T; ++k) s[k][i] = s[k + 1][i];
        s[T - 1][i] = t0;
    }
}

// One permutation, in place. Sink::emit(s) sees the state after the pre-add and after every round's linear layer.
template <int T, class Sink>
IMT_HD void spec_permute(uint32_t (*s)[8], const Fr* __restrict__ P, const SpecLayout L, Sink& sink) {
    {
        uint32_t c[8];
#pragma unroll
        for (int j = 0; j < T; ++j) {
            spec_load_const(c, P + L.pre(j));
            add_semi(s[j], s[j], c);
        }
    }
    sink.emit(s);
    const unsigned half = L.r_f / 2;
#pragma unroll 1
    for (unsigned r = 0; r < L.r_f + L.r_p; ++r) {
        const bool full = r < half || r >= half + L.r_p;
        const unsigned fr = r < half ? r : r - L.r_p;  // index among the full rounds
        const unsigned pk = full ? 0 : r - half;       // index among the partial rounds
        // ---- S-box layer (+ the constants folded behind it): all lanes, or lane 0 only
        const unsigned lanes = full ? T : 1;
#pragma unroll 1
        for (unsigned j = 0; j < lanes; ++j) {
            spec_sbox_add(s[0], s[0], full ? P + L.full(fr, j) : P + L.partial_c(pk), sink);
            if (full) spec_rotate<T>(s);  // T rotations bring the lanes back in order
        }
        // ---- linear layer: dense T x T (full rounds) or sparse (row . s ; s_j + col_j * s0)
        const Fr* dense = P + ((r + 1 == half) ? L.pre_sparse(0, 0) : L.mds(0, 0));
        uint32_t n[T][8] = {};
#pragma unroll 1
        for (unsigned j = 0; j < (unsigned)T; ++j) {
            uint32_t v[8];
            if (full || j == 0) {
                spec_dot<T>(v, s, full ? dense + (size_t)j * T : P + L.partial_row(pk, 0));
            } else {
                uint32_t sj[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    sj[i] = s[1][i];
#pragma unroll
                    for (int k = 2; k < T; ++k) sj[i] = (j == (unsigned)k) ? s[k][i] : sj[i];
                }
                spec_mul_add(v, s[0], P + L.partial_col(pk, j - 1), sj);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {  // n <- (n1, ..., nT-1, v): after T steps n = (v0, ..., vT-1)
#pragma unroll
                for (int k = 0; k + 1 < T; ++k) n[k][i] = n[k + 1][i];
                n[T - 1][i] = v[i];
            }
        }
#pragma unroll
        for (int k = 0; k < T; ++k)
#pragma unroll
            for (int i = 0; i < 8; ++i) s[k][i] = n[k][i];
        sink.emit(s);
    }
}

// The sponge: `update(inputs)` then `squeeze_and_reset()` (pse-poseidon; call sites utils.rs:46-47, indexed_merkle_tree.rs:374-375).
// state = [2^64, 0, ...]; every full RATE-chunk is added into state[1..] and permuted; the remainder (< RATE elements)
// followed by the padding element 1 is added and permuted once more; the digest is state[1].
// `load(j, x)` yields input j in Montgomery form (semi-reduced). arity / RATE + 1 permutations, ONE call site.
template <int T, class Load, class Sink>
IMT_HD void spec_sponge(uint32_t* digest, size_t arity, Load& load, const Fr* __restrict__ P, const SpecLayout L,
                                            Sink& sink) {
    constexpr int RATE = T - 1;
    uint32_t s[T][8];
    spec_load_const(s[0], P + L.cap());
#pragma unroll
    for (int k = 1; k < T; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) s[k][i] = 0;
    uint32_t one[8];
    spec_load_const(one, P + L.one());
#pragma unroll 1
    for (size_t base = 0;; base += RATE) {
        const size_t rem = arity - base;
        const bool last = rem < (size_t)RATE;
#pragma unroll
        for (int k = 0; k < RATE; ++k) {
            if ((size_t)k < rem) {
                uint32_t x[8];
                load(base + k, x);
                add_semi(s[1 + k], s[1 + k], x);
            } else if ((size_t)k == rem) {
                add_semi(s[1 + k], s[1 + k], one);
            }
        }
        spec_permute<T>(s, P, L, sink);
        if (last) break;
    }
    canonicalize(s[1]);
#pragma unroll
    for (int i = 0; i < 8; ++i) digest[i] = s[1][i];
}

}  // namespace imt

// BN254 scalar field (Fr) arithmetic for sm_100a: 8 x 32-bit Montgomery limbs, R = 2^256.
//
// Replaces the arithmetic the reference gets from halo2curves (BN254 Fr, named grumpkin::Fq at
// /root/reference/src/indexed_merkle_tree.rs:327; modulus literal at :382-385). In memory an element is
// bit-identical to halo2curves' [u64; 4] little-endian Montgomery form.
//
// Design (see DESIGN.md "Field multiplication"):
//   * A 512-bit product is held as TWO interleaved accumulators, `e` (64-bit slots at even limb positions) and
//     `o` (slots at odd limb positions). Every 32x32->64 partial product then lands on an aligned register
//     pair, so each is ONE IMAD.WIDE.U32(.X) with the carry riding the predicate chain — no lo/hi split, no
//     register moves. ptxas fuses each mad.lo.cc/madc.hi.cc pair below into that instruction.
//   * Products and the Montgomery reduction are SEPARATE steps, which buys (a) a dedicated squaring (36 wide
//     MACs instead of 64) for the x^2, x^4 of the S-box, and (b) lazy reduction: the three products of an
//     MDS / sparse-row dot product are summed in the wide accumulator and reduced once.
//   * Values are kept "semi-reduced" in [0, 2p) between operations (4p < 2^256, so a Montgomery product of
//     two semi-reduced values is again semi-reduced without any conditional subtraction); they are made
//     canonical only when stored.
//
// The carry primitives have a host emulation (a thread-local carry flag) so that exactly this source is unit
// tested on the CPU against big-integer arithmetic (tests/test_host_field.py). The emulation is test-only; the
// shipped library has no CPU compute path.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define IMT_HD __host__ __device__ __forceinline__
#else
#define IMT_HD inline
#endif

namespace imt {

struct alignas(16) Fr {
    uint32_t l[8];
};

// p = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
#define IMT_P0 0xf0000001u
#define IMT_P1 0x43e1f593u
#define IMT_P2 0x79b97091u
#define IMT_P3 0x2833e848u
#define IMT_P4 0x8181585du
#define IMT_P5 0xb85045b6u
#define IMT_P6 0xe131a029u
#define IMT_P7 0x30644e72u
// 2p
#define IMT_2P0 0xe0000002u
#define IMT_2P1 0x87c3eb27u
#define IMT_2P2 0xf372e122u
#define IMT_2P3 0x5067d090u
#define IMT_2P4 0x0302b0bau
#define IMT_2P5 0x70a08b6du
#define IMT_2P6 0xc2634053u
#define IMT_2P7 0x60c89ce5u
#define IMT_INV32 0xefffffffu  // -p^-1 mod 2^32

// ------------------------------------------------------------------------------------------ carry primitives
namespace cc {
#ifndef __CUDA_ARCH__
inline uint32_t& flag() {
    static thread_local uint32_t f = 0;
    return f;
}
#endif

IMT_HD uint32_t add_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
#else
    uint64_t t = (uint64_t)a + b;
    flag() = (uint32_t)(t >> 32);
    return (uint32_t)t;
#endif
}
IMT_HD uint32_t addc_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
#else
    uint64_t t = (uint64_t)a + b + flag();
    flag() = (uint32_t)(t >> 32);
    return (uint32_t)t;
#endif
}
IMT_HD uint32_t addc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
#else
    return a + b + flag();
#endif
}
IMT_HD uint32_t sub_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
#else
    uint64_t t = (uint64_t)a - b;
    flag() = (uint32_t)(t >> 63);  // borrow
    return (uint32_t)t;
#endif
}
IMT_HD uint32_t subc_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
#else
    uint64_t t = (uint64_t)a - b - flag();
    flag() = (uint32_t)(t >> 63);
    return (uint32_t)t;
#endif
}
IMT_HD uint32_t subc(uint32_t a, uint32_t b) {  // a - b - borrow, borrow not updated
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
#else
    return a - b - flag();
#endif
}
// INVARIANT: the carry flag is 0 between field operations. Every MAC chain STARTS with madwc_cc (it consumes that
// zero) and every chain ends by writing a provably-zero carry back. PTX has one carry flag, so this strings all
// chains of a thread into program order; without it ptxas overlaps so many independent chains that their carry
// predicates (7 per thread) spill into a GPR bitmask — thousands of extra LOP3/P2R per hash.
// clear the flag; `x` must be < 2^31 (top limb of a value < 2p) — ties the clear to the data just produced
IMT_HD void clear_after(uint32_t x) {
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm volatile("add.cc.u32 %0, %1, %1;" : "=r"(d) : "r"(x));
#else
    flag() = (uint32_t)(((uint64_t)x + x) >> 32);
#endif
}
IMT_HD void clear() {
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm volatile("add.cc.u32 %0, 0, 0;" : "=r"(d));
#else
    flag() = 0;
#endif
}
// {hi,lo} += a*b, carry out (starts a chain without reading the flag)                      -> IMAD.WIDE.U32   Rd, Pout, a, b, Rd
IMT_HD void madw_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
#else
    uint64_t pr = (uint64_t)a * b;
    uint64_t t = (uint64_t)lo + (uint32_t)pr;
    lo = (uint32_t)t;
    t = (uint64_t)hi + (uint32_t)(pr >> 32) + (t >> 32);
    hi = (uint32_t)t;
    flag() = (uint32_t)(t >> 32);
#endif
}
// {hi,lo} += a*b + carry in, carry out (continues a chain)         -> IMAD.WIDE.U32.X Rd, Pout, a, b, Rd, Pin
IMT_HD void madwc_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
#else
    uint64_t pr = (uint64_t)a * b;
    uint64_t t = (uint64_t)lo + (uint32_t)pr + flag();
    lo = (uint32_t)t;
    t = (uint64_t)hi + (uint32_t)(pr >> 32) + (t >> 32);
    hi = (uint32_t)t;
    flag() = (uint32_t)(t >> 32);
#endif
}
// First MAC of a chain. Default: it consumes the (provably zero) carry the previous chain left behind, which strings every
// chain of a thread into program order (see INVARIANT above) — right for the throughput kernels. A chain whose head does NOT
// read the flag is independent of what precedes it, so ptxas may overlap it with the previous chain (one more carry predicate
// live): the latency kernels (a lone warp per scheduler, poseidon_coop.cuh) gain from that. IMT_FREE_MASK selects which heads
// are free — bit 0 / 1: the even / odd accumulator chain of a product row, bit 2 / 3: of a reduction row, bit 4: squaring rows;
// IMT_FREE_CHAINS = all of them. Measured per combination by tools/latency_lab.cu.
#if defined(IMT_FREE_CHAINS) && !defined(IMT_FREE_MASK)
#define IMT_FREE_MASK 31
#endif
#ifndef IMT_FREE_MASK
#define IMT_FREE_MASK 0
#endif
template <int KIND>
IMT_HD void madw_head(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    if constexpr (((IMT_FREE_MASK) >> KIND) & 1) madw_cc(lo, hi, a, b);
    else madwc_cc(lo, hi, a, b);
}
constexpr int kHeadProdE = 0, kHeadProdO = 1, kHeadRedE = 2, kHeadRedO = 3, kHeadSqr = 4;
}  // namespace cc

// ------------------------------------------------------------------------------------------ wide accumulator
// value = sum e[i] 2^(32 i)  +  sum o[i] 2^(32 (i+1))  +  sum k[j] 2^(32 (8+j))
// `k` collects the carries that leave the top of a MAC chain whenever the limb above the chain may already hold
// data (accumulating a second product, or a reduction row), so no carry ever has to ripple.
struct Wide {
    uint32_t e[16];
    uint32_t o[16];
    uint32_t k[8];
};

IMT_HD void wide_zero(Wide& w) {
#pragma unroll
    for (int i = 0; i < 16; ++i) w.e[i] = 0, w.o[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) w.k[i] = 0;
}

// carry out of a chain whose next limb is at absolute limb position `pos`
template <bool FRESH, bool IS_E>
IMT_HD void chain_top(Wide& w, int pos) {
    if (pos >= 16) {  // the total is < 2^512, so this carry is always 0: just hand the (zero) flag on
        (void)cc::addc_cc(0, 0);
        return;
    }
    if (FRESH) {
        // the limb above the chain holds at most a few earlier carry bits: absorb the carry there
        if (IS_E) w.e[pos] = cc::addc_cc(w.e[pos], 0);
        else w.o[pos - 1] = cc::addc_cc(w.o[pos - 1], 0);
    } else {
        w.k[pos - 8] = cc::addc_cc(w.k[pos - 8], 0);
    }
}

// one row: w += (a[0..7] * s) << (32 * i).  `a` limbs with even index feed one accumulator, odd the other.
template <bool FRESH, int I>
IMT_HD void mac_row(Wide& w, const uint32_t* a, uint32_t s) {
    if constexpr ((I & 1) == 0) {
        cc::madw_head<cc::kHeadProdE>(w.e[I], w.e[I + 1], a[0], s);
        cc::madwc_cc(w.e[I + 2], w.e[I + 3], a[2], s);
        cc::madwc_cc(w.e[I + 4], w.e[I + 5], a[4], s);
        cc::madwc_cc(w.e[I + 6], w.e[I + 7], a[6], s);
        chain_top<FRESH, true>(w, I + 8);
        cc::madw_head<cc::kHeadProdO>(w.o[I], w.o[I + 1], a[1], s);
        cc::madwc_cc(w.o[I + 2], w.o[I + 3], a[3], s);
        cc::madwc_cc(w.o[I + 4], w.o[I + 5], a[5], s);
        cc::madwc_cc(w.o[I + 6], w.o[I + 7], a[7], s);
        chain_top<FRESH, false>(w, I + 9);
    } else {
        cc::madw_head<cc::kHeadProdO>(w.o[I - 1], w.o[I], a[0], s);
        cc::madwc_cc(w.o[I + 1], w.o[I + 2], a[2], s);
        cc::madwc_cc(w.o[I + 3], w.o[I + 4], a[4], s);
        cc::madwc_cc(w.o[I + 5], w.o[I + 6], a[6], s);
        chain_top<FRESH, false>(w, I + 8);
        cc::madw_head<cc::kHeadProdE>(w.e[I + 1], w.e[I + 2], a[1], s);
        cc::madwc_cc(w.e[I + 3], w.e[I + 4], a[3], s);
        cc::madwc_cc(w.e[I + 5], w.e[I + 6], a[5], s);
        cc::madwc_cc(w.e[I + 7], w.e[I + 8], a[7], s);
        chain_top<FRESH, true>(w, I + 9);
    }
}

// w = a * b (w must be zero on entry; ptxas folds the zero addends into RZ)
IMT_HD void mul_wide(Wide& w, const uint32_t* a, const uint32_t* b) {
    mac_row<true, 0>(w, a, b[0]);
    mac_row<true, 1>(w, a, b[1]);
    mac_row<true, 2>(w, a, b[2]);
    mac_row<true, 3>(w, a, b[3]);
    mac_row<true, 4>(w, a, b[4]);
    mac_row<true, 5>(w, a, b[5]);
    mac_row<true, 6>(w, a, b[6]);
    mac_row<true, 7>(w, a, b[7]);
}
// w += a * b (any w; carries that leave a chain are parked in w.k)
IMT_HD void mac_wide(Wide& w, const uint32_t* a, const uint32_t* b) {
    mac_row<false, 0>(w, a, b[0]);
    mac_row<false, 1>(w, a, b[1]);
    mac_row<false, 2>(w, a, b[2]);
    mac_row<false, 3>(w, a, b[3]);
    mac_row<false, 4>(w, a, b[4]);
    mac_row<false, 5>(w, a, b[5]);
    mac_row<false, 6>(w, a, b[6]);
    mac_row<false, 7>(w, a, b[7]);
}
// w += c * 2^256  (c is a field element; used to fold "+ constant" / "+ s_i" into a reduction)
IMT_HD void add_hi(Wide& w, const uint32_t* c) {
    w.e[8] = cc::add_cc(w.e[8], c[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) w.e[8 + i] = cc::addc_cc(w.e[8 + i], c[i]);
    // no carry out: the total stays < 2^512
}

// w = a^2 (w must be zero on entry). 28 off-diagonal + 8 diagonal wide MACs.
// On exit the value is entirely in w.e (w.o and w.k are zero again).
IMT_HD void sqr_wide(Wide& w, const uint32_t* a) {
    // ---- off-diagonal a_i * a_j, i < j, landing at limb position i + j
    // row 0
    cc::madw_head<cc::kHeadSqr>(w.o[0], w.o[1], a[0], a[1]);
    cc::madwc_cc(w.o[2], w.o[3], a[0], a[3]);
    cc::madwc_cc(w.o[4], w.o[5], a[0], a[5]);
    cc::madwc_cc(w.o[6], w.o[7], a[0], a[7]);
    w.o[8] = cc::addc_cc(w.o[8], 0);
    cc::madw_head<cc::kHeadSqr>(w.e[2], w.e[3], a[0], a[2]);
    cc::madwc_cc(w.e[4], w.e[5], a[0], a[4]);
    cc::madwc_cc(w.e[6], w.e[7], a[0], a[6]);
    w.e[8] = cc::addc_cc(w.e[8], 0);
    // row 1
    cc::madw_head<cc::kHeadSqr>(w.o[2], w.o[3], a[1], a[2]);
    cc::madwc_cc(w.o[4], w.o[5], a[1], a[4]);
    cc::madwc_cc(w.o[6], w.o[7], a[1], a[6]);
    w.o[8] = cc::addc_cc(w.o[8], 0);
    cc::madw_head<cc::kHeadSqr>(w.e[4], w.e[5], a[1], a[3]);
    cc::madwc_cc(w.e[6], w.e[7], a[1], a[5]);
    cc::madwc_cc(w.e[8], w.e[9], a[1], a[7]);
    w.e[10] = cc::addc_cc(w.e[10], 0);
    // row 2
    cc::madw_head<cc::kHeadSqr>(w.o[4], w.o[5], a[2], a[3]);
    cc::madwc_cc(w.o[6], w.o[7], a[2], a[5]);
    cc::madwc_cc(w.o[8], w.o[9], a[2], a[7]);
    w.o[10] = cc::addc_cc(w.o[10], 0);
    cc::madw_head<cc::kHeadSqr>(w.e[6], w.e[7], a[2], a[4]);
    cc::madwc_cc(w.e[8], w.e[9], a[2], a[6]);
    w.e[10] = cc::addc_cc(w.e[10], 0);
    // row 3
    cc::madw_head<cc::kHeadSqr>(w.o[6], w.o[7], a[3], a[4]);
    cc::madwc_cc(w.o[8], w.o[9], a[3], a[6]);
    w.o[10] = cc::addc_cc(w.o[10], 0);
    cc::madw_head<cc::kHeadSqr>(w.e[8], w.e[9], a[3], a[5]);
    cc::madwc_cc(w.e[10], w.e[11], a[3], a[7]);
    w.e[12] = cc::addc_cc(w.e[12], 0);
    // row 4
    cc::madw_head<cc::kHeadSqr>(w.o[8], w.o[9], a[4], a[5]);
    cc::madwc_cc(w.o[10], w.o[11], a[4], a[7]);
    w.o[12] = cc::addc_cc(w.o[12], 0);
    cc::madw_head<cc::kHeadSqr>(w.e[10], w.e[11], a[4], a[6]);
    w.e[12] = cc::addc_cc(w.e[12], 0);
    // row 5
    cc::madw_head<cc::kHeadSqr>(w.o[10], w.o[11], a[5], a[6]);
    w.o[12] = cc::addc_cc(w.o[12], 0);
    cc::madw_head<cc::kHeadSqr>(w.e[12], w.e[13], a[5], a[7]);
    w.e[14] = cc::addc_cc(w.e[14], 0);
    // row 6
    cc::madw_head<cc::kHeadSqr>(w.o[12], w.o[13], a[6], a[7]);
    w.o[14] = cc::addc_cc(w.o[14], 0);
    // ---- merge: e[pos] += o[pos-1]  (e[0] = e[1]'s pair is still zero; position 0 holds nothing)
    w.e[1] = cc::add_cc(w.e[1], w.o[0]);
#pragma unroll
    for (int pos = 2; pos < 16; ++pos) w.e[pos] = cc::addc_cc(w.e[pos], w.o[pos - 1]);
    // ---- double
    w.e[1] = cc::add_cc(w.e[1], w.e[1]);
#pragma unroll
    for (int pos = 2; pos < 16; ++pos) w.e[pos] = cc::addc_cc(w.e[pos], w.e[pos]);
    // ---- diagonal a_i^2 at position 2i: one chain over aligned pairs
    cc::madw_head<cc::kHeadSqr>(w.e[0], w.e[1], a[0], a[0]);
    cc::madwc_cc(w.e[2], w.e[3], a[1], a[1]);
    cc::madwc_cc(w.e[4], w.e[5], a[2], a[2]);
    cc::madwc_cc(w.e[6], w.e[7], a[3], a[3]);
    cc::madwc_cc(w.e[8], w.e[9], a[4], a[4]);
    cc::madwc_cc(w.e[10], w.e[11], a[5], a[5]);
    cc::madwc_cc(w.e[12], w.e[13], a[6], a[6]);
    cc::madwc_cc(w.e[14], w.e[15], a[7], a[7]);
#pragma unroll
    for (int i = 0; i < 16; ++i) w.o[i] = 0;
}

// one Montgomery reduction row: make limb I of the running total zero by adding m * p << (32 I)
template <int I>
IMT_HD void redc_row(Wide& w, uint32_t& cin) {
    uint32_t m;
    if constexpr ((I & 1) == 0) {
        // limb I lives in e[I] (home, aligned pair (e[I], e[I+1])) and o[I-1] (other)
        uint32_t t = w.e[I];
        uint32_t c1 = 0;
        if constexpr (I > 0) {
            t = cc::add_cc(t, w.o[I - 1]);
            c1 = cc::addc(0, 0);
            t = cc::add_cc(t, cin);
            c1 = cc::addc_cc(c1, 0);  // c1 <= 2: leaves the flag clear for the chain below
        }
        w.e[I] = t;
        cin = c1;
        m = t * IMT_INV32;
        cc::madw_head<cc::kHeadRedE>(w.e[I], w.e[I + 1], m, IMT_P0);
        cc::madwc_cc(w.e[I + 2], w.e[I + 3], m, IMT_P2);
        cc::madwc_cc(w.e[I + 4], w.e[I + 5], m, IMT_P4);
        cc::madwc_cc(w.e[I + 6], w.e[I + 7], m, IMT_P6);
        chain_top<false, true>(w, I + 8);
        cc::madw_head<cc::kHeadRedO>(w.o[I], w.o[I + 1], m, IMT_P1);
        cc::madwc_cc(w.o[I + 2], w.o[I + 3], m, IMT_P3);
        cc::madwc_cc(w.o[I + 4], w.o[I + 5], m, IMT_P5);
        cc::madwc_cc(w.o[I + 6], w.o[I + 7], m, IMT_P7);
        chain_top<false, false>(w, I + 9);
    } else {
        // limb I lives in o[I-1] (home, aligned pair (o[I-1], o[I])) and e[I] (other)
        uint32_t t = cc::add_cc(w.o[I - 1], w.e[I]);
        uint32_t c1 = cc::addc(0, 0);
        t = cc::add_cc(t, cin);
        c1 = cc::addc_cc(c1, 0);
        w.o[I - 1] = t;
        cin = c1;
        m = t * IMT_INV32;
        cc::madw_head<cc::kHeadRedO>(w.o[I - 1], w.o[I], m, IMT_P0);
        cc::madwc_cc(w.o[I + 1], w.o[I + 2], m, IMT_P2);
        cc::madwc_cc(w.o[I + 3], w.o[I + 4], m, IMT_P4);
        cc::madwc_cc(w.o[I + 5], w.o[I + 6], m, IMT_P6);
        chain_top<false, false>(w, I + 8);
        cc::madw_head<cc::kHeadRedE>(w.e[I + 1], w.e[I + 2], m, IMT_P1);
        cc::madwc_cc(w.e[I + 3], w.e[I + 4], m, IMT_P3);
        cc::madwc_cc(w.e[I + 5], w.e[I + 6], m, IMT_P5);
        cc::madwc_cc(w.e[I + 7], w.e[I + 8], m, IMT_P7);
        chain_top<false, true>(w, I + 9);
    }
}

// r = w / 2^256 mod p, r < w / 2^256 + p  (NOT conditionally reduced). w is consumed.
IMT_HD void redc(uint32_t* r, Wide& w) {
    uint32_t cin = 0;
    redc_row<0>(w, cin);
    redc_row<1>(w, cin);
    redc_row<2>(w, cin);
    redc_row<3>(w, cin);
    redc_row<4>(w, cin);
    redc_row<5>(w, cin);
    redc_row<6>(w, cin);
    redc_row<7>(w, cin);
    w.k[0] += cin;  // both are tiny
    r[0] = cc::add_cc(w.e[8], w.o[7]);
#pragma unroll
    for (int j = 1; j < 8; ++j) r[j] = cc::addc_cc(w.e[8 + j], w.o[7 + j]);
    r[0] = cc::add_cc(r[0], w.k[0]);
#pragma unroll
    for (int j = 1; j < 8; ++j) r[j] = cc::addc_cc(r[j], w.k[j]);
}

// ------------------------------------------------------------------------------------------ field helpers
// x in [0, 2m) -> [0, m) where m = (m0..m7)
#define IMT_COND_SUB(x, M0, M1, M2, M3, M4, M5, M6, M7)                 \
    do {                                                                \
        uint32_t d0 = cc::sub_cc((x)[0], M0);                           \
        uint32_t d1 = cc::subc_cc((x)[1], M1);                          \
        uint32_t d2 = cc::subc_cc((x)[2], M2);                          \
        uint32_t d3 = cc::subc_cc((x)[3], M3);                          \
        uint32_t d4 = cc::subc_cc((x)[4], M4);                          \
        uint32_t d5 = cc::subc_cc((x)[5], M5);                          \
        uint32_t d6 = cc::subc_cc((x)[6], M6);                          \
        uint32_t d7 = cc::subc_cc((x)[7], M7);                          \
        uint32_t bw = cc::subc(0, 0); /* 0xffffffff when x < m */       \
        if (bw == 0) {                                                  \
            (x)[0] = d0; (x)[1] = d1; (x)[2] = d2; (x)[3] = d3;         \
            (x)[4] = d4; (x)[5] = d5; (x)[6] = d6; (x)[7] = d7;         \
        }                                                               \
        cc::clear_after((x)[7]);                                        \
    } while (0)

IMT_HD void cond_sub_p(uint32_t* x) { IMT_COND_SUB(x, IMT_P0, IMT_P1, IMT_P2, IMT_P3, IMT_P4, IMT_P5, IMT_P6, IMT_P7); }
IMT_HD void cond_sub_2p(uint32_t* x) { IMT_COND_SUB(x, IMT_2P0, IMT_2P1, IMT_2P2, IMT_2P3, IMT_2P4, IMT_2P5, IMT_2P6, IMT_2P7); }

// semi-reduced [0,2p) -> canonical [0,p)
IMT_HD void canonicalize(uint32_t* x) { cond_sub_p(x); }

IMT_HD bool is_canonical(const uint32_t* x) {  // x < p ?
    (void)cc::sub_cc(x[0], IMT_P0);
    (void)cc::subc_cc(x[1], IMT_P1);
    (void)cc::subc_cc(x[2], IMT_P2);
    (void)cc::subc_cc(x[3], IMT_P3);
    (void)cc::subc_cc(x[4], IMT_P4);
    (void)cc::subc_cc(x[5], IMT_P5);
    (void)cc::subc_cc(x[6], IMT_P6);
    (void)cc::subc_cc(x[7], IMT_P7);
    const bool below = cc::subc(0, 0) != 0;
    cc::clear();
    return below;
}

// r = a * b * R^-1, semi-reduced in -> semi-reduced out (no conditional subtraction needed: 4p < 2^256)
IMT_HD void mont_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    Wide w;
    wide_zero(w);
    mul_wide(w, a, b);
    redc(r, w);
}
IMT_HD void mont_sqr(uint32_t* r, const uint32_t* a) {
    Wide w;
    wide_zero(w);
    sqr_wide(w, a);
    redc(r, w);
}
// r = a + b, inputs semi-reduced, output semi-reduced
IMT_HD void add_semi(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    r[0] = cc::add_cc(a[0], b[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) r[i] = cc::addc_cc(a[i], b[i]);
    cond_sub_2p(r);  // a + b < 4p < 2^256
}

// R^2 mod p (to Montgomery form) — same constant halo2curves uses
#define IMT_R2_LIMBS {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u}

IMT_HD void to_mont(uint32_t* r, const uint32_t* canon) {
    const uint32_t r2[8] = IMT_R2_LIMBS;
    mont_mul(r, canon, r2);
}
// Montgomery (semi-reduced) -> canonical integer
IMT_HD void from_mont(uint32_t* r, const uint32_t* a) {
    Wide w;
    wide_zero(w);
#pragma unroll
    for (int i = 0; i < 8; ++i) w.e[i] = a[i];
    redc(r, w);  // < 2p/2^256 + p
    cond_sub_p(r);
}

}  // namespace imt

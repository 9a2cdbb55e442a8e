// TEST-ONLY: compiles the product's field/Poseidon headers in HOST mode (emulated carry flag) so that the exact
// algorithm the GPU kernels run can be checked against big-integer arithmetic without a GPU.
// Built on demand by tests/test_host_field.py into tests/_build/; never part of the shipped library.
#include <cstring>

#include "poseidon_params.h"
#include "poseidon_lh_math.cuh"

using namespace imt;

static PoseidonParams g_params;
static bool g_ready = false;
static const PoseidonParams& params() {
    if (!g_ready) {
        poseidon_params_generate(&g_params);
        g_ready = true;
    }
    return g_params;
}

extern "C" {
void shim_params(void* out) { std::memcpy(out, &params(), sizeof(PoseidonParams)); }
unsigned shim_params_size() { return sizeof(PoseidonParams); }
void shim_mont_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) { mont_mul(r, a, b); }
void shim_mont_sqr(uint32_t* r, const uint32_t* a) { mont_sqr(r, a); }
void shim_to_mont(uint32_t* r, const uint32_t* a) { to_mont(r, a); }
void shim_from_mont(uint32_t* r, const uint32_t* a) { from_mont(r, a); }
void shim_add_semi(uint32_t* r, const uint32_t* a, const uint32_t* b) { add_semi(r, a, b); }
void shim_dot3(uint32_t* r, const uint32_t* u, const uint32_t* m) { dot3(r, u, u + 8, u + 16, m, m + 8, m + 16); }
void shim_mul_add(uint32_t* r, const uint32_t* u, const uint32_t* m, const uint32_t* s) { mul_add(r, u, m, s); }
void shim_sqr_wide(uint32_t* out16, const uint32_t* a) {
    Wide w;
    wide_zero(w);
    sqr_wide(w, a);
    std::memcpy(out16, w.e, 64);
}
void shim_mul_wide_merged(uint32_t* out16, const uint32_t* a, const uint32_t* b) {
    Wide w;
    wide_zero(w);
    mul_wide(w, a, b);
    uint64_t c = 0;
    for (int pos = 0; pos < 16; ++pos) {
        c += (uint64_t)w.e[pos] + (pos ? w.o[pos - 1] : 0);
        out16[pos] = (uint32_t)c;
        c >>= 32;
    }
}
// The 57 partial rounds of a permutation on a canonical state (3 x 8 words in / out): which = 0 the product's partial_round
// (poseidon.cuh), which = 1 the lead / helper recurrence with its derived tables (poseidon_lh_math.cuh: what k_hash_lh runs)
void shim_partial_rounds(uint32_t* state, int which) {
    uint32_t s[3][8];
    for (int i = 0; i < 3; ++i) {
        to_mont(s[i], state + 8 * i);
        cond_sub_p(s[i]);
    }
    const PoseidonParams& P = params();
    if (which == 0) {
        NoTrace nt;
        for (int k = 0; k < kRP; ++k) partial_round(s, P.partial[k], nt);
    } else {
        static LhAux aux;
        static bool made = false;
        if (!made) {
            for (int k = 0; k < kRP; ++k) lh_make_round(&aux.round[k], P.partial[k], k + 1 < kRP ? &P.partial[k + 1] : nullptr);
            lh_make_kc0(&aux.kc0, P.partial[0]);
            made = true;
        }
        lh_partial_rounds(s, P, aux);
    }
    for (int i = 0; i < 3; ++i) from_mont(state + 8 * i, s[i]);
}
struct Collect {
    uint32_t* dst;
    uint32_t* sbox = nullptr;  // optional: the extended S-box trace, (x^2, x^4, x^5 + c) per S-box
    void emit(const uint32_t (*s)[8]) {
        for (int i = 0; i < 3; ++i) {
            uint32_t t[8];
            for (int j = 0; j < 8; ++j) t[j] = s[i][j];
            from_mont(dst, t);
            dst += 8;
        }
    }
    void emit_sbox(const uint32_t* x2, const uint32_t* x4, const uint32_t* u) {
        if (!sbox) return;
        const uint32_t* v[3] = {x2, x4, u};
        for (int i = 0; i < 3; ++i) {
            uint32_t t[8];
            for (int j = 0; j < 8; ++j) t[j] = v[i][j];
            from_mont(sbox, t);
            sbox += 8;
        }
    }
};
// canonical in (arity x 8 words) -> canonical digest; states (132 x 3 x 8 words, canonical) optional;
// sbox (162 x 3 x 8 words, canonical) optional, only with states
void shim_hash_ext(uint32_t* out, const uint32_t* in, int arity, uint32_t* states, uint32_t* sbox);
void shim_hash(uint32_t* out, const uint32_t* in, int arity, uint32_t* states) { shim_hash_ext(out, in, arity, states, nullptr); }
void shim_hash_ext(uint32_t* out, const uint32_t* in, int arity, uint32_t* states, uint32_t* sbox) {
    uint32_t m[3][8], d[8];
    for (int i = 0; i < arity; ++i) {
        to_mont(m[i], in + 8 * i);
        cond_sub_p(m[i]);
    }
    if (states) {
        Collect c{states, sbox};
        if (arity == 3) hash_fixed<3>(d, m, params(), c); else hash_fixed<2>(d, m, params(), c);
    } else {
        NoTrace n;
        if (arity == 3) hash_fixed<3>(d, m, params(), n); else hash_fixed<2>(d, m, params(), n);
    }
    from_mont(out, d);
}
}  // extern "C"

// any-width instance <t, t-1>(r_f, r_p): canonical in (arity x 8 words) -> canonical digest; states optional
// ((arity / (t-1) + 1) x (1 + r_f + r_p) x t x 8 words, canonical). Returns 0, or 1 for an unsupported instance.
template <int T>
struct CollectT {
    uint32_t* dst;
    void emit(const uint32_t (*s)[8]) {
        if (!dst) return;
        for (int i = 0; i < T; ++i) {
            uint32_t t[8];
            for (int j = 0; j < 8; ++j) t[j] = s[i][j];
            from_mont(dst, t);
            dst += 8;
        }
    }
    void emit_sbox(const uint32_t*, const uint32_t*, const uint32_t*) {}
};
struct HostLoad {
    const uint32_t* in;
    void operator()(size_t j, uint32_t* x) {
        to_mont(x, in + 8 * j);
        cond_sub_p(x);
    }
};
template <int T>
static void spec_hash_t(uint32_t* out, const uint32_t* in, size_t arity, const Fr* P, SpecLayout L, uint32_t* states) {
    uint32_t d[8];
    HostLoad load{in};
    CollectT<T> sink{states};
    spec_sponge<T>(d, arity, load, P, L, sink);
    from_mont(out, d);
}
extern "C" int shim_spec_hash(unsigned t, unsigned r_f, unsigned r_p, const uint32_t* in, size_t arity, uint32_t* out, uint32_t* states) {
    if (t < kSpecMinT || t > kSpecMaxT) return 1;
    const SpecLayout L{t, r_f, r_p};
    static Fr params[4096];
    if (L.total() > 4096 || !poseidon_spec_generate(t, r_f, r_p, params)) return 1;
    switch (t) {
        case 2: spec_hash_t<2>(out, in, arity, params, L, states); break;
        case 3: spec_hash_t<3>(out, in, arity, params, L, states); break;
        case 4: spec_hash_t<4>(out, in, arity, params, L, states); break;
        default: spec_hash_t<5>(out, in, arity, params, L, states); break;
    }
    return 0;
}

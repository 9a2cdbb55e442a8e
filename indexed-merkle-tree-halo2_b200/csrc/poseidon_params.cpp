// Host-side derivation of the Poseidon parameter set, done once per context — the native analogue of what
// `Poseidon::<Fr,3,2>::new(8, 57)` does at construction (/root/reference/src/indexed_merkle_tree.rs:370, 663,
// 681, 807; the generator itself lives in the un-vendored pse-poseidon dependency, Cargo.toml:16):
//   Grain LFSR -> 65 x 3 round constants (rejection sampled) + Cauchy MDS from 3 + 3 further elements
//   -> "optimized" constants (partial-round constants pushed through MDS^-1 into a single lane)
//   -> factorisation of the 57 partial-round linear layers into sparse matrices + one dense pre-matrix.
// Runs a few thousand field operations on the host through fr.cuh's emulated carry primitives; the result is
// uploaded to __constant__ memory. Nothing here is on the hashing path.
#include "poseidon_params.h"

#include <cstring>
#include <utility>
#include <vector>

namespace imt {
namespace {

struct F {
    uint32_t l[8];
};  // Montgomery, canonical

const F kZero = {{0, 0, 0, 0, 0, 0, 0, 0}};

F mul(const F& a, const F& b) {
    F r;
    mont_mul(r.l, a.l, b.l);
    cond_sub_p(r.l);
    return r;
}
F add(const F& a, const F& b) {
    F r;
    r.l[0] = cc::add_cc(a.l[0], b.l[0]);
    for (int i = 1; i < 8; ++i) r.l[i] = cc::addc_cc(a.l[i], b.l[i]);
    cond_sub_p(r.l);
    return r;
}
F neg(const F& a) {
    const uint32_t p[8] = {IMT_P0, IMT_P1, IMT_P2, IMT_P3, IMT_P4, IMT_P5, IMT_P6, IMT_P7};
    bool zero = true;
    for (int i = 0; i < 8; ++i) zero = zero && a.l[i] == 0;
    if (zero) return a;
    F r;
    r.l[0] = cc::sub_cc(p[0], a.l[0]);
    for (int i = 1; i < 8; ++i) r.l[i] = cc::subc_cc(p[i], a.l[i]);
    return r;
}
F sub(const F& a, const F& b) { return add(a, neg(b)); }
bool is_zero(const F& a) {
    uint32_t x = 0;
    for (int i = 0; i < 8; ++i) x |= a.l[i];
    return x == 0;
}
F from_canonical(const uint32_t* c) {
    F r;
    to_mont(r.l, c);
    cond_sub_p(r.l);
    return r;
}
F from_u64(uint64_t v) {
    uint32_t c[8] = {(uint32_t)v, (uint32_t)(v >> 32), 0, 0, 0, 0, 0, 0};
    return from_canonical(c);
}
F inverse(const F& a) {  // a^(p-2)
    const uint32_t e[8] = {IMT_P0 - 2, IMT_P1, IMT_P2, IMT_P3, IMT_P4, IMT_P5, IMT_P6, IMT_P7};
    F r = from_u64(1);
    for (int i = 255; i >= 0; --i) {
        r = mul(r, r);
        if ((e[i / 32] >> (i % 32)) & 1) r = mul(r, a);
    }
    return r;
}

// ---- Grain LFSR in self-shrinking mode (Poseidon paper, parameter generation)
class Grain {
  public:
    Grain(unsigned field_bits, unsigned t, unsigned r_f, unsigned r_p) {
        unsigned n = 0;
        auto push = [&](unsigned width, uint32_t v) {
            for (int i = (int)width - 1; i >= 0; --i) reg_[n++] = (v >> i) & 1u;
        };
        push(2, 1);  // GF(p)
        push(4, 0);  // x^alpha
        push(12, field_bits);
        push(12, t);
        push(10, r_f);
        push(10, r_p);
        push(30, 0x3fffffffu);
        for (int i = 0; i < 160; ++i) step();
    }
    // 254 filtered bits, most significant first, as 8 little-endian 32-bit limbs
    void draw254(uint32_t* out) {
        std::memset(out, 0, 32);
        for (int bit = 253; bit >= 0; --bit)
            if (filtered()) out[bit / 32] |= 1u << (bit % 32);
    }

  private:
    uint8_t reg_[80];
    unsigned head_ = 0;
    unsigned at(unsigned k) const { return reg_[(head_ + k) % 80]; }
    unsigned step() {
        unsigned b = at(62) ^ at(51) ^ at(38) ^ at(23) ^ at(13) ^ at(0);
        reg_[head_] = (uint8_t)b;
        head_ = (head_ + 1) % 80;
        return b;
    }
    unsigned filtered() {  // bits come in pairs; the second is kept only when the first is 1
        for (;;) {
            unsigned first = step(), second = step();
            if (first) return second;
        }
    }
};

bool below_p(const uint32_t* x) { return is_canonical(x); }

F draw_rejecting(Grain& g) {
    uint32_t w[8];
    do g.draw254(w);
    while (!below_p(w));
    return from_canonical(w);
}
F draw_reducing(Grain& g) {  // 254-bit value < 2p
    uint32_t w[8];
    g.draw254(w);
    cond_sub_p(w);
    return from_canonical(w);
}

struct M3 {
    F v[3][3];
};
M3 transpose(const M3& a) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.v[j][i] = a.v[i][j];
    return r;
}
M3 matmul(const M3& a, const M3& b) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            F acc = kZero;
            for (int k = 0; k < 3; ++k) acc = add(acc, mul(a.v[i][k], b.v[k][j]));
            r.v[i][j] = acc;
        }
    return r;
}
void matvec(const M3& m, const F* x, F* y) {
    F t[3];
    for (int i = 0; i < 3; ++i) {
        F acc = kZero;
        for (int j = 0; j < 3; ++j) acc = add(acc, mul(m.v[i][j], x[j]));
        t[i] = acc;
    }
    for (int i = 0; i < 3; ++i) y[i] = t[i];
}
// inverse via the adjugate (3x3) — cofactor expansion
M3 inverse3(const M3& m) {
    auto minor2 = [&](int r0, int r1, int c0, int c1) {
        return sub(mul(m.v[r0][c0], m.v[r1][c1]), mul(m.v[r0][c1], m.v[r1][c0]));
    };
    M3 cof;
    const int o[3][2] = {{1, 2}, {0, 2}, {0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            F d = minor2(o[i][0], o[i][1], o[j][0], o[j][1]);
            cof.v[i][j] = ((i + j) & 1) ? neg(d) : d;
        }
    F det = kZero;
    for (int j = 0; j < 3; ++j) det = add(det, mul(m.v[0][j], cof.v[0][j]));
    F dinv = inverse(det);
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.v[i][j] = mul(cof.v[j][i], dinv);
    return r;
}

void store(Fr& dst, const F& src) { std::memcpy(dst.l, src.l, 32); }

}  // namespace

void poseidon_params_generate(PoseidonParams* out) {
    constexpr int kRounds = kRF + kRP;
    Grain grain(254, kT, kRF, kRP);
    static F rc[kRounds][3];
    for (int r = 0; r < kRounds; ++r)
        for (int i = 0; i < 3; ++i) rc[r][i] = draw_rejecting(grain);
    F xs[3], ys[3];
    for (int i = 0; i < 3; ++i) xs[i] = draw_reducing(grain);
    for (int i = 0; i < 3; ++i) ys[i] = draw_reducing(grain);
    M3 mds;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) mds.v[i][j] = inverse(add(xs[i], ys[j]));  // Cauchy matrix
    const M3 mds_inv = inverse3(mds);

    std::memset(out, 0, sizeof(*out));
    // ---- constants, re-associated so that every round adds its constants AFTER the S-box
    for (int i = 0; i < 3; ++i) store(out->pre[i], rc[0][i]);
    F v[3];
    for (int r = 1; r < kHalfF; ++r) {  // full rounds 0..2 of the first half
        matvec(mds_inv, rc[r], v);
        for (int i = 0; i < 3; ++i) store(out->full[r - 1][i], v[i]);
    }
    // partial rounds, walked backwards: only lane 0 keeps a constant, the rest is pushed one round earlier
    F acc[3] = {rc[kHalfF + kRP][0], rc[kHalfF + kRP][1], rc[kHalfF + kRP][2]};
    for (int k = kRP - 1; k >= 0; --k) {
        matvec(mds_inv, acc, v);
        store(out->partial[k].c, v[0]);
        v[0] = kZero;
        for (int i = 0; i < 3; ++i) acc[i] = add(v[i], rc[kHalfF + k][i]);
    }
    matvec(mds_inv, acc, v);
    for (int i = 0; i < 3; ++i) store(out->full[kHalfF - 1][i], v[i]);
    for (int r = 0; r < kHalfF - 1; ++r) {  // full rounds 4..6; round 7 adds nothing
        matvec(mds_inv, rc[kHalfF + kRP + 1 + r], v);
        for (int i = 0; i < 3; ++i) store(out->full[kHalfF + r][i], v[i]);
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) store(out->mds[i][j], mds.v[i][j]);

    // ---- sparse factorisation of the partial-round linear layers, last round first:
    //      A = M^T;  A = [[a00, a0*], [w, Mh]]  ->  sparse = (row = [a00, Mh^-1 w], col = a0*),  A <- M^T * diag(1, Mh)
    const M3 mt = transpose(mds);
    M3 a = mt;
    const F one = from_u64(1);
    for (int k = kRP - 1; k >= 0; --k) {
        const F h00 = a.v[1][1], h01 = a.v[1][2], h10 = a.v[2][1], h11 = a.v[2][2];
        const F dinv = inverse(sub(mul(h00, h11), mul(h01, h10)));
        const F w0 = a.v[1][0], w1 = a.v[2][0];
        // Mh^-1 = 1/det * [[h11, -h01], [-h10, h00]]
        const F wh0 = mul(dinv, sub(mul(h11, w0), mul(h01, w1)));
        const F wh1 = mul(dinv, sub(mul(h00, w1), mul(h10, w0)));
        store(out->partial[k].row[0], a.v[0][0]);
        store(out->partial[k].row[1], wh0);
        store(out->partial[k].row[2], wh1);
        store(out->partial[k].col[0], a.v[0][1]);
        store(out->partial[k].col[1], a.v[0][2]);
        M3 mp;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) mp.v[i][j] = kZero;
        mp.v[0][0] = one;
        mp.v[1][1] = h00, mp.v[1][2] = h01, mp.v[2][1] = h10, mp.v[2][2] = h11;
        a = matmul(mt, mp);
    }
    const M3 pre = transpose(a);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) store(out->pre_sparse[i][j], pre.v[i][j]);

    store(out->one, one);
    const uint32_t cap[8] = {0, 0, 1, 0, 0, 0, 0, 0};  // 2^64
    store(out->cap, from_canonical(cap));
}

// ------------------------------------------------------------------------------------------------- any width
// The same derivation for any instance `Poseidon::<Fr, T, RATE>::new(r_f, r_p)` — the reference's tree and chip are
// generic over T and RATE (/root/reference/src/utils.rs:6, 19; src/indexed_merkle_tree.rs:65, 127, 231) although only
// <3, 2>(8, 57) is instantiated (indexed_merkle_tree.rs:362-365). Output: one dense Fr array laid out by SpecLayout.
namespace {

struct Mat {
    int n;
    std::vector<F> v;  // row-major
    explicit Mat(int n_) : n(n_), v((size_t)n_ * n_, kZero) {}
    F& at(int i, int j) { return v[(size_t)i * n + j]; }
    const F& at(int i, int j) const { return v[(size_t)i * n + j]; }
};
Mat transpose(const Mat& a) {
    Mat r(a.n);
    for (int i = 0; i < a.n; ++i)
        for (int j = 0; j < a.n; ++j) r.at(j, i) = a.at(i, j);
    return r;
}
Mat matmul(const Mat& a, const Mat& b) {
    Mat r(a.n);
    for (int i = 0; i < a.n; ++i)
        for (int j = 0; j < a.n; ++j) {
            F acc = kZero;
            for (int k = 0; k < a.n; ++k) acc = add(acc, mul(a.at(i, k), b.at(k, j)));
            r.at(i, j) = acc;
        }
    return r;
}
std::vector<F> matvec(const Mat& m, const std::vector<F>& x) {
    std::vector<F> y((size_t)m.n, kZero);
    for (int i = 0; i < m.n; ++i)
        for (int j = 0; j < m.n; ++j) y[i] = add(y[i], mul(m.at(i, j), x[j]));
    return y;
}
// Gauss-Jordan; false when singular
bool invert(const Mat& m, Mat* out) {
    const int n = m.n;
    Mat a = m, b(n);
    const F one = from_u64(1);
    for (int i = 0; i < n; ++i) b.at(i, i) = one;
    for (int c = 0; c < n; ++c) {
        int piv = c;
        while (piv < n && is_zero(a.at(piv, c))) ++piv;
        if (piv == n) return false;
        for (int j = 0; j < n; ++j) std::swap(a.at(c, j), a.at(piv, j)), std::swap(b.at(c, j), b.at(piv, j));
        const F k = inverse(a.at(c, c));
        for (int j = 0; j < n; ++j) a.at(c, j) = mul(a.at(c, j), k), b.at(c, j) = mul(b.at(c, j), k);
        for (int r = 0; r < n; ++r) {
            if (r == c || is_zero(a.at(r, c))) continue;
            const F f = a.at(r, c);
            for (int j = 0; j < n; ++j) {
                a.at(r, j) = sub(a.at(r, j), mul(f, a.at(c, j)));
                b.at(r, j) = sub(b.at(r, j), mul(f, b.at(c, j)));
            }
        }
    }
    *out = b;
    return true;
}

}  // namespace

bool poseidon_spec_generate(unsigned t_, unsigned r_f_, unsigned r_p_, Fr* out) {
    const int t = (int)t_, r_f = (int)r_f_, r_p = (int)r_p_, half = r_f / 2, rounds = r_f + r_p;
    const SpecLayout L{t_, r_f_, r_p_};
    Grain grain(254, t_, r_f_, r_p_);
    std::vector<std::vector<F>> rc((size_t)rounds, std::vector<F>((size_t)t));
    for (int r = 0; r < rounds; ++r)
        for (int i = 0; i < t; ++i) rc[r][i] = draw_rejecting(grain);
    std::vector<F> xs((size_t)t), ys((size_t)t);
    for (int i = 0; i < t; ++i) xs[i] = draw_reducing(grain);
    for (int i = 0; i < t; ++i) ys[i] = draw_reducing(grain);
    Mat mds(t), mds_inv(t);
    for (int i = 0; i < t; ++i)
        for (int j = 0; j < t; ++j) {
            const F d = add(xs[i], ys[j]);
            if (is_zero(d)) return false;
            mds.at(i, j) = inverse(d);  // Cauchy matrix
        }
    if (!invert(mds, &mds_inv)) return false;

    for (size_t i = 0; i < L.total(); ++i) std::memset(out[i].l, 0, 32);
    const F one = from_u64(1);
    const uint32_t cap[8] = {0, 0, 1, 0, 0, 0, 0, 0};  // 2^64
    store(out[L.cap()], from_canonical(cap));
    store(out[L.one()], one);
    // ---- constants, re-associated so that every round adds its constants AFTER the S-box
    for (int i = 0; i < t; ++i) store(out[L.pre(i)], rc[0][i]);
    for (int r = 1; r < half; ++r) {
        const std::vector<F> v = matvec(mds_inv, rc[r]);
        for (int i = 0; i < t; ++i) store(out[L.full(r - 1, i)], v[i]);
    }
    std::vector<F> acc = rc[half + r_p];
    for (int k = r_p - 1; k >= 0; --k) {
        std::vector<F> v = matvec(mds_inv, acc);
        store(out[L.partial_c(k)], v[0]);
        v[0] = kZero;
        for (int i = 0; i < t; ++i) acc[i] = add(v[i], rc[half + k][i]);
    }
    {
        const std::vector<F> v = matvec(mds_inv, acc);
        for (int i = 0; i < t; ++i) store(out[L.full(half - 1, i)], v[i]);
    }
    for (int r = 0; r < half - 1; ++r) {  // second half; the last full round adds nothing
        const std::vector<F> v = matvec(mds_inv, rc[half + r_p + 1 + r]);
        for (int i = 0; i < t; ++i) store(out[L.full(half + r, i)], v[i]);
    }
    for (int i = 0; i < t; ++i)
        for (int j = 0; j < t; ++j) store(out[L.mds(i, j)], mds.at(i, j));
    // ---- sparse factorisation, last partial round first (see poseidon_params_generate)
    const Mat mt = transpose(mds);
    Mat a = mt;
    for (int k = r_p - 1; k >= 0; --k) {
        Mat mh(t - 1), mh_inv(t - 1);
        std::vector<F> w((size_t)(t - 1));
        for (int i = 1; i < t; ++i) {
            w[i - 1] = a.at(i, 0);
            for (int j = 1; j < t; ++j) mh.at(i - 1, j - 1) = a.at(i, j);
        }
        if (t > 1 && !invert(mh, &mh_inv)) return false;
        const std::vector<F> wh = matvec(mh_inv, w);
        store(out[L.partial_row(k, 0)], a.at(0, 0));
        for (int i = 1; i < t; ++i) store(out[L.partial_row(k, i)], wh[i - 1]);
        for (int i = 1; i < t; ++i) store(out[L.partial_col(k, i - 1)], a.at(0, i));
        Mat mp(t);
        mp.at(0, 0) = one;
        for (int i = 1; i < t; ++i)
            for (int j = 1; j < t; ++j) mp.at(i, j) = mh.at(i - 1, j - 1);
        a = matmul(mt, mp);
    }
    const Mat pre = transpose(a);
    for (int i = 0; i < t; ++i)
        for (int j = 0; j < t; ++j) store(out[L.pre_sparse(i, j)], pre.at(i, j));
    return true;
}


}  // namespace imt

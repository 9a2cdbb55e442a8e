// Latency-oriented Poseidon, lead / helper form: the S-box chain of a hash runs in a warp of its OWN.
//
// What a lone warp pays is the NUMBER of IMAD.WIDE in its instruction stream: one warp issues one every ~7 cycles whatever
// the dependencies, and the four sub-partitions of an SM issue independently (tools/lab/issue_probe.cu, profiles/
// r02_latency_lab.md section 5). poseidon_coop.cuh puts the three state lanes of a hash into ONE warp, so the warp that
// carries the S-box also carries the linear layer: x^2, x^4, (x^4 x + c | s_i row_i), (u row_0 + P | u col_i + s_i) = 456
// multiplies per partial round. Here the partial rounds (57 of the 65 rounds of a permutation) are split by ROLE between warps
// on different sub-partitions. Every output of a partial round is affine in u = x^5 + c:
//     x'   = row_0 u + row_1 s_1 + row_2 s_2 = x^4 (row_0 x) + K          K   = row_0 c + row_1 s_1 + row_2 s_2
//     s_i' = col_i u + s_i                   = x^4 (col_i x) + (col_i c + s_i)
//     K'   = row_0' c' + row_1' s_1' + row_2' s_2' = x^4 (rho x) + (kappa + row_1' s_1 + row_2' s_2)
//            rho = row_1' col_1 + row_2' col_2,  kappa = rho c + row_0' c'      (' = the next round; rho, kappa, col_i c: tables)
// so the LEAD warp (one lane per hash) runs only   x^2 ; x^4 ; x' = x^4 y_0 + K    (two real squarings + one fused multiply-add
// = 328 multiplies per round, the minimum multiplicative depth of x^5), and a HELPER warp (8 lanes per hash) runs beside it
//     slot A (while the lead squares):  y_0 = row_0 x | w_1 = col_1 x | w_2 = col_2 x | z = rho x | d_1 = row_1' s_1 | d_2 = row_2' s_2
//     slot C (beside the lead's last product, needs x^4):  s_1' = x^4 w_1 + (col_1 c + s_1) | s_2' = ... | K' = x^4 z + (kappa + d_1 + d_2)
// = 256 multiplies per round (y_0 is the one product the lead waits for, ~200 cycles a round: lab section 6). Operands cross warps
// through shared memory at three points per round
// (x, x^4 from the lead; y_0, K from the helpers) with named barriers: the producer ARRIVES and goes on, only the consumer waits.
// The 8 full rounds run on the helper warps as in poseidon_coop.cuh (three lanes, one S-box each). A block = 1 lead warp + 3 helper
// warps = 12 hashes, one warp per sub-partition; used while all hashes in flight fit one block per SM (imt_latency.cu).
// Results are the same field elements as the other kernels after canonicalisation, bit for bit (the representatives in [0, 2p)
// differ in between); imt_ctx_create cross-checks the three families on every context.
#pragma once
#include "poseidon_coop.cuh"
#include "poseidon_lh_math.cuh"

// tools/lab/lh_prof.cu defines IMT_LH_PROF and reads where the cycles of a round go; the library build has no hooks
#ifdef IMT_LH_PROF
__device__ long long g_lh_prof[32];
#define LH_PROF_DECL long long lhp_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, lhp_t_ = clock64()
#define LH_PROF(i)                         \
    do {                                   \
        const long long now_ = clock64();  \
        lhp_[i] += now_ - lhp_t_;          \
        lhp_t_ = now_;                     \
    } while (0)
#define LH_PROF_OUT(off, cond)                                      \
    do {                                                            \
        if (cond)                                                   \
            for (int i_ = 0; i_ < 8; ++i_) g_lh_prof[(off) + i_] = lhp_[i_]; \
    } while (0)
#else
#define LH_PROF_DECL (void)0
#define LH_PROF(i) (void)0
#define LH_PROF_OUT(off, cond) (void)0
#endif

namespace imt {

constexpr int kLhSlots = 12;  // hashes per block: 3 helper warps x 4 groups of 8 lanes; lanes 0..11 of the lead warp
constexpr int kLhThreads = 128;

// the tables above, once per context (the arithmetic is lh_make_round of poseidon_lh_math.cuh, which the CPU tests run too)
__global__ void k_lh_aux(const PoseidonParams* __restrict__ G, LhAux* __restrict__ A) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= kRP) return;
    lh_make_round(&A->round[k], G->partial[k], k + 1 < kRP ? &G->partial[k + 1] : nullptr);
    if (k == 0) lh_make_kc0(&A->kc0, G->partial[0]);
}

// Named barriers of a block (0 is __syncthreads): the producer side arrives and goes on, the consumer side waits; 128 threads take part
// in each. Safe under ARBITRARY delays of any warp (a co-running kernel may slow a helper warp by thousands of cycles), not just
// under the usual timing: a barrier may only be re-armed by a warp that knows every participant has LEFT its previous phase.
//   kBarX   lead arrives for round k + 1 after it passed YK_k  <- every helper arrived at YK_k <- after it passed X_k
//   kBarYK  helpers arrive for round k + 1 after they passed X_{k+1} <- the lead arrived at X_{k+1} <- after it passed YK_k
//   kBarX4  the lead arrives for round k + 1 after it passed YK_k only, and a helper arrives at YK_k BEFORE it waits at X4_k: two
//           phases of one barrier could overlap (and x^4 of round k + 1 overwrite x^4 of round k before a late helper has read it),
//           so this barrier and its buffer alternate by round parity: phase k + 2 follows YK_{k+1}, which every helper reaches
//           only after it passed X4_k and loaded x^4 of round k.
//           (Round 56 of the first permutation and round 0 of the second share a parity: the lead arms the latter after kBarInit,
//           where every helper arrives after its 57 rounds.)
// The same chains order every shared-memory buffer (a writer of round k + 1 runs after every reader of round k).
constexpr int kBarInit = 1;  // helpers -> lead: s_0 at the start of the partial rounds
constexpr int kBarX = 2;     // lead -> helpers: x of the round (and the final s_0)
constexpr int kBarYK = 3;    // helpers -> lead: y_0 and K
constexpr int kBarX4 = 4;    // lead -> helpers: x^4; ids 4 and 5 by round parity
template <int ID>
__device__ __forceinline__ void bar_sync() {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(kLhThreads) : "memory");
}
template <int ID>
__device__ __forceinline__ void bar_arrive() {
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(kLhThreads) : "memory");
}
__device__ __forceinline__ void bar_sync_id(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kLhThreads) : "memory"); }
__device__ __forceinline__ void bar_arrive_id(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kLhThreads) : "memory"); }
// one field element per hash slot, word-major (the lead's lanes write consecutive banks, a helper group reads one address)
typedef uint32_t LhBuf[8][kLhSlots];
__device__ __forceinline__ void lh_store(LhBuf& b, int slot, const uint32_t* x) {
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i][slot] = x[i];
}
__device__ __forceinline__ void lh_load(uint32_t* x, const LhBuf& b, int slot) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = b[i][slot];
}

// the pre-add and the first four full rounds / the last four full rounds, three lanes per hash as in permute_coop
template <bool FIRST>
__device__ __forceinline__ void lh_full_rounds(uint32_t* x, const PoseidonParams* __restrict__ G, int rr, int base) {
    if (FIRST) {
        uint32_t c[8];
        ld_fe(c, &G->pre[rr]);
        add_semi(x, x, c);
    }
#pragma unroll 1
    for (int fr = FIRST ? 0 : kHalfF; fr < (FIRST ? kHalfF : kRF); ++fr) {
        uint32_t c[8], u0[8], u1[8], u2[8], m0[8], m1[8], m2[8];
        ld_fe(c, &G->full[fr][rr]);
        sbox_add(x, x, c);
        shfl_fe(u0, x, base);
        shfl_fe(u1, x, base + 1);
        shfl_fe(u2, x, base + 2);
        const Fr(*m)[3] = (FIRST && fr == kHalfF - 1) ? G->pre_sparse : G->mds;
        ld_fe(m0, &m[rr][0]);
        ld_fe(m1, &m[rr][1]);
        ld_fe(m2, &m[rr][2]);
        dot3(x, u0, u1, u2, m0, m1, m2);
    }
}

// out[h] = H(in[ARITY*h .. ARITY*h + ARITY)): same contract as k_hash_coop, for batches of at most one block per SM
template <int ARITY>
__global__ void __launch_bounds__(kLhThreads) k_hash_lh(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n, int in_fmt, int out_fmt,
                                                        const PoseidonParams* __restrict__ G, const LhAux* __restrict__ A,
                                                        uint32_t* __restrict__ err) {
    __shared__ LhBuf sX, sX4[2], sY, sK;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        // ------------------------------------------------------------------ lead: lane = hash slot; x^2, x^4, x^4 y_0 + K
        const bool act = lane < kLhSlots;
        const int slot = act ? lane : kLhSlots - 1;  // the other lanes shadow the last slot (they load, never store)
        uint32_t x[8];
        LH_PROF_DECL;
#pragma unroll 1
        for (int perm = 0; perm < 2; ++perm) {
            bar_sync<kBarInit>();
            lh_load(x, sX, slot);
            LH_PROF(0);  // waiting for the helpers' full rounds
#pragma unroll 1
            for (int k = 0; k < kRP; ++k) {
                uint32_t x2[8], x4[8], y[8], kk[8];
                if (act) lh_store(sX, slot, x);
                bar_arrive<kBarX>();
                LH_PROF(1);
                mont_sqr(x2, x);
                mont_sqr(x4, x2);
                LH_PROF(2);
                if (act) lh_store(sX4[k & 1], slot, x4);
                bar_arrive_id(kBarX4 + (k & 1));
                LH_PROF(3);
                bar_sync<kBarYK>();
                LH_PROF(4);
                lh_load(y, sY, slot);
                lh_load(kk, sK, slot);
                mul_add(x, x4, y, kk);
                LH_PROF(5);
            }
            if (act) lh_store(sX, slot, x);
            bar_arrive<kBarX>();
        }
        LH_PROF_OUT(0, blockIdx.x == 0 && lane == 0);
        return;
    }
    // ---------------------------------------------------------------------- helpers: 8 lanes per hash
    const int role = lane & 7, base = lane & ~7;
    const int slot = (warp - 1) * 4 + (lane >> 3);
    const int rr = role < 3 ? role : 2;  // state element this lane carries through the full rounds (roles >= 3 mirror role 2)
    const size_t h = (size_t)blockIdx.x * kLhSlots + slot;
    const size_t hc = h < n ? h : n - 1;  // slots past the end recompute the last hash: every lane reaches the shuffles and barriers
    const bool side = role == 1 || role == 2;
    uint32_t x[8], second[8];
    bool ok = true;
    if (rr == 0) {
        ld_fe(x, &G->cap);
    } else {
        load_fe(x, in + 2 * (ARITY * hc + (rr == 1 ? 0 : 1)));
        ok &= ingest(x, in_fmt);
    }
    if (ARITY == 3 && rr == 1) {
        load_fe(second, in + 2 * (ARITY * hc + 2));
        ok &= ingest(second, in_fmt);
    } else {
        ld_fe(second, &G->one);
        const bool pad_here = ARITY == 3 ? rr == 2 : rr == 1;
#pragma unroll
        for (int i = 0; i < 8; ++i) second[i] = pad_here ? second[i] : 0u;
    }
    if (!ok) atomicOr(err, kErrNonCanonical);
    LH_PROF_DECL;
#pragma unroll 1
    for (int perm = 0; perm < 2; ++perm) {
        if (perm) add_semi(x, x, second);
        lh_full_rounds<true>(x, G, rr, base);
        LH_PROF(0);  // full rounds
        // ---- partial rounds
        if (role == 0) lh_store(sX, slot, x);
        bar_arrive<kBarInit>();
        uint32_t S[8], K[8];  // S: this lane's s_i (roles 1, 2; zero elsewhere)   K: the K of the coming round (role 3)
#pragma unroll
        for (int i = 0; i < 8; ++i) S[i] = side ? x[i] : 0u;
        {  // first K = row_0 c + row_1 s_1 + row_2 s_2 with the rows of partial round 0
            uint32_t sh[8], m[8], t[8], d1[8], d2[8], c[8];
            shfl_fe(sh, S, role == 4 ? base + 1 : base + 2);
            ld_fe(m, role == 4 ? &G->partial[0].row[1] : &G->partial[0].row[2]);
            mont_mul(t, sh, m);
            shfl_fe(d1, t, base + 4);
            shfl_fe(d2, t, base + 5);
            ld_fe(c, &A->kc0);
            add_semi(K, c, d1);
            add_semi(K, K, d2);
        }
        // the tables of round k + 1 are fetched during round k: a first touch comes from L2 (~700 cycles), and slot A sits on the
        // critical path of the lead (its y_0 is what the lead's last product waits for)
        uint32_t m[8], c[8];
        ld_fe(m, &A->round[0].mul_a[role]);
        ld_fe(c, &A->round[0].add_c[role < 4 ? role : 0]);
#pragma unroll 1
        for (int k = 0; k < kRP; ++k) {
            const LhRound* nx = &A->round[k + 1 < kRP ? k + 1 : k];
            uint32_t sh[8], X[8], ta[8], mn[8], cn[8];
            ld_fe(mn, &nx->mul_a[role]);
            ld_fe(cn, &nx->add_c[role < 4 ? role : 0]);
            shfl_fe(sh, S, role == 4 ? base + 1 : base + 2);  // role 4 takes s_1, role 5 takes s_2
            LH_PROF(1);
            bar_sync<kBarX>();
            LH_PROF(2);
            lh_load(X, sX, slot);
#pragma unroll
            for (int i = 0; i < 8; ++i) X[i] = role < 4 ? X[i] : sh[i];
            mont_mul(ta, X, m);  // slot A
            if (role == 0) lh_store(sY, slot, ta);
            if (role == 3) lh_store(sK, slot, K);
            bar_arrive<kBarYK>();
            LH_PROF(3);
            uint32_t d1[8], d2[8], add[8], X4[8];
            shfl_fe(d1, ta, base + 4);
            shfl_fe(d2, ta, base + 5);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                d1[i] = role == 3 ? d1[i] : S[i];
                d2[i] = role == 3 ? d2[i] : 0u;
            }
            add_semi(add, c, d1);  // role 3: kappa + d_1 + d_2     roles 1, 2: col_i c + s_i
            add_semi(add, add, d2);
            LH_PROF(4);
            bar_sync_id(kBarX4 + (k & 1));
            LH_PROF(5);
            lh_load(X4, sX4[k & 1], slot);
            mul_add(K, X4, ta, add);  // slot C: role 1, 2: s_i'   role 3: the next K
            LH_PROF(6);
#pragma unroll
            for (int i = 0; i < 8; ++i) S[i] = side ? K[i] : 0u, m[i] = mn[i], c[i] = cn[i];
        }
        bar_sync<kBarX>();
        {
            uint32_t X[8], s2[8];
            lh_load(X, sX, slot);
            shfl_fe(s2, S, base + 2);
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = role == 0 ? X[i] : (side ? S[i] : s2[i]);
        }
        lh_full_rounds<false>(x, G, rr, base);
        LH_PROF(0);
    }
    LH_PROF_OUT(8, blockIdx.x == 0 && warp == 1 && lane == 0);
    if (role == 1 && h < n) {
        canonicalize(x);
        egress(x, out_fmt);
        store_fe(out + 2 * h, x);
    }
}

}  // namespace imt

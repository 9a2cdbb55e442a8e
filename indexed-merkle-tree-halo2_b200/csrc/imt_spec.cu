// Any-width Poseidon instances: `Poseidon::<Fr, T, RATE>::new(r_f, r_p)` with T = 2..5, RATE = T - 1, any r_f / r_p and any
// input length — the generics of /root/reference/src/utils.rs:6, 19 and src/indexed_merkle_tree.rs:65, 127, 231.
// The reference's own instance <3, 2>(8, 57) keeps its tuned kernels (imt_capi.cu); a context made by imt_ctx_create_spec
// routes every hash of the library (tree levels, leaf hashing, folds, traces, insert levels, sharding cap) through the
// kernels below instead, so the whole C-ABI works for any instance. One thread = one hash; T is a template parameter,
// everything else is a run-time value read from the SpecLayout-ordered parameter array in global memory.
#include <cuda_runtime.h>

#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "imt_b200.h"
#include "imt_internal.h"
#include "kernels_common.cuh"
#include "poseidon_params.h"
#include "poseidon_spec.cuh"

using namespace imt;
using namespace imt_host;

namespace imt {

struct SpecNoTrace {
    __device__ __forceinline__ void emit(const uint32_t (*)[8]) {}
    __device__ __forceinline__ void emit_sbox(const uint32_t*, const uint32_t*, const uint32_t*) {}
};
// every traced state = T FE in the user format, hash after hash: perms x (1 + r_f + r_p) x T FE, contiguous per hash
// `sbox` (may be null): the extended trace, (x^2, x^4, x^5 + c) of every S-box in execution order
template <int T>
struct SpecTraceSink {
    uint4* dst;
    int fmt;
    uint4* sbox = nullptr;
    __device__ __forceinline__ void put(uint4*& p, const uint32_t* x) {
        uint32_t t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = x[i];
        if (fmt == kFmtCanonical) from_mont(t, t);
        else canonicalize(t);
        store_fe(p, t);
        p += 2;
    }
    __device__ __forceinline__ void emit(const uint32_t (*s)[8]) {
#pragma unroll
        for (int j = 0; j < T; ++j) put(dst, s[j]);
    }
    __device__ __forceinline__ void emit_sbox(const uint32_t* x2, const uint32_t* x4, const uint32_t* u) {
        if (sbox) {
            put(sbox, x2);
            put(sbox, x4);
            put(sbox, u);
        }
    }
};
// sponge inputs straight from global memory (validated, converted at the edge)
struct SpecGlobalLoad {
    const uint4* base;
    int fmt;
    bool ok;
    __device__ __forceinline__ void operator()(size_t j, uint32_t* x) {
        load_fe(x, base + 2 * j);
        ok &= ingest(x, fmt);
    }
};
// sponge inputs = the two children of a tree node, in registers
struct SpecPairLoad {
    const uint32_t* a;
    const uint32_t* b;
    __device__ __forceinline__ void operator()(size_t j, uint32_t* x) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = j == 0 ? a[i] : b[i];
    }
};

// out[i] = H(in[arity*i .. arity*i + arity)); `states` (optional) receives the witness trace of every hash
// occupancy hint per width (same reasoning as k_hash, kernels.cuh): blocks of 128 threads per SM
#ifndef IMT_SPEC_MIN_BLOCKS
#define IMT_SPEC_MIN_BLOCKS(T) ((T) <= 2 ? 7 : (T) == 3 ? 6 : (T) == 4 ? 5 : 4)
#endif
template <int T>
__global__ void __launch_bounds__(kHashThreads, IMT_SPEC_MIN_BLOCKS(T)) k_spec_hash(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n, size_t arity,
                                                            const Fr* __restrict__ P, SpecLayout L, int in_fmt, int out_fmt,
                                                            uint4* __restrict__ states, size_t state_fe, uint32_t* __restrict__ err,
                                                            uint4* __restrict__ sbox, size_t sbox_fe) {
    const size_t i = blockIdx.x * (size_t)kHashThreads + threadIdx.x;
    if (i >= n) return;
    SpecGlobalLoad load{in + 2 * arity * i, in_fmt, true};
    uint32_t d[8];
    if (states) {
        SpecTraceSink<T> sink{states + 2 * state_fe * i, out_fmt, sbox ? sbox + 2 * sbox_fe * i : nullptr};
        spec_sponge<T>(d, arity, load, P, L, sink);
    } else {
        SpecNoTrace nt;
        spec_sponge<T>(d, arity, load, P, L, nt);
    }
    if (!load.ok) atomicOr(err, kErrNonCanonical);
    egress(d, out_fmt);
    if (out) store_fe(out + 2 * i, d);
}

// the bare permutation on n states of T FE (published permutation test vectors; not used by the tree)
template <int T>
__global__ void __launch_bounds__(kHashThreads) k_spec_permute(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n,
                                                               const Fr* __restrict__ P, SpecLayout L, int fmt, uint32_t* __restrict__ err) {
    const size_t i = blockIdx.x * (size_t)kHashThreads + threadIdx.x;
    if (i >= n) return;
    uint32_t s[T][8];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < T; ++j) {
        load_fe(s[j], in + 2 * (T * i + j));
        ok &= ingest(s[j], fmt);
    }
    if (!ok) atomicOr(err, kErrNonCanonical);
    SpecNoTrace nt;
    spec_permute<T>(s, P, L, nt);
#pragma unroll
    for (int j = 0; j < T; ++j) {
        canonicalize(s[j]);
        egress(s[j], fmt);
        store_fe(out + 2 * (T * i + j), s[j]);
    }
}

// batched verify_proof / compute_merkle_root (utils.rs:87-107, indexed_merkle_tree.rs:78-96): one thread folds one path
template <int T>
__global__ void __launch_bounds__(kHashThreads) k_spec_fold(const uint4* __restrict__ leaves, const uint64_t* __restrict__ idx,
                                                            const uint4* __restrict__ siblings, const uint4* __restrict__ roots, size_t q,
                                                            unsigned depth, const Fr* __restrict__ P, SpecLayout L, int fmt,
                                                            uint8_t* __restrict__ ok_out, uint4* __restrict__ roots_out,
                                                            uint4* __restrict__ states, size_t state_fe, uint32_t* __restrict__ err) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= q) return;
    uint32_t h[8];
    load_fe(h, leaves + 2 * i);
    bool ok = ingest(h, fmt);
    uint64_t index = idx[i];
#pragma unroll 1
    for (unsigned l = 0; l < depth; ++l) {
        uint32_t s[8], lo[8], hi[8];
        load_fe(s, siblings + 2 * (i * depth + l));
        ok &= ingest(s, fmt);
        const bool left = (index & 1) == 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            lo[k] = left ? h[k] : s[k];
            hi[k] = left ? s[k] : h[k];
        }
        SpecPairLoad load{lo, hi};
        if (states) {
            SpecTraceSink<T> sink{states + 2 * state_fe * (i * depth + l), fmt};
            spec_sponge<T>(h, 2, load, P, L, sink);
        } else {
            SpecNoTrace nt;
            spec_sponge<T>(h, 2, load, P, L, nt);
        }
        index >>= 1;
    }
    canonicalize(h);
    if (ok_out) {
        uint32_t r[8];
        load_fe(r, roots + 2 * i);
        ok &= ingest(r, fmt);
        canonicalize(r);
        bool same = true;
#pragma unroll
        for (int k = 0; k < 8; ++k) same &= r[k] == h[k];
        ok_out[i] = (uint8_t)same;
    }
    if (!ok) atomicOr(err, kErrNonCanonical);
    if (roots_out) {
        egress(h, fmt);
        store_fe(roots_out + 2 * i, h);
    }
}

// k_trace_tree_paths (kernels.cuh) for any width: the traced hashes of the paths of a resident tree are independent
template <int T>
__global__ void __launch_bounds__(kHashThreads) k_spec_trace_tree_paths(const uint4* __restrict__ levels, const uint4* __restrict__ cap,
                                                                        size_t n_local, unsigned depth_local, unsigned cap_depth, unsigned rank,
                                                                        const uint64_t* __restrict__ idx, size_t q, const Fr* __restrict__ P,
                                                                        SpecLayout L, int fmt, uint4* __restrict__ states, size_t state_fe,
                                                                        uint32_t* __restrict__ err, uint4* __restrict__ sbox, size_t sbox_fe) {
    const unsigned depth = depth_local + cap_depth;
    const size_t t = blockIdx.x * (size_t)kHashThreads + threadIdx.x;
    if (t >= q * depth) return;
    const size_t qi = t / depth;
    const unsigned lvl = (unsigned)(t % depth);
    const uint64_t g = idx[qi];
    const uint64_t base = (uint64_t)rank * n_local;
    if (g < base || g >= base + n_local) {
        atomicOr(err, kErrIndexOob);
        return;
    }
    const uint4* src;
    if (lvl < depth_local) src = levels + 2 * (level_offset(n_local, lvl) + (((g - base) >> lvl) & ~(uint64_t)1));
    else src = cap + 2 * (level_offset((size_t)1 << cap_depth, lvl - depth_local) + (((uint64_t)rank >> (lvl - depth_local)) & ~(uint64_t)1));
    uint32_t lo[8], hi[8], d[8];
    load_fe(lo, src);
    load_fe(hi, src + 2);
    SpecPairLoad load{lo, hi};
    SpecTraceSink<T> sink{states + 2 * state_fe * t, fmt, sbox ? sbox + 2 * sbox_fe * t : nullptr};
    spec_sponge<T>(d, 2, load, P, L, sink);
}

}  // namespace imt

namespace imt_host {

imt_status ensure_spec(imt_ctx* ctx) {
    if (ctx->d_spec) return IMT_OK;
    const SpecLayout L = ctx->spec;
    std::vector<Fr> host(L.total());
    if (!poseidon_spec_generate(L.t, L.r_f, L.r_p, host.data())) return fail(ctx, IMT_ERR_INVALID_ARG, "parameter generation hit a singular matrix");
    IMT_TRY_CUDA(ctx, cudaMalloc((void**)&ctx->d_spec, host.size() * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, cudaMemcpy(ctx->d_spec, host.data(), host.size() * sizeof(Fr), cudaMemcpyHostToDevice));
    return IMT_OK;
}

#define IMT_SPEC_DISPATCH(T_, CALL)   \
    switch (T_) {                     \
        case 2: { constexpr int T = 2; CALL; } break; \
        case 3: { constexpr int T = 3; CALL; } break; \
        case 4: { constexpr int T = 4; CALL; } break; \
        case 5: { constexpr int T = 5; CALL; } break; \
        default: return fail(ctx, IMT_ERR_INVALID_ARG, "unsupported width"); \
    }

imt_status launch_spec_hash(imt_ctx* ctx, size_t arity, const void* d_in, void* d_out, size_t n, int in_fmt, int out_fmt, void* d_states,
                            cudaStream_t s, void* d_sbox) {
    if (n == 0) return IMT_OK;
    IMT_TRY(ensure_spec(ctx));
    const SpecLayout L = ctx->spec;
    const size_t state_fe = trace_fe_per_hash(ctx, arity);
    imt_ctx::Timed tm{nullptr, nullptr, arity == 3 ? 3 : 2, n};
    const bool timed = ctx->timing && (arity == 2 || arity == 3);
    if (timed) {
        IMT_TRY_CUDA(ctx, cudaEventCreate(&tm.a));
        IMT_TRY_CUDA(ctx, cudaEventCreate(&tm.b));
        IMT_TRY_CUDA(ctx, cudaEventRecord(tm.a, s));
    }
    IMT_SPEC_DISPATCH(L.t, (k_spec_hash<T><<<grid_for(n, kHashThreads), kHashThreads, 0, s>>>(
                               (const uint4*)d_in, (uint4*)d_out, n, arity, ctx->d_spec, L, in_fmt, out_fmt, (uint4*)d_states, state_fe, ctx->d_err,
                               (uint4*)d_sbox, sbox_fe_per_hash(ctx, arity))));
    ++ctx->launches;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    if (timed) {
        IMT_TRY_CUDA(ctx, cudaEventRecord(tm.b, s));
        ctx->pending.push_back(tm);
    }
    return IMT_OK;
}

imt_status launch_spec_fold(imt_ctx* ctx, const void* d_leaves, const uint64_t* d_indices, const void* d_roots, const void* d_siblings, size_t q,
                            unsigned depth, uint8_t* d_ok, void* d_roots_out, void* d_states) {
    if (q == 0) return IMT_OK;
    IMT_TRY(ensure_spec(ctx));
    const SpecLayout L = ctx->spec;
    const size_t state_fe = trace_fe_per_hash(ctx, 2);
    const unsigned threads = q >= (size_t)1 << 20 ? kHashThreads : 32;  // small batches: one warp per block spreads over the SMs
    IMT_SPEC_DISPATCH(L.t, (k_spec_fold<T><<<grid_for(q, threads), threads, 0, ctx->stream>>>(
                               (const uint4*)d_leaves, d_indices, (const uint4*)d_siblings, (const uint4*)d_roots, q, depth, ctx->d_spec, L,
                               ctx->fmt, d_ok, (uint4*)d_roots_out, (uint4*)d_states, state_fe, ctx->d_err)));
    ++ctx->launches;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    return IMT_OK;
}

imt_status launch_spec_tree_trace(imt_tree* t, const uint64_t* d_idx, size_t q, void* d_states, void* d_sbox) {
    imt_ctx* ctx = t->ctx;
    IMT_TRY(ensure_spec(ctx));
    const SpecLayout L = ctx->spec;
    const unsigned cap_depth = t->cap_valid ? t->cap_depth : 0;
    const unsigned depth = t->depth + cap_depth;
    const size_t state_fe = trace_fe_per_hash(ctx, 2);
    IMT_SPEC_DISPATCH(L.t, (k_spec_trace_tree_paths<T><<<grid_for(q * depth, kHashThreads), kHashThreads, 0, ctx->stream>>>(
                               (const uint4*)t->d_levels, (const uint4*)t->d_cap, t->n, t->depth, cap_depth, t->cap_valid ? t->rank : 0u,
                               d_idx, q, ctx->d_spec, L, ctx->fmt, (uint4*)d_states, state_fe, ctx->d_err, (uint4*)d_sbox,
                               sbox_fe_per_hash(ctx, 2))));
    ++ctx->launches;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    return IMT_OK;
}

}  // namespace imt_host

// ------------------------------------------------------------------------------------------------- C-ABI
static bool spec_args_ok(unsigned t, unsigned rate, unsigned r_f, unsigned r_p) {
    return t >= kSpecMinT && t <= kSpecMaxT && rate + 1 == t && r_f >= 2 && (r_f & 1) == 0 && r_f + r_p <= kSpecMaxRounds;
}

extern "C" imt_status imt_ctx_create_spec(int device, imt_fe_format format, unsigned t, unsigned rate, unsigned r_f, unsigned r_p, imt_ctx** out) {
    if (!out) return IMT_ERR_INVALID_ARG;
    *out = nullptr;
    if (!spec_args_ok(t, rate, r_f, r_p)) return IMT_ERR_INVALID_ARG;
    imt_ctx* ctx = nullptr;
    IMT_TRY(imt_ctx_create(device, format, &ctx));
    ctx->spec = SpecLayout{t, r_f, r_p};
    ctx->generic = true;
    const imt_status st = ensure_spec(ctx);
    if (st != IMT_OK) {
        imt_ctx_destroy(ctx);
        return st;
    }
    *out = ctx;
    return IMT_OK;
}

extern "C" imt_status imt_ctx_spec(const imt_ctx* ctx, unsigned* t, unsigned* rate, unsigned* r_f, unsigned* r_p, int* generic_kernels) {
    if (!ctx) return IMT_ERR_INVALID_ARG;
    if (t) *t = ctx->spec.t;
    if (rate) *rate = ctx->spec.t - 1;
    if (r_f) *r_f = ctx->spec.r_f;
    if (r_p) *r_p = ctx->spec.r_p;
    if (generic_kernels) *generic_kernels = ctx->generic ? 1 : 0;
    return IMT_OK;
}

extern "C" imt_status imt_trace_fe_per_hash(const imt_ctx* ctx, size_t arity, size_t* fe) {
    if (!ctx || !fe) return IMT_ERR_INVALID_ARG;
    *fe = trace_fe_per_hash(ctx, arity);
    return IMT_OK;
}
extern "C" imt_status imt_trace_sbox_fe_per_hash(const imt_ctx* ctx, size_t arity, size_t* fe) {
    if (!ctx || !fe) return IMT_ERR_INVALID_ARG;
    *fe = sbox_fe_per_hash(ctx, arity);
    return IMT_OK;
}

// host-only: the derived parameter array of an instance (no device work; CPU tests compare it with the oracle's)
extern "C" imt_status imt_spec_params_host(unsigned t, unsigned rate, unsigned r_f, unsigned r_p, void* out, size_t capacity_fe, size_t* count_fe) {
    if (!spec_args_ok(t, rate, r_f, r_p)) return IMT_ERR_INVALID_ARG;
    const SpecLayout L{t, r_f, r_p};
    if (count_fe) *count_fe = L.total();
    if (!out) return IMT_OK;
    if (capacity_fe < L.total()) return IMT_ERR_INVALID_ARG;
    return poseidon_spec_generate(t, r_f, r_p, static_cast<Fr*>(out)) ? IMT_OK : IMT_ERR_INVALID_ARG;
}

// the tuned kernels serve input lengths 2 and 3 of the default instance; everything else takes the any-width kernels
static imt_status hash_any_launch(imt_ctx* ctx, size_t arity, const void* d_in, void* d_out, size_t n, void* d_states) {
    if (!ctx->generic && !d_states && (arity == 2 || arity == 3))
        return launch_hash(ctx, (int)arity, d_in, d_out, n, ctx->fmt, ctx->fmt, ctx->stream);
    return launch_spec_hash(ctx, arity, d_in, d_out, n, ctx->fmt, ctx->fmt, d_states, ctx->stream);
}

extern "C" imt_status imt_poseidon_hash_dev(imt_ctx* ctx, const void* d_in, size_t arity, size_t n, void* d_out) {
    if (!ctx) return IMT_ERR_INVALID_ARG;
    if (n && ((arity && !d_in) || !d_out)) return fail(ctx, IMT_ERR_INVALID_ARG, "null buffer");
    if (n == 0) return IMT_OK;
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    IMT_TRY(clear_err(ctx));
    IMT_TRY(hash_any_launch(ctx, arity, d_in, d_out, n, nullptr));
    return finish(ctx);
}

extern "C" imt_status imt_poseidon_hash(imt_ctx* ctx, const void* in, size_t arity, size_t n, void* out) {
    if (!ctx) return IMT_ERR_INVALID_ARG;
    if (n && ((arity && !in) || !out)) return fail(ctx, IMT_ERR_INVALID_ARG, "null buffer");
    if (n == 0) return IMT_OK;
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf din(ctx), dout(ctx);
    IMT_TRY_CUDA(ctx, din.alloc(n * arity * sizeof(Fr)));
    IMT_TRY_CUDA(ctx, dout.alloc(n * sizeof(Fr)));
    if (arity) IMT_TRY_CUDA(ctx, cudaMemcpyAsync(din.p, in, n * arity * sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY(clear_err(ctx));
    IMT_TRY(hash_any_launch(ctx, arity, din.p, dout.p, n, nullptr));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(out, dout.p, n * sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
    return finish(ctx);
}

extern "C" imt_status imt_poseidon_permute(imt_ctx* ctx, const void* in_states, size_t n, void* out_states) {
    if (!ctx) return IMT_ERR_INVALID_ARG;
    if (n && (!in_states || !out_states)) return fail(ctx, IMT_ERR_INVALID_ARG, "null buffer");
    if (n == 0) return IMT_OK;
    IMT_TRY_CUDA(ctx, cudaSetDevice(ctx->device));
    IMT_TRY(ensure_spec(ctx));
    const SpecLayout L = ctx->spec;
    DevBuf din(ctx), dout(ctx);
    const size_t bytes = n * L.t * sizeof(Fr);
    IMT_TRY_CUDA(ctx, din.alloc(bytes));
    IMT_TRY_CUDA(ctx, dout.alloc(bytes));
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(din.p, in_states, bytes, cudaMemcpyHostToDevice, ctx->stream));
    IMT_TRY(clear_err(ctx));
    IMT_SPEC_DISPATCH(L.t, (k_spec_permute<T><<<grid_for(n, kHashThreads), kHashThreads, 0, ctx->stream>>>(
                               din.as<uint4>(), dout.as<uint4>(), n, ctx->d_spec, L, ctx->fmt, ctx->d_err)));
    ++ctx->launches;
    IMT_TRY_CUDA(ctx, cudaGetLastError());
    IMT_TRY_CUDA(ctx, cudaMemcpyAsync(out_states, dout.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return finish(ctx);
}

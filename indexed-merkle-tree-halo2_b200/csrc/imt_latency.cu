// The LATENCY kernels of the library (3 lanes per hash, poseidon_coop.cuh) in their own translation unit, because they want
// another carry discipline than the thread-per-hash kernels of imt_capi.cu (-DIMT_FREE_MASK, see fr.cuh): the levels near the
// root of a tree, the levels of an insert batch, small hash batches and small path folds run ONE warp per scheduler, where what
// counts is the length of a single instruction stream: ONE warp issues an IMAD.WIDE every ~7 cycles whatever its dependencies
// (tools/lab/issue_probe.cu; ptxas schedules for 4), and with every chain strung through the carry flag the IADD3.X / predicate
// instructions around the multiplies do not hide in the idle issue slots between them (tools/lab/latency_lab.cu). A chain whose
// head does not read the flag may be moved across the previous one. All 32 head masks were swept on a B200
// (profiles/r02_latency_lab.md): one level of <= 4096 nodes takes 283 us with every chain serialised (round 1), 233 us with mask
// 29 (here), 240 us with mask 22 (the best mask for the thread-per-hash kernels, where 7 warps share a scheduler and the carry
// predicates of too many open chains spill: mask 31 = everything free costs them 8 %).
// Same field elements out as the throughput kernels, bit for bit (tests/test_gpu_parity.py runs both on every tree).
#include <cuda_runtime.h>

#include "imt_b200.h"
#include "imt_internal.h"
#include "kernels_common.cuh"
#include "poseidon_coop.cuh"
#include "poseidon_lh.cuh"

#include <cstdlib>

using namespace imt;

namespace imt_host {

// this unit's copy of the parameters in constant memory (the trace sinks' conversions use it)
cudaError_t latency_upload_params(const PoseidonParams* host_params) { return cudaMemcpyToSymbol(c_params, host_params, sizeof(PoseidonParams)); }

}  // namespace imt_host

namespace imt_host {

cudaError_t latency_setup(imt_ctx* ctx) {
    cudaError_t e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, ctx->device);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_lh_aux, sizeof(LhAux));
    if (e != cudaSuccess) return e;
    k_lh_aux<<<1, 64, 0, ctx->stream>>>(ctx->d_params, (LhAux*)ctx->d_lh_aux);
    e = cudaGetLastError();
    return e == cudaSuccess ? cudaStreamSynchronize(ctx->stream) : e;
}
void latency_teardown(imt_ctx* ctx) {
    if (ctx->d_lh_aux) cudaFree(ctx->d_lh_aux);
    ctx->d_lh_aux = nullptr;
}

void launch_hash_latency(imt_ctx* ctx, int which, int arity, const void* d_in, void* d_out, size_t n, int in_fmt, int out_fmt, cudaStream_t s) {
    if (which == 1) {
        const unsigned grid = grid_for(n, kLhSlots);
        if (arity == 3)
            k_hash_lh<3><<<grid, kLhThreads, 0, s>>>((const uint4*)d_in, (uint4*)d_out, n, in_fmt, out_fmt, ctx->d_params, (const LhAux*)ctx->d_lh_aux,
                                                    ctx->d_err);
        else
            k_hash_lh<2><<<grid, kLhThreads, 0, s>>>((const uint4*)d_in, (uint4*)d_out, n, in_fmt, out_fmt, ctx->d_params, (const LhAux*)ctx->d_lh_aux,
                                                    ctx->d_err);
        return;
    }
    const unsigned grid = grid_for(4 * n, 128);
    if (arity == 3)
        k_hash_coop<3><<<grid, 128, 0, s>>>((const uint4*)d_in, (uint4*)d_out, n, in_fmt, out_fmt, ctx->d_params, ctx->d_err);
    else
        k_hash_coop<2><<<grid, 128, 0, s>>>((const uint4*)d_in, (uint4*)d_out, n, in_fmt, out_fmt, ctx->d_params, ctx->d_err);
}

// The lead / helper kernel wants a sub-partition per warp: it is chosen while every hash in flight fits ONE block (12 hashes) per SM —
// 1776 on a B200 — and the 3-lanes-per-hash kernel above that. IMT_LH_MAX_NODES overrides the bound (0 switches the kernel off;
// tuning / A-B measurements only).
void launch_hash_coop(imt_ctx* ctx, int arity, const void* d_in, void* d_out, size_t n, int in_fmt, int out_fmt, cudaStream_t s,
                      unsigned concurrent) {
    static const long long forced = [] {
        const char* e = std::getenv("IMT_LH_MAX_NODES");
        return e ? std::atoll(e) : -1ll;
    }();
    const size_t lh_max = forced >= 0 ? (size_t)forced : (size_t)kLhSlots * (size_t)ctx->sm_count;
    launch_hash_latency(ctx, ctx->d_lh_aux && n * concurrent <= lh_max ? 1 : 0, arity, d_in, d_out, n, in_fmt, out_fmt, s);
}

void launch_fold_coop(imt_ctx* ctx, const void* d_leaves, const uint64_t* d_indices, const void* d_roots, const void* d_siblings, size_t q,
                      unsigned depth, uint8_t* d_ok, void* d_roots_out, void* d_states) {
    k_fold_paths_coop<<<grid_for(4 * q, 128), 128, 0, ctx->stream>>>((const uint4*)d_leaves, d_indices, (const uint4*)d_siblings,
                                                                     (const uint4*)d_roots, q, depth, ctx->fmt, d_ok, (uint4*)d_roots_out,
                                                                     (uint4*)d_states, ctx->d_params, ctx->d_err);
}

}  // namespace imt_host

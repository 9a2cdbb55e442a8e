// The LATENCY kernels of the library in their own translation unit, compiled with FREE carry chains (-DIMT_FREE_MASK, see
// fr.cuh): the levels near the root of a tree, the levels of an insert batch, small hash batches and small path folds run one
// warp per scheduler, where what counts is the dependent-issue latency of a single instruction stream, not the pipe. The
// throughput kernels (imt_capi.cu) string every multiply-accumulate chain of a thread through the carry flag — right when
// eight warps share a scheduler; here the same field source is compiled so that a chain does not wait for the previous one:
// ptxas overlaps the even / odd accumulator chains of a row and the reduction with the product. Measured on one warp
// (tools/lab/latency_lab.cu, profiles/r02_latency_lab.md): a dependent IMAD.WIDE.X through the carry predicate issues every 7.05
// cycles, through the accumulator every 3.35; Montgomery product 1018 -> 844 cycles, squaring 902 -> 712; one tree level of
// <= 4096 nodes 283 -> 233 us (all 32 chain-head masks swept; bit-exact digests under every one).
// Same field elements out as the throughput kernels, bit for bit (tests/test_gpu_parity.py runs both on every tree).
#include <cuda_runtime.h>

#include "imt_b200.h"
#include "imt_internal.h"
#include "kernels_common.cuh"
#include "poseidon_coop.cuh"

using namespace imt;

namespace imt_host {

// this unit's copy of the parameters in constant memory (thread-per-hash latency variant below)
cudaError_t latency_upload_params(const PoseidonParams* host_params) { return cudaMemcpyToSymbol(c_params, host_params, sizeof(PoseidonParams)); }

}  // namespace imt_host

namespace {

// One thread per hash, free carry chains: tree levels too large for the 3-lanes-per-hash kernel but too small to fill the
// schedulers (8192 < nodes <= kLatencyMaxNodes: at most ~2 warps per scheduler). Same contract as k_hash (kernels.cuh).
template <int ARITY>
__global__ void __launch_bounds__(kHashThreads) k_hash_lat(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n, int in_fmt, int out_fmt,
                                                           uint32_t* __restrict__ err) {
    const size_t i = blockIdx.x * (size_t)kHashThreads + threadIdx.x;
    if (i >= n) return;
    uint32_t x[ARITY][8], d[8];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < ARITY; ++j) {
        load_fe(x[j], in + 2 * (ARITY * i + j));
        ok &= ingest(x[j], in_fmt);
    }
    if (!ok) atomicOr(err, kErrNonCanonical);
    NoTrace nt;
    hash_fixed<ARITY>(d, x, c_params, nt);
    egress(d, out_fmt);
    store_fe(out + 2 * i, d);
}

}  // namespace

namespace imt_host {

void launch_hash_coop(imt_ctx* ctx, int arity, const void* d_in, void* d_out, size_t n, int in_fmt, int out_fmt, cudaStream_t s) {
    const unsigned grid = grid_for(4 * n, 128);
    if (arity == 3)
        k_hash_coop<3><<<grid, 128, 0, s>>>((const uint4*)d_in, (uint4*)d_out, n, in_fmt, out_fmt, ctx->d_params, ctx->d_err);
    else
        k_hash_coop<2><<<grid, 128, 0, s>>>((const uint4*)d_in, (uint4*)d_out, n, in_fmt, out_fmt, ctx->d_params, ctx->d_err);
}

void launch_hash_lat(imt_ctx* ctx, int arity, const void* d_in, void* d_out, size_t n, int in_fmt, int out_fmt, cudaStream_t s) {
    // one warp per block: the few warps of such a level spread over all SMs (592 schedulers) instead of piling four to a block
    const unsigned grid = grid_for(n, kHashThreads);
    if (arity == 3) k_hash_lat<3><<<grid, kHashThreads, 0, s>>>((const uint4*)d_in, (uint4*)d_out, n, in_fmt, out_fmt, ctx->d_err);
    else k_hash_lat<2><<<grid, kHashThreads, 0, s>>>((const uint4*)d_in, (uint4*)d_out, n, in_fmt, out_fmt, ctx->d_err);
}

void launch_fold_coop(imt_ctx* ctx, const void* d_leaves, const uint64_t* d_indices, const void* d_roots, const void* d_siblings, size_t q,
                      unsigned depth, uint8_t* d_ok, void* d_roots_out, void* d_states) {
    k_fold_paths_coop<<<grid_for(4 * q, 128), 128, 0, ctx->stream>>>((const uint4*)d_leaves, d_indices, (const uint4*)d_siblings,
                                                                     (const uint4*)d_roots, q, depth, ctx->fmt, d_ok, (uint4*)d_roots_out,
                                                                     (uint4*)d_states, ctx->d_params, ctx->d_err);
}

}  // namespace imt_host

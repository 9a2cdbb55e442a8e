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

using namespace imt;

namespace imt_host {

// this unit's copy of the parameters in constant memory (the trace sinks' conversions use it)
cudaError_t latency_upload_params(const PoseidonParams* host_params) { return cudaMemcpyToSymbol(c_params, host_params, sizeof(PoseidonParams)); }

}  // namespace imt_host

namespace imt_host {

void launch_hash_coop(imt_ctx* ctx, int arity, const void* d_in, void* d_out, size_t n, int in_fmt, int out_fmt, cudaStream_t s) {
    const unsigned grid = grid_for(4 * n, 128);
    if (arity == 3)
        k_hash_coop<3><<<grid, 128, 0, s>>>((const uint4*)d_in, (uint4*)d_out, n, in_fmt, out_fmt, ctx->d_params, ctx->d_err);
    else
        k_hash_coop<2><<<grid, 128, 0, s>>>((const uint4*)d_in, (uint4*)d_out, n, in_fmt, out_fmt, ctx->d_params, ctx->d_err);
}

void launch_fold_coop(imt_ctx* ctx, const void* d_leaves, const uint64_t* d_indices, const void* d_roots, const void* d_siblings, size_t q,
                      unsigned depth, uint8_t* d_ok, void* d_roots_out, void* d_states) {
    k_fold_paths_coop<<<grid_for(4 * q, 128), 128, 0, ctx->stream>>>((const uint4*)d_leaves, d_indices, (const uint4*)d_siblings,
                                                                     (const uint4*)d_roots, q, depth, ctx->fmt, d_ok, (uint4*)d_roots_out,
                                                                     (uint4*)d_states, ctx->d_params, ctx->d_err);
}

}  // namespace imt_host

// LAB ONLY — where the cycles of the lead / helper latency kernel (csrc/poseidon_lh.cuh) go: clock64 sums per segment of the
// lead warp's and of one helper warp's partial-round loop, one block of 12 hashes. Digests are checked against k_hash_coop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Iinclude -Iindexed-merkle-tree-halo2_b200/csrc \
//        -DIMT_FREE_MASK=29 tools/lab/lh_prof.cu indexed-merkle-tree-halo2_b200/csrc/poseidon_params.cpp -o tools/_build/lh_prof
#define IMT_LH_PROF 1
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "imt_internal.h"
#include "kernels_common.cuh"
#include "poseidon_lh.cuh"
#include "poseidon_params.h"

using namespace imt;
#define CK(x)                                                             \
    do {                                                                  \
        cudaError_t e_ = (x);                                             \
        if (e_ != cudaSuccess) {                                          \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); \
            std::exit(1);                                                 \
        }                                                                 \
    } while (0)

// the lead warp's arithmetic alone (no barriers, no shared memory): the floor of a partial round
__global__ void k_lead_only(const uint4* __restrict__ in, uint4* __restrict__ out, long long* __restrict__ cyc, int rounds, int mode) {
    uint32_t x[8], y[8], k[8];
    load_fe(x, in + 2 * (threadIdx.x & 7));
    load_fe(y, in + 2 * (8 + (threadIdx.x & 3)));
    load_fe(k, in + 2 * (12 + (threadIdx.x & 3)));
    cc::clear();
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < rounds; ++r) {
        uint32_t x2[8], x4[8];
        if (mode == 0 || mode == 1) {
            mont_sqr(x2, x);
            mont_sqr(x4, x2);
        }
        if (mode == 0) mul_add(x, x4, y, k);
        if (mode == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = x4[i];
        }
        if (mode == 2) mul_add(x, x, y, k);
        if (mode == 3) mont_mul(x, x, y);
        if (mode == 4) {  // warp 0: the lead's round; the other warps: OTHER code of similar size (instruction-cache neighbours)
            if (threadIdx.x < 32) {
                mont_sqr(x2, x);
                mont_sqr(x4, x2);
                mul_add(x, x4, y, k);
            } else {
                sbox_add(x2, x, k);
                dot3(x, x2, y, k, k, y, y);
            }
        }
        if (mode == 5) {  // the same with the neighbours running the SAME code as warp 0, but started later (out of phase)
            if (threadIdx.x >= 32 && r == 0)
                for (int d = 0; d < (int)(threadIdx.x >> 5); ++d) mont_mul(y, y, k);
            mont_sqr(x2, x);
            mont_sqr(x4, x2);
            mul_add(x, x4, y, k);
        }
    }
    const long long t1 = clock64();
    store_fe(out + 2 * threadIdx.x, x);
    if ((threadIdx.x & 31) == 0) cyc[threadIdx.x >> 5] = t1 - t0;
}

int main(int argc, char** argv) {
    std::setvbuf(stdout, nullptr, _IONBF, 0);
    const size_t n = argc > 1 ? (size_t)std::atoll(argv[1]) : 12;
    static PoseidonParams hp;
    poseidon_params_generate(&hp);
    PoseidonParams* d_params;
    LhAux* d_aux;
    uint32_t* d_err;
    Fr *d_in, *d_out, *d_ref;
    CK(cudaMalloc(&d_params, sizeof(hp)));
    CK(cudaMemcpy(d_params, &hp, sizeof(hp), cudaMemcpyHostToDevice));
    CK(cudaMemcpyToSymbol(c_params, &hp, sizeof(hp)));
    CK(cudaMalloc(&d_aux, sizeof(LhAux)));
    CK(cudaMalloc(&d_err, 4));
    CK(cudaMemset(d_err, 0, 4));
    CK(cudaMalloc(&d_in, 2 * n * sizeof(Fr)));
    CK(cudaMalloc(&d_out, n * sizeof(Fr)));
    CK(cudaMalloc(&d_ref, n * sizeof(Fr)));
    std::vector<Fr> in(2 * n);
    const Fr* src = reinterpret_cast<const Fr*>(&hp.partial[0]);  // canonical Montgomery values
    for (size_t i = 0; i < 2 * n; ++i) in[i] = src[i % 300];
    CK(cudaMemcpy(d_in, in.data(), 2 * n * sizeof(Fr), cudaMemcpyHostToDevice));
    k_lh_aux<<<1, 64>>>(d_params, d_aux);
    const unsigned grid = (unsigned)((n + kLhSlots - 1) / kLhSlots);
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float best_lh = 1e9f, best_coop = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        std::printf("rep %d\n", rep);
        CK(cudaEventRecord(a));
        k_hash_lh<2><<<grid, kLhThreads>>>((const uint4*)d_in, (uint4*)d_out, n, kFmtMontgomery, kFmtMontgomery, d_params, d_aux, d_err);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        best_lh = ms < best_lh ? ms : best_lh;
        CK(cudaEventRecord(a));
        k_hash_coop<2><<<(unsigned)((4 * n + 127) / 128), 128>>>((const uint4*)d_in, (uint4*)d_ref, n, kFmtMontgomery, kFmtMontgomery, d_params, d_err);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        CK(cudaEventElapsedTime(&ms, a, b));
        best_coop = ms < best_coop ? ms : best_coop;
    }
    std::vector<Fr> o(n), r(n);
    CK(cudaMemcpy(o.data(), d_out, n * sizeof(Fr), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(r.data(), d_ref, n * sizeof(Fr), cudaMemcpyDeviceToHost));
    long long prof[32];
    CK(cudaMemcpyFromSymbol(prof, g_lh_prof, sizeof(prof)));
    std::printf("n = %zu: lead/helper %.1f us, 3 lanes per hash %.1f us, digests %s\n", n, best_lh * 1e3, best_coop * 1e3,
                std::memcmp(o.data(), r.data(), n * sizeof(Fr)) == 0 ? "equal" : "DIFFERENT");
    const double rounds = 2.0 * kRP;
    const char* lead[6] = {"wait for the helpers' full rounds (total, both permutations)", "store x + arrive", "x^2, x^4", "store x^4 + arrive",
                           "WAIT for y_0, K", "load y_0, K + x^4 y_0 + K"};
    std::printf("lead warp (cycles per partial round unless noted):\n");
    for (int i = 0; i < 6; ++i) std::printf("  %-64s %9.1f\n", lead[i], i == 0 ? (double)prof[i] : prof[i] / rounds);
    const char* help[7] = {"full rounds + hand-offs outside the loop (total)", "loop top: shuffle s_i, constants", "WAIT for x", "load x + slot A (multiply)",
                           "store y_0, K + arrive + shuffles + addend", "WAIT for x^4", "load x^4 + slot C (multiply-add)"};
    std::printf("helper warp 1:\n");
    for (int i = 0; i < 7; ++i) std::printf("  %-64s %9.1f\n", help[i], i == 0 ? (double)prof[8 + i] : prof[8 + i] / rounds);
    long long* d_cyc;
    CK(cudaMalloc(&d_cyc, 64));
    Fr* d_o2;
    CK(cudaMalloc(&d_o2, 512 * sizeof(Fr)));
    const char* names[6] = {"x^2, x^4, x^4 y + k", "x^2, x^4", "x y + k", "x y", "lead | other code", "lead | same, skewed"};
    for (int mode = 0; mode < 6; ++mode)
        for (int threads : {32, 128}) {
            k_lead_only<<<1, threads>>>((const uint4*)d_in, (uint4*)d_o2, d_cyc, 1024, mode);
            k_lead_only<<<1, threads>>>((const uint4*)d_in, (uint4*)d_o2, d_cyc, 1024, mode);
            CK(cudaDeviceSynchronize());
            long long c[4];
            CK(cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost));
            std::printf("arithmetic alone, %-22s %d warp(s) on the SM: %8.1f cycles per round\n", names[mode], threads / 32, c[0] / 1024.0);
        }
    return 0;
}

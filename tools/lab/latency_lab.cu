// Latency lab for the top-of-tree ("cooperative") kernel: what ONE warp pays per dependent field operation, and what the
// instruction forms it is built from cost when nothing else runs on the SM sub-partition. Evidence for DESIGN.md
// "Cooperative kernel"; not part of the library.
//
// Build, from the repo root, once per carry discipline (IMT_FREE_MASK = 0 .. 31, see csrc/fr.cuh; -DLAB_SWEEP keeps only the level
// timings and the known-answer check, -DLAB_PROBES adds the instruction-form probes):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Iinclude -Iindexed-merkle-tree-halo2_b200/csrc -Itools/lab \
//        -DIMT_FREE_MASK=29 tools/lab/latency_lab.cu indexed-merkle-tree-halo2_b200/csrc/poseidon_params.cpp -o tools/_build/lab_m29
// Run on a B200: prints cycles per operation (clock64 inside the kernel, one warp) and microseconds per level (CUDA events).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <vector>

#include "imt_internal.h"
#include "kernels_common.cuh"
#include "poseidon_coop.cuh"
#include "poseidon_quad.cuh"
#include "poseidon_params.h"

using namespace imt;

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));          \
            std::exit(1);                                                          \
        }                                                                          \
    } while (0)

constexpr int kChain = 2048;

// ---------------------------------------------------------------------------------- dependent chains of field operations
// op 0: x = x * y          op 1: x = x^2          op 2: x = x * y + z          op 3: x = x^5 + z          op 4: x = dot3(x, y, z; constants)
template <int OP>
__global__ void k_chain(const uint4* __restrict__ in, uint4* __restrict__ out, long long* __restrict__ cycles, int active_lanes) {
    if ((int)threadIdx.x >= active_lanes) return;
    uint32_t x[8], y[8], z[8];
    load_fe(x, in + 2 * (threadIdx.x & 31));
    load_fe(y, in + 2 * (40 + (threadIdx.x & 3)));  // lane dependent, as in the kernels: regular registers, not uniform ones
    load_fe(z, in + 2 * (44 + (threadIdx.x & 3)));
    cc::clear();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < kChain; ++i) {
        if (OP == 0) mont_mul(x, x, y);
        if (OP == 1) mont_sqr(x, x);
        if (OP == 2) mul_add(x, x, y, z);
        if (OP == 3) sbox_add(x, x, z);
        if (OP == 4) dot3(x, x, y, z, z, y, y);
        if (OP == 5) fma2(x, x, y, z, z);
    }
    const long long t1 = clock64();
    store_fe(out + 2 * threadIdx.x, x);
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ---------------------------------------------------------------------------------- one thread per hash (k_hash of the library)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_hash_tph(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
    const size_t i = blockIdx.x * (size_t)128 + threadIdx.x;
    if (i >= n) return;
    uint32_t x[2][8], d[8];
    load_fe(x[0], in + 2 * (2 * i));
    load_fe(x[1], in + 2 * (2 * i + 1));
    NoTrace nt;
    hash_fixed<2>(d, x, c_params, nt);
    store_fe(out + 2 * i, d);
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_hash3_tph(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
    const size_t i = blockIdx.x * (size_t)128 + threadIdx.x;
    if (i >= n) return;
    uint32_t x[3][8], d[8];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        load_fe(x[j], in + 2 * (3 * i + j));
        (void)ingest(x[j], kFmtMontgomery);
    }
    NoTrace nt;
    hash_fixed<3>(d, x, c_params, nt);
    store_fe(out + 2 * i, d);
}

// ---------------------------------------------------------------------------------- radix-2^29 carry-free multiplication
// 9 limbs of 29 bits (R = 2^261). Column sums are plain 64-bit accumulators: no carry flag anywhere, so every multiply-add of
// a row is independent of its neighbours. 81 + 81 + 9 multiplies instead of 64 + 64 + 8, but no carry bookkeeping.
constexpr uint32_t kM29 = (1u << 29) - 1;
__device__ __forceinline__ void mul29(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* P, uint32_t inv) {
    uint64_t c[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) c[k] = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = 0; j < 9; ++j) c[i + j] += (uint64_t)a[i] * b[j];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const uint32_t m = ((uint32_t)c[i] * inv) & kM29;
#pragma unroll
        for (int j = 0; j < 9; ++j) c[i + j] += (uint64_t)m * P[j];
        c[i + 1] += c[i] >> 29;
    }
    uint32_t t[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) t[k] = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {  // one pass, three-way split: every limb ends below 2^30 + 64
        const uint64_t v = c[9 + k];
        t[k] += (uint32_t)v & kM29;
        t[k + 1] += (uint32_t)(v >> 29) & kM29;
        t[k + 2] += (uint32_t)(v >> 58);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) r[k] = t[k];
}
__device__ __forceinline__ void sqr29(uint32_t* r, const uint32_t* a, const uint32_t* P, uint32_t inv) {
    uint64_t c[18];
    uint32_t d[9];
#pragma unroll
    for (int k = 0; k < 18; ++k) c[k] = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) d[k] = a[k] << 1;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        c[2 * i] += (uint64_t)a[i] * a[i];
#pragma unroll
        for (int j = i + 1; j < 9; ++j) c[i + j] += (uint64_t)a[i] * d[j];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const uint32_t m = ((uint32_t)c[i] * inv) & kM29;
#pragma unroll
        for (int j = 0; j < 9; ++j) c[i + j] += (uint64_t)m * P[j];
        c[i + 1] += c[i] >> 29;
    }
    uint32_t t[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) t[k] = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const uint64_t v = c[9 + k];
        t[k] += (uint32_t)v & kM29;
        t[k + 1] += (uint32_t)(v >> 29) & kM29;
        t[k + 2] += (uint32_t)(v >> 58);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) r[k] = t[k];
}
struct P29 {
    uint32_t p[9];
    uint32_t inv;
};
template <int OP>
__global__ void k_chain29(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, long long* __restrict__ cycles, P29 prm) {
    uint32_t x[9], y[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) x[k] = in[threadIdx.x * 9 + k] & kM29, y[k] = in[400 + 9 * (threadIdx.x & 3) + k] & kM29;
    uint32_t P[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) P[k] = prm.p[k] ^ (in[500 + (threadIdx.x & 1)] & 0);  // the modulus in regular registers
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < kChain; ++i) {
        if (OP == 0) mul29(x, x, y, P, prm.inv);
        if (OP == 1) sqr29(x, x, P, prm.inv);
        if (OP == 2) {  // product phase only: 81 independent-ish multiply-adds into 17 column sums, folded back to 9 limbs
            uint64_t c[18];
#pragma unroll
            for (int k = 0; k < 18; ++k) c[k] = 0;
#pragma unroll
            for (int i = 0; i < 9; ++i)
#pragma unroll
                for (int j = 0; j < 9; ++j) c[i + j] += (uint64_t)x[i] * y[j];
#pragma unroll
            for (int k = 0; k < 9; ++k) x[k] = ((uint32_t)c[k] ^ (uint32_t)(c[k + 9] >> 29)) & kM29;
        }
        if (OP == 3) {  // reduction phase only, on column sums made from x
            uint64_t c[18];
#pragma unroll
            for (int k = 0; k < 18; ++k) c[k] = (uint64_t)x[k % 9] << (k & 15);
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const uint32_t m = ((uint32_t)c[i] * prm.inv) & kM29;
#pragma unroll
                for (int j = 0; j < 9; ++j) c[i + j] += (uint64_t)m * P[j];
                c[i + 1] += c[i] >> 29;
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) x[k] = (uint32_t)c[9 + k] & kM29;
        }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int k = 0; k < 9; ++k) out[threadIdx.x * 9 + k] = x[k];
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ---------------------------------------------------------------------------------- instruction-form probes (one warp)
// MODE 0: one serial IMAD.WIDE.U32.X chain through the carry flag (what a MAC chain is)
// MODE 1: 8 independent mad.wide (no carry): the multiply pipe's issue interval for one warp
// MODE 2: serial IADD3.X chain (addc.cc)
// MODE 3: serial SHFL chain (each shuffle's source is the previous result)
// MODE 4: 8 independent SHFLs per iteration
// MODE 5: the carry chain of MODE 0 with 8 independent 32-bit adds between every 8 MACs (do ALU instructions hide behind it?)
// MODE 6: dependent L1-hit loads (pointer chase inside 1 KB of global memory)
// MODE 7: 8 independent mad.wide accumulators and nothing else: the issue interval of the multiply pipe for one warp
// MODE 8: mad.wide dependent through its accumulator (no carry flag): latency through the 64-bit addend
// MODE 9: two carry chains of 8, each started without the flag (ptxas may overlap them)
template <int MODE>
__global__ void k_probe(uint32_t* __restrict__ buf, long long* __restrict__ cycles, int active_lanes, int iters) {
    if ((int)threadIdx.x >= active_lanes) return;
    uint32_t lo[8], hi[8], s[8];
    uint64_t acc[8];
    uint32_t acc32[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc32[j] = buf[200 + j];
#pragma unroll
    for (int j = 0; j < 8; ++j) lo[j] = buf[threadIdx.x + j], hi[j] = j, s[j] = buf[j] ^ threadIdx.x, acc[j] = buf[j];
    uint32_t b = buf[33] | 1u, v = threadIdx.x, ptr = buf[64 + threadIdx.x] & 255;
    cc::clear();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 5) {
#pragma unroll
            for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
                for (int j = 0; j < 8; ++j) cc::madwc_cc(lo[j], hi[j], lo[(j + 3) & 7], b);
                if (MODE == 5) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) s[j] = s[j] * 3u + (s[(j + 1) & 7] >> 3);  // LEA / SHF / IADD on the ALU pipe
                }
            }
        }
        if (MODE == 1) {
#pragma unroll
            for (int rep = 0; rep < 4; ++rep)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(lo[j]), "r"(b));
                    lo[j] = (uint32_t)(acc[j] >> 7);
                }
        }
        if (MODE == 2) {
#pragma unroll
            for (int rep = 0; rep < 4; ++rep)
#pragma unroll
                for (int j = 0; j < 8; ++j) lo[j] = cc::addc_cc(lo[j], lo[(j + 3) & 7]);
        }
        if (MODE == 3) {
#pragma unroll
            for (int rep = 0; rep < 32; ++rep) v = __shfl_sync(0xffffffffu, v, (v + 1) & 31) + 1;
        }
        if (MODE == 4) {
#pragma unroll
            for (int rep = 0; rep < 4; ++rep)
#pragma unroll
                for (int j = 0; j < 8; ++j) s[j] = __shfl_xor_sync(0xffffffffu, s[j], 1 + j);
        }
        if (MODE == 7) {
#pragma unroll
            for (int rep = 0; rep < 4; ++rep)
#pragma unroll
                for (int j = 0; j < 8; ++j) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(lo[j]), "r"(b));
        }
        if (MODE == 8) {
#pragma unroll
            for (int rep = 0; rep < 32; ++rep) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[0]) : "r"(lo[rep & 7]), "r"(b));
        }
        if (MODE == 9) {
#pragma unroll
            for (int rep = 0; rep < 2; ++rep) {
                cc::madw_cc(lo[0], hi[0], s[0], b);
#pragma unroll
                for (int j = 1; j < 8; ++j) cc::madwc_cc(lo[j], hi[j], s[j], b);
                b += cc::addc(0, 0);
                cc::madw_cc(acc32[0], acc32[1], s[1], b);
#pragma unroll
                for (int j = 1; j < 8; ++j) cc::madwc_cc(acc32[2 * j], acc32[2 * j + 1], s[(j + 1) & 7], b);
                v += cc::addc(0, 0);
            }
        }
        if (MODE == 6) {
#pragma unroll
            for (int rep = 0; rep < 32; ++rep) ptr = __ldg(buf + 128 + ptr) & 255;
        }
    }
    const long long t1 = clock64();
    uint32_t x = v ^ ptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) x ^= lo[j] ^ hi[j] ^ s[j] ^ (uint32_t)acc[j] ^ (uint32_t)(acc[j] >> 32) ^ acc32[j] ^ acc32[j + 8];
    buf[1024 + blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static double run_cycles(void (*launch)(long long*), long long* d_cyc) {
    launch(d_cyc);  // warm (instruction cache)
    launch(d_cyc);
    CK(cudaDeviceSynchronize());
    long long h = 0;
    CK(cudaMemcpy(&h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost));
    return (double)h;
}

int main() {
    char variant[64];
    std::snprintf(variant, sizeof variant, "free-chain mask %d", (int)(IMT_FREE_MASK));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clock_khz = 0;
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
    std::printf("# latency lab (%s) on %s, %d SMs, %.0f MHz nominal\n", variant, prop.name, prop.multiProcessorCount, clock_khz / 1e3);

    static PoseidonParams hp;
    poseidon_params_generate(&hp);
    PoseidonParams* d_params;
    CK(cudaMalloc(&d_params, sizeof(hp)));
    CK(cudaMemcpy(d_params, &hp, sizeof(hp), cudaMemcpyHostToDevice));
    CK(cudaMemcpyToSymbol(c_params, &hp, sizeof(hp)));

    // inputs: canonical field elements (any constants of the parameter set will do)
    const size_t n_max = 1 << 15;
    std::vector<Fr> h_in(2 * n_max);
    const Fr* pool = &hp.partial[0].c;
    for (size_t i = 0; i < h_in.size(); ++i) h_in[i] = pool[i % (kRP * 6)];
    Fr *d_in, *d_out;
    CK(cudaMalloc(&d_in, h_in.size() * sizeof(Fr)));
    CK(cudaMalloc(&d_out, n_max * sizeof(Fr)));
    CK(cudaMemcpy(d_in, h_in.data(), h_in.size() * sizeof(Fr), cudaMemcpyHostToDevice));
    long long* d_cyc;
    CK(cudaMalloc(&d_cyc, 1024 * sizeof(long long)));
    uint32_t* d_buf;
    CK(cudaMalloc(&d_buf, 1 << 20));
    {
        std::vector<uint32_t> hb(1 << 18);
        for (size_t i = 0; i < hb.size(); ++i) hb[i] = (uint32_t)(i * 2654435761u + 12345u);
        CK(cudaMemcpy(d_buf, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
    }
    uint32_t* d_err;
    CK(cudaMalloc(&d_err, 4));
    CK(cudaMemset(d_err, 0, 4));

#ifndef LAB_SWEEP
    // ---- field-operation chains, one warp
    const char* names[6] = {"mont_mul", "mont_sqr", "mul_add", "sbox_add (x^5 + c)", "dot3", "fma2 (a b + c + d)"};
    for (int lanes : {32}) {
        std::printf("## dependent chains, one warp, %d active lanes: cycles per operation\n", lanes);
        auto run = [&](int op) {
            for (int rep = 0; rep < 2; ++rep) {
                switch (op) {
                    case 0: k_chain<0><<<1, 32>>>((const uint4*)d_in, (uint4*)d_out, d_cyc, lanes); break;
                    case 1: k_chain<1><<<1, 32>>>((const uint4*)d_in, (uint4*)d_out, d_cyc, lanes); break;
                    case 2: k_chain<2><<<1, 32>>>((const uint4*)d_in, (uint4*)d_out, d_cyc, lanes); break;
                    case 3: k_chain<3><<<1, 32>>>((const uint4*)d_in, (uint4*)d_out, d_cyc, lanes); break;
                    case 4: k_chain<4><<<1, 32>>>((const uint4*)d_in, (uint4*)d_out, d_cyc, lanes); break;
                    case 5: k_chain<5><<<1, 32>>>((const uint4*)d_in, (uint4*)d_out, d_cyc, lanes); break;
                }
            }
            CK(cudaDeviceSynchronize());
            long long h = 0;
            CK(cudaMemcpy(&h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost));
            return (double)h / kChain;
        };
        for (int op = 0; op < 6; ++op) std::printf("%-22s %9.1f\n", names[op], run(op));
    }
    // two warps of the same block on the same sub-partition (warps 0 and 4): per-warp cycles per mont_mul
    {
        k_chain<0><<<1, 160>>>((const uint4*)d_in, (uint4*)d_out, d_cyc, 160);
        k_chain<0><<<1, 160>>>((const uint4*)d_in, (uint4*)d_out, d_cyc, 160);
        CK(cudaDeviceSynchronize());
        long long h = 0;
        CK(cudaMemcpy(&h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost));
        std::printf("mont_mul, 5 warps in the block (2 on sub-partition 0): %9.1f cycles per operation per warp\n", (double)h / kChain);
    }
#endif  // LAB_SWEEP
#if IMT_FREE_MASK == 0 && defined(LAB_PROBES)
    {
        std::printf("## radix-2^29 carry-free arithmetic (9 limbs, 64-bit column sums), one warp: cycles per operation\n");
        P29 prm;
        for (int k = 0; k < 9; ++k) prm.p[k] = (0x10000001u * (k + 3)) & kM29;
        prm.p[0] = 0x10000001u;
        prm.inv = 0x0fffffffu;
        for (int op = 0; op < 4; ++op) {
            for (int rep = 0; rep < 2; ++rep) {
                if (op == 0) k_chain29<0><<<1, 32>>>(d_buf, d_buf + 4096, d_cyc, prm);
                if (op == 1) k_chain29<1><<<1, 32>>>(d_buf, d_buf + 4096, d_cyc, prm);
                if (op == 2) k_chain29<2><<<1, 32>>>(d_buf, d_buf + 4096, d_cyc, prm);
                if (op == 3) k_chain29<3><<<1, 32>>>(d_buf, d_buf + 4096, d_cyc, prm);
            }
            CK(cudaDeviceSynchronize());
            long long h = 0;
            CK(cudaMemcpy(&h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost));
            const char* nm29[4] = {"mul29", "sqr29", "mul29 product only", "mul29 reduction only"};
            std::printf("%-22s %9.1f\n", nm29[op], (double)h / kChain);
        }
    }
    {
        std::printf("## instruction forms, one warp: cycles per instruction\n");
        const int iters = 512;
        struct Row {
            const char* name;
            int mode;
            int per_iter;
        } rows[] = {{"IMAD.WIDE.U32.X serial carry chain", 0, 32}, {"mad.wide x8 independent", 1, 32}, {"IADD3.X serial carry chain", 2, 32},
                    {"SHFL serial", 3, 32},                        {"SHFL x8 independent", 4, 32},      {"carry chain + 8 ALU per 8 MAC (per MAC)", 5, 32},
                    {"LDG L1-hit pointer chase", 6, 32},           {"mad.wide x8 independent, nothing else", 7, 32}, {"mad.wide serial through the addend", 8, 32},
                    {"two free carry chains of 8 (per MAC)", 9, 32}};
        for (int lanes : {32}) {
            for (const Row& r : rows) {
                for (int rep = 0; rep < 2; ++rep) {
                    switch (r.mode) {
                        case 0: k_probe<0><<<1, 32>>>(d_buf, d_cyc, lanes, iters); break;
                        case 1: k_probe<1><<<1, 32>>>(d_buf, d_cyc, lanes, iters); break;
                        case 2: k_probe<2><<<1, 32>>>(d_buf, d_cyc, lanes, iters); break;
                        case 3: k_probe<3><<<1, 32>>>(d_buf, d_cyc, lanes, iters); break;
                        case 4: k_probe<4><<<1, 32>>>(d_buf, d_cyc, lanes, iters); break;
                        case 5: k_probe<5><<<1, 32>>>(d_buf, d_cyc, lanes, iters); break;
                        case 6: k_probe<6><<<1, 32>>>(d_buf, d_cyc, lanes, iters); break;
                        case 7: k_probe<7><<<1, 32>>>(d_buf, d_cyc, lanes, iters); break;
                        case 8: k_probe<8><<<1, 32>>>(d_buf, d_cyc, lanes, iters); break;
                        case 9: k_probe<9><<<1, 32>>>(d_buf, d_cyc, lanes, iters); break;
                    }
                }
                CK(cudaDeviceSynchronize());
                long long h = 0;
                CK(cudaMemcpy(&h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost));
                std::printf("%-44s lanes=%2d %8.2f\n", r.name, lanes, (double)h / ((double)iters * r.per_iter));
            }
        }
    }
#endif
    // ---- the cooperative level kernel: microseconds per level for n nodes (CUDA events, best of 5)
    std::printf("## k_hash_coop<2>: microseconds per level\n");
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (size_t n : {(size_t)8, (size_t)64, (size_t)512, (size_t)1024, (size_t)2048, (size_t)4096, (size_t)8192, (size_t)16384}) {
        float best = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            CK(cudaEventRecord(e0));
            k_hash_coop<2><<<(unsigned)((4 * n + 127) / 128), 128>>>((const uint4*)d_in, (uint4*)d_out, n, kFmtMontgomery, kFmtMontgomery, d_params, d_err);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
        }
        std::printf("n = %6zu  %9.1f us\n", n, best * 1e3);
    }
    // ---- one thread per hash at the sizes of the middle levels (occupancy hint 7 = the library's throughput kernel, 1 = none)
    std::printf("## thread-per-hash kernel: microseconds per level  (launch bounds 128 x 7 | 128 x 1)\n");
    {
        const size_t big = (size_t)1 << 22;
        Fr *b_in, *b_out;
        CK(cudaMalloc(&b_in, 2 * big * sizeof(Fr)));
        CK(cudaMalloc(&b_out, big * sizeof(Fr)));
        for (size_t off = 0; off < 2 * big; off += h_in.size())
            CK(cudaMemcpy(b_in + off, h_in.data(), std::min(h_in.size(), 2 * big - off) * sizeof(Fr), cudaMemcpyHostToDevice));
        for (size_t n : {(size_t)4096, (size_t)8192, (size_t)16384, (size_t)32768, (size_t)65536, (size_t)131072, (size_t)262144, (size_t)1 << 22}) {
            float best7 = 1e9f, best1 = 1e9f;
            for (int rep = 0; rep < 5; ++rep) {
                float ms;
                CK(cudaEventRecord(e0));
                k_hash_tph<7><<<(unsigned)((n + 127) / 128), 128>>>((const uint4*)b_in, (uint4*)b_out, n);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep && ms < best7) best7 = ms;
                CK(cudaEventRecord(e0));
                k_hash_tph<1><<<(unsigned)((n + 127) / 128), 128>>>((const uint4*)b_in, (uint4*)b_out, n);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep && ms < best1) best1 = ms;
            }
            std::printf("n = %6zu  %9.1f us | %9.1f us\n", n, best7 * 1e3, best1 * 1e3);
        }
    }
#ifdef LAB_LEAF
    {   // the leaf kernel of the build: H3 of 2^22 preimages, one thread per hash, launch bounds 128 x 7 (the library's) and 128 x 6
        const size_t n = (size_t)1 << 22;
        Fr *b_in, *b_out;
        CK(cudaMalloc(&b_in, 3 * n * sizeof(Fr)));
        CK(cudaMalloc(&b_out, n * sizeof(Fr)));
        for (size_t off = 0; off < 3 * n; off += h_in.size())
            CK(cudaMemcpy(b_in + off, h_in.data(), std::min(h_in.size(), 3 * n - off) * sizeof(Fr), cudaMemcpyHostToDevice));
        float best7 = 1e9f, best6 = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            float ms;
            CK(cudaEventRecord(e0));
            k_hash3_tph<7><<<(unsigned)((n + 127) / 128), 128>>>((const uint4*)b_in, (uint4*)b_out, n);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best7) best7 = ms;
            CK(cudaEventRecord(e0));
            k_hash3_tph<6><<<(unsigned)((n + 127) / 128), 128>>>((const uint4*)b_in, (uint4*)b_out, n);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best6) best6 = ms;
        }
        std::printf("## leaf kernel H3, n = 2^22: %9.1f us (128 x 7) | %9.1f us (128 x 6)\n", best7 * 1e3, best6 * 1e3);
        float n6 = 1e9f, n5 = 1e9f, l5 = 1e9f, n8 = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            float ms;
            CK(cudaEventRecord(e0));
            k_hash_tph<6><<<(unsigned)((n + 127) / 128), 128>>>((const uint4*)b_in, (uint4*)b_out, n);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < n6) n6 = ms;
            CK(cudaEventRecord(e0));
            k_hash_tph<5><<<(unsigned)((n + 127) / 128), 128>>>((const uint4*)b_in, (uint4*)b_out, n);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < n5) n5 = ms;
            CK(cudaEventRecord(e0));
            k_hash_tph<8><<<(unsigned)((n + 127) / 128), 128>>>((const uint4*)b_in, (uint4*)b_out, n);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < n8) n8 = ms;
            CK(cudaEventRecord(e0));
            k_hash3_tph<5><<<(unsigned)((n + 127) / 128), 128>>>((const uint4*)b_in, (uint4*)b_out, n);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < l5) l5 = ms;
        }
        std::printf("## node kernel H2, n = 2^22: %9.1f us (128 x 8) | %9.1f us (128 x 6) | %9.1f us (128 x 5);  leaf 128 x 5: %9.1f us\n", n8 * 1e3, n6 * 1e3, n5 * 1e3, l5 * 1e3);
        CK(cudaFree(b_in));
        CK(cudaFree(b_out));
    }
#endif
    // ---- second generation: 4 lanes per hash, 3 slots per partial round (poseidon_quad.cuh); must give the same digests
    QuadAux* d_aux;
    CK(cudaMalloc(&d_aux, sizeof(QuadAux)));
    k_quad_aux<<<1, 64>>>(d_params, d_aux);
    Fr* d_out2;
    CK(cudaMalloc(&d_out2, n_max * sizeof(Fr)));
    std::printf("## k_hash_quad<2>: microseconds per level\n");
    for (size_t n : {(size_t)8, (size_t)64, (size_t)512, (size_t)1024, (size_t)2048, (size_t)4096, (size_t)8192, (size_t)16384}) {
        float best = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            CK(cudaEventRecord(e0));
            k_hash_quad<2><<<(unsigned)((4 * n + 127) / 128), 128>>>((const uint4*)d_in, (uint4*)d_out2, n, kFmtMontgomery, kFmtMontgomery, d_params, d_aux, d_err);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
        }
        std::printf("n = %6zu  %9.1f us\n", n, best * 1e3);
    }
    {
        const size_t n = 16384;
        k_hash_coop<2><<<(unsigned)((4 * n + 127) / 128), 128>>>((const uint4*)d_in, (uint4*)d_out, n, kFmtMontgomery, kFmtMontgomery, d_params, d_err);
        k_hash_quad<2><<<(unsigned)((4 * n + 127) / 128), 128>>>((const uint4*)d_in, (uint4*)d_out2, n, kFmtMontgomery, kFmtMontgomery, d_params, d_aux, d_err);
        std::vector<Fr> a(n), b(n);
        CK(cudaMemcpy(a.data(), d_out, n * sizeof(Fr), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), d_out2, n * sizeof(Fr), cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t i = 0; i < n; ++i) bad += std::memcmp(&a[i], &b[i], sizeof(Fr)) != 0;
        std::printf("k_hash_quad<2> vs k_hash_coop<2> on %zu pairs: %zu differ%s\n", n, bad, bad ? "  <-- MISMATCH" : " (bit-exact)");
        k_hash_coop<3><<<(unsigned)((4 * n + 127) / 128), 128>>>((const uint4*)d_in, (uint4*)d_out, n / 2, kFmtCanonical, kFmtCanonical, d_params, d_err);
        k_hash_quad<3><<<(unsigned)((4 * n + 127) / 128), 128>>>((const uint4*)d_in, (uint4*)d_out2, n / 2, kFmtCanonical, kFmtCanonical, d_params, d_aux, d_err);
        CK(cudaMemcpy(a.data(), d_out, n / 2 * sizeof(Fr), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), d_out2, n / 2 * sizeof(Fr), cudaMemcpyDeviceToHost));
        bad = 0;
        for (size_t i = 0; i < n / 2; ++i) bad += std::memcmp(&a[i], &b[i], sizeof(Fr)) != 0;
        std::printf("k_hash_quad<3> vs k_hash_coop<3> on %zu triples (canonical format): %zu differ%s\n", n / 2, bad, bad ? "  <-- MISMATCH" : " (bit-exact)");
    }
    {   // known answer: H2(0, 0) = 0x2b2ceb...: both latency kernels on canonical zeros (SURVEY.md 8c derived vector)
        CK(cudaMemset(d_in, 0, 64 * sizeof(Fr)));
        k_hash_coop<2><<<1, 128>>>((const uint4*)d_in, (uint4*)d_out, 4, kFmtCanonical, kFmtCanonical, d_params, d_err);
        k_hash_quad<2><<<1, 128>>>((const uint4*)d_in, (uint4*)d_out2, 4, kFmtCanonical, kFmtCanonical, d_params, d_aux, d_err);
        uint32_t a[8], b[8];
        CK(cudaMemcpy(a, d_out, 32, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b, d_out2, 32, cudaMemcpyDeviceToHost));
        std::printf("H2(0,0) coop %08x%08x...%08x quad %08x%08x...%08x   (want 2b2ceb8eb042a119...80eeb04f)\n", a[7], a[6], a[0], b[7], b[6], b[0]);
    }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}

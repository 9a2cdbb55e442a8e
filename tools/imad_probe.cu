// Integer-pipe microbenchmarks for sm_100a: how many 32x32(+64) multiply-accumulates per clock per SM, in the forms
// the field arithmetic uses. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/imad_probe tools/imad_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define ITERS 4096

// 8 independent mad.wide.u32 chains per thread
__global__ void __launch_bounds__(256) k_wide_indep(uint64_t* out, uint32_t seed) {
    uint64_t acc[8]; uint32_t a[8]; uint32_t b = seed + threadIdx.x;
    for (int j = 0; j < 8; ++j) { acc[j] = seed + j; a[j] = seed * 31 + j + blockIdx.x; }
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // the multiplier follows the accumulator so the product is not loop invariant
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(a[j]), "r"(b));
                a[j] = (uint32_t)acc[j];
            }
    uint64_t x = 0; for (int j = 0; j < 8; ++j) x ^= acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
// 8 independent 32-bit mad.lo chains
__global__ void __launch_bounds__(256) k_lo_indep(uint64_t* out, uint32_t seed) {
    uint32_t acc[8]; uint32_t a[8]; uint32_t b = seed + threadIdx.x;
    for (int j = 0; j < 8; ++j) { acc[j] = seed + j; a[j] = seed * 31 + j + blockIdx.x; }
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("mad.lo.u32 %0, %0, %2, %1;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
    uint64_t x = 0; for (int j = 0; j < 8; ++j) x ^= acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
// ONE serial carry chain per thread: 32 wide MACs chained through the carry flag (what fr.cuh emits)
__global__ void __launch_bounds__(256) k_wide_carry_serial(uint64_t* out, uint32_t seed) {
    uint32_t lo[8], hi[8], a[8]; uint32_t b = seed + threadIdx.x;
    for (int j = 0; j < 8; ++j) { lo[j] = seed + j; hi[j] = j; a[j] = seed * 31 + j + blockIdx.x; }
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(lo[(j + 3) & 7]), "r"(b));
        }
    }
    uint64_t x = 0; for (int j = 0; j < 8; ++j) x ^= ((uint64_t)hi[j] << 32) | lo[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
// TWO independent carry chains is not expressible in PTX (one flag); instead: chains of 4 that restart (mad.lo.cc starts)
__global__ void __launch_bounds__(256) k_wide_carry_chains4(uint64_t* out, uint32_t seed) {
    uint32_t lo[8], hi[8], a[8]; uint32_t b = seed + threadIdx.x;
    for (int j = 0; j < 8; ++j) { lo[j] = seed + j; hi[j] = j; a[j] = seed * 31 + j + blockIdx.x; }
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if ((j & 3) == 0)
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(lo[(j + 3) & 7]), "r"(b));
                else
                    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(lo[(j + 3) & 7]), "r"(b));
            }
        }
    }
    uint64_t x = 0; for (int j = 0; j < 8; ++j) x ^= ((uint64_t)hi[j] << 32) | lo[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
// wide MACs interleaved 1:1 with independent 3-input adds (does the ALU pipe co-issue for free?)
__global__ void __launch_bounds__(256) k_wide_plus_alu(uint64_t* out, uint32_t seed) {
    uint64_t acc[8]; uint32_t a[8], s[8]; uint32_t b = seed + threadIdx.x;
    for (int j = 0; j < 8; ++j) { acc[j] = seed + j; a[j] = seed * 31 + j + blockIdx.x; s[j] = j * seed; }
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(a[j]), "r"(b));
                a[j] = (uint32_t)acc[j];
                asm volatile("add.u32 %0, %0, %1;" : "+r"(s[j]) : "r"(s[(j + 1) & 7]));
            }
    uint64_t x = 0; for (int j = 0; j < 8; ++j) x ^= acc[j] + s[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

// 8 independent double-precision FMA chains (is the FP64 pipe a usable second multiplier on this part?)
__global__ void __launch_bounds__(256) k_dfma_indep(uint64_t* out, uint32_t seed) {
    double acc[8]; double a = 1.0 + 1e-9 * (seed + threadIdx.x), b = 1e-12 * blockIdx.x;
    for (int j = 0; j < 8; ++j) acc[j] = 1.0 + j * 1e-3;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(acc[j]) : "d"(a), "d"(b));
    double x = 0; for (int j = 0; j < 8; ++j) x += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint64_t)__double_as_longlong(x);
}
// the serial wide+carry chain of the field arithmetic with one independent DFMA per wide MAC: do the two pipes overlap?
__global__ void __launch_bounds__(256) k_wide_carry_plus_dfma(uint64_t* out, uint32_t seed) {
    uint32_t lo[8], hi[8]; uint32_t b = seed + threadIdx.x;
    double acc[8]; double a = 1.0 + 1e-9 * (seed + threadIdx.x), c = 1e-12 * blockIdx.x;
    for (int j = 0; j < 8; ++j) { lo[j] = seed + j; hi[j] = j; acc[j] = 1.0 + j * 1e-3; }
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[j]), "+r"(hi[j]) : "r"(lo[(j + 3) & 7]), "r"(b));
                asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(acc[j]) : "d"(a), "d"(c));
            }
        }
    }
    double y = 0; for (int j = 0; j < 8; ++j) y += acc[j];
    uint64_t x = (uint64_t)__double_as_longlong(y); for (int j = 0; j < 8; ++j) x ^= ((uint64_t)hi[j] << 32) | lo[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <class K>
static void run(const char* name, K kernel, int blocks_per_sm, int sms, double clock_hz, double ops_per_thread_iter) {
    int blocks = blocks_per_sm * sms, threads = 256;
    uint64_t* out; cudaMalloc(&out, (size_t)blocks * threads * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kernel<<<blocks, threads>>>(out, 7); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); kernel<<<blocks, threads>>>(out, 11 + rep); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double ops = (double)blocks * threads * ITERS * ops_per_thread_iter;
    double rate = ops / (best * 1e-3);
    printf("%-28s blocks/SM=%d  %8.3f ms  %10.3e MAC/s  %6.2f MAC/clk/SM @%.0f MHz (nominal max clock)\n", name, blocks_per_sm, best, rate,
           rate / sms / clock_hz, clock_hz / 1e6);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double hz = clk_khz * 1e3;
    printf("%s: %d SMs, clockRate %.0f MHz\n", p.name, p.multiProcessorCount, hz / 1e6);
    for (int bps : {1, 2, 4, 8}) {
        run("mad.wide.u32 x8 independent", k_wide_indep, bps, p.multiProcessorCount, hz, 32);
        run("mad.lo.u32 x8 independent", k_lo_indep, bps, p.multiProcessorCount, hz, 32);
        run("wide+carry, one serial chain", k_wide_carry_serial, bps, p.multiProcessorCount, hz, 32);
        run("wide+carry, chains of 4", k_wide_carry_chains4, bps, p.multiProcessorCount, hz, 32);
        run("mad.wide + add 1:1", k_wide_plus_alu, bps, p.multiProcessorCount, hz, 32);
        run("fma.rz.f64 x8 independent", k_dfma_indep, bps, p.multiProcessorCount, hz, 32);
        run("wide+carry chain + dfma 1:1", k_wide_carry_plus_dfma, bps, p.multiProcessorCount, hz, 32);
    }
    return 0;
}

// LAB ONLY — what ONE warp can issue on the integer-multiply pipe of a B200 SM sub-partition, by instruction form. Evidence for
// DESIGN.md "Carry discipline" / profiles/r02_latency_lab.md section 5: is the lone-warp hash latency bounded by the dependent
// carry-predicate latency (then interleaving chains would help) or by the per-warp issue interval of IMAD.WIDE itself (then the
// latency kernel already sits at the hardware floor)?  ptxas spaces ANY two IMAD.WIDE of one warp 4 cycles apart in its static
// schedule (cuobjdump control words: stall = 4), dependent or not.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/lab/issue_probe.cu -o tools/_build/issue_probe
//
// Every probe keeps its operands loop-carried (nothing can be hoisted or strength-reduced) and was checked in the SASS.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                             \
    do {                                                                  \
        cudaError_t e_ = (x);                                             \
        if (e_ != cudaSuccess) {                                          \
            std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); \
            std::exit(1);                                                 \
        }                                                                 \
    } while (0)

__device__ __forceinline__ void madw_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
__device__ __forceinline__ void madwc_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
__device__ __forceinline__ uint32_t addc0(uint32_t x) {
    uint32_t d;
    asm volatile("addc.u32 %0, %1, 0;" : "=r"(d) : "r"(x));
    return d;
}
__device__ __forceinline__ void madw(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {  // 64-bit accumulate, no carry out
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}

constexpr int kUnroll = 4;   // bodies per loop iteration
constexpr int kMacs = 8;     // wide multiply-accumulates per body

// MODE 0: 8 INDEPENDENT chain heads (carry out, no carry in), each carry consumed by its own IADD3.X
// MODE 1: 8 independent 64-bit accumulates without carry out
// MODE 2: one serial chain of 8 through the carry (a product row: the reference point, 7.05 cycles per instruction)
// MODE 3: the serial chain of MODE 2 with one independent carry-less accumulate between every two links
// MODE 4: the serial chain of MODE 2 with one independent ALU add (IADD3) between every two links
// MODE 5: two serial chains of 4 (the even / odd strands of one product row), the second one free to overlap the first
// MODE 6: 8 independent 32-bit IMAD + 8 IMAD.HI (the product split into halves, no accumulation into pairs)
template <int MODE>
__global__ void k_issue(uint32_t* __restrict__ buf, long long* __restrict__ cycles, int iters) {
    uint32_t lo[kMacs], hi[kMacs], a[kMacs], k[kMacs], xlo[kMacs], xhi[kMacs];
#pragma unroll
    for (int j = 0; j < kMacs; ++j) {
        lo[j] = buf[threadIdx.x + j], hi[j] = buf[64 + j], a[j] = buf[128 + j] ^ threadIdx.x, k[j] = 0;
        xlo[j] = buf[192 + j], xhi[j] = buf[256 + j];
    }
    uint32_t b = buf[33] | 1u;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < kUnroll; ++rep) {
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < kMacs; ++j) {
                    madw_cc(lo[j], hi[j], a[j], b);
                    k[j] = addc0(k[j]);
                }
            }
            if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < kMacs; ++j) madw(lo[j], hi[j], a[j], b);
            }
            if (MODE == 2 || MODE == 3 || MODE == 4) {
                madw_cc(lo[0], hi[0], a[0], b);
#pragma unroll
                for (int j = 1; j < kMacs; ++j) {
                    if (MODE == 3) madw(xlo[j], xhi[j], a[j], b);
                    if (MODE == 4) xlo[j] += xhi[j];
                    madwc_cc(lo[j], hi[j], a[j], b);
                }
                k[0] = addc0(k[0]);
            }
            if (MODE == 5) {
                madw_cc(lo[0], hi[0], a[0], b);
                madwc_cc(lo[2], hi[2], a[2], b);
                madwc_cc(lo[4], hi[4], a[4], b);
                madwc_cc(lo[6], hi[6], a[6], b);
                k[0] = addc0(k[0]);
                madw_cc(lo[1], hi[1], a[1], b);
                madwc_cc(lo[3], hi[3], a[3], b);
                madwc_cc(lo[5], hi[5], a[5], b);
                madwc_cc(lo[7], hi[7], a[7], b);
                k[1] = addc0(k[1]);
            }
            if (MODE == 6) {
#pragma unroll
                for (int j = 0; j < kMacs; ++j) {
                    lo[j] = a[j] * b + lo[j];
                    hi[j] = __umulhi(a[j], b) + hi[j];
                }
            }
            b += 2;  // loop-carried multiplier: nothing is invariant
        }
    }
    const long long t1 = clock64();
    uint32_t x = b;
#pragma unroll
    for (int j = 0; j < kMacs; ++j) x ^= lo[j] ^ hi[j] ^ k[j] ^ xlo[j] ^ xhi[j];
    buf[1024 + blockIdx.x * blockDim.x + threadIdx.x] = x;
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int MODE>
static void run(const char* name, uint32_t* d_buf, long long* d_cyc, int macs_per_body, int extra_per_body) {
    const int iters = 2048;
    // (threads per block) : 32 = one warp alone on the SM; 128 = one warp on each of the four sub-partitions; 256 = two per sub-partition
    for (int threads : {32, 128, 256, 512}) {
        k_issue<MODE><<<1, threads>>>(d_buf, d_cyc, iters);  // warm-up (instruction cache)
        k_issue<MODE><<<1, threads>>>(d_buf, d_cyc, iters);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(threads / 32);
        CK(cudaMemcpy(h.data(), d_cyc, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        long long worst = 0;
        for (long long v : h) worst = v > worst ? v : worst;
        const double per_body = (double)worst / ((double)iters * kUnroll);
        std::printf("%-58s warps/SM %2d  cycles per body %7.2f  per wide MAC %5.2f%s\n", name, threads / 32, per_body, per_body / macs_per_body,
                    extra_per_body ? "  (+ the interleaved independent instructions)" : "");
    }
}

int main() {
    uint32_t* d_buf;
    long long* d_cyc;
    CK(cudaMalloc(&d_buf, 1 << 20));
    CK(cudaMalloc(&d_cyc, 4096));
    std::vector<uint32_t> h(1 << 18);
    uint64_t s = 0x494D54;
    for (auto& v : h) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        v = (uint32_t)(s >> 32);
    }
    CK(cudaMemcpy(d_buf, h.data(), 1 << 20, cudaMemcpyHostToDevice));
    run<0>("0: 8 independent heads (carry out) + 8 IADD3.X", d_buf, d_cyc, 8, 0);
    run<1>("1: 8 independent 64-bit accumulates, no carry out", d_buf, d_cyc, 8, 0);
    run<2>("2: one serial carry chain of 8", d_buf, d_cyc, 8, 0);
    run<3>("3: serial chain of 8 + 7 independent carry-less MACs between", d_buf, d_cyc, 8, 7);
    run<4>("4: serial chain of 8 + 7 independent IADD between", d_buf, d_cyc, 8, 7);
    run<5>("5: two chains of 4, second free to overlap", d_buf, d_cyc, 8, 0);
    run<6>("6: 8 x (IMAD lo + IMAD.HI), 32-bit accumulators", d_buf, d_cyc, 8, 0);
    return 0;
}

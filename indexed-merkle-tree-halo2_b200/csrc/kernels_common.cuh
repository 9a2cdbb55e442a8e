// Device helpers shared by the kernels of every translation unit: the Poseidon parameters in constant memory and
// 128-bit vectorised field-element loads / stores with format conversion at the edges.
#pragma once
#include <cuda_runtime.h>

#include "imt_internal.h"
#include "poseidon.cuh"

namespace imt {

// private per translation unit (no relocatable device code): each TU that hashes uploads its own copy
static __constant__ PoseidonParams c_params;

__device__ __forceinline__ void load_fe(uint32_t* x, const uint4* p) {
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    x[0] = a.x, x[1] = a.y, x[2] = a.z, x[3] = a.w;
    x[4] = b.x, x[5] = b.y, x[6] = b.z, x[7] = b.w;
}
__device__ __forceinline__ void store_fe(uint4* p, const uint32_t* x) {
    p[0] = make_uint4(x[0], x[1], x[2], x[3]);
    p[1] = make_uint4(x[4], x[5], x[6], x[7]);
}
// user format -> Montgomery (semi-reduced). Returns false when the input is not < p.
__device__ __forceinline__ bool ingest(uint32_t* x, int fmt) {
    const bool ok = is_canonical(x);
    if (fmt == kFmtCanonical) to_mont(x, x);
    return ok;
}
// Montgomery canonical -> user format
__device__ __forceinline__ void egress(uint32_t* x, int fmt) {
    if (fmt == kFmtCanonical) from_mont(x, x);
}

}  // namespace imt

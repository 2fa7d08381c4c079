// Box.Sample() of an arbitrary box on device: the reference's four-way split per component
// (src/Gym/Spaces/Box.cs:69-90):
//   low and high finite   uniform(low, high)                       :84
//   only low finite       low + exponential(1)                     :82
//   only high finite      high + exponential(1)                    :83  (the reference ADDS to High; upstream gym subtracts)
//   neither               normal(0.5, 1)                           :81  (mean 0.5: the reference's constant)
// and floor() for integer dtypes (:85-88).  NumSharp's generator is replaced by the engine's Philox stream (RNG
// spec): sample c, component j draws block (seed, env id = j, index + c, STREAM_SPACE); w0 feeds the uniform / exponential
// (u = (w0 >> 8) * 2^-24 in [0, 1), exponential = -log1p(-u)), w1 and w2 the normal (Box-Muller with u1 = ((w1 >> 8) + 1) *
// 2^-24 in (0, 1], u2 = (w2 >> 8) * 2^-24: z = sqrt(-2 ln u1) * cos(2 pi u2)).
// One thread per (sample, component); the output row-major [count][dim] is written coalesced.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "philox.cuh"

namespace gymcuda {

constexpr uint32_t STREAM_SPACE = 4;

struct BoxSampleArgs {
    const float* low;    // [dim]
    const float* high;   // [dim]
    float* out;          // [count][dim]
    int dim;
    long long total;     // count * dim
    int as_int;
    uint64_t seed, index;
};

__device__ __forceinline__ bool box_finite(float x) { return fabsf(x) <= 3.4028234663852886e38f; }   // false for +-inf and NaN

__global__ void __launch_bounds__(256) box_sample_kernel(const BoxSampleArgs p) {
    const long long g = (long long)blockIdx.x * 256 + threadIdx.x;
    if (g >= p.total) return;
    const int j = (int)(g % p.dim);
    const uint64_t c = (uint64_t)(g / p.dim);
    const float lo = p.low[j], hi = p.high[j];
    const bool bl = box_finite(lo) && lo == lo, bh = box_finite(hi) && hi == hi;
    const Block b = draw(p.seed, (uint32_t)j, p.index + c, STREAM_SPACE);
    float v;
    if (bl && bh) {
        v = uniformf(lo, hi, b.w0);
    } else if (bl || bh) {
        const float e = -log1pf(-u01(b.w0));
        v = __fadd_rn(bl ? lo : hi, e);
    } else {
        const float u1 = __fmul_rn((float)((b.w1 >> 8) + 1u), 0x1p-24f);
        const float u2 = u01(b.w2);
        const float z = __fmul_rn(sqrtf(__fmul_rn(-2.0f, logf(u1))), cospif(__fmul_rn(2.0f, u2)));
        v = __fadd_rn(0.5f, z);
    }
    if (p.as_int) v = floorf(v);
    p.out[g] = v;
}

}  // namespace gymcuda

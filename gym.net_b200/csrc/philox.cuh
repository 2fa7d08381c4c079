// Counter-based RNG of the engine ("gymcuda RNG spec v1").
//
// Replaces the reference's NumSharp NumPyRandom draws -- CartPoleEnv.cs:65 (reset uniform),
// LunarLanderEnv.cs:496,507,611-612 (reset + per-step dispersion), Discrete.cs:27 / Box.cs:84
// (random policy) -- with Philox4x32-10 keyed by (seed, global env id): stateless, so a thread
// regenerates any draw from (t, env) alone and nothing RNG-related is stored in HBM.
//
//   key = (seed_lo, global_env_id)
//   ctr = (index_lo, index_hi, stream | sub << 8, seed_hi)
// index: RESET = episode ordinal of the env (how many resets it has had), ACTION = block number
// of the random-policy draw at global step t, DYNAMICS = global step t.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gymcuda {

enum : uint32_t { STREAM_RESET = 0, STREAM_ACTION = 1, STREAM_DYNAMICS = 2, STREAM_CTOR = 3 };

struct Block { uint32_t w0, w1, w2, w3; };

__device__ __forceinline__ Block philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    return Block{c0, c1, c2, c3};
}

__device__ __forceinline__ Block draw(uint64_t seed, uint32_t env_id, uint64_t index, uint32_t stream,
                                      uint32_t sub = 0) {
    return philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), stream | (sub << 8),
                         (uint32_t)(seed >> 32), (uint32_t)seed, env_id);
}

__device__ __forceinline__ uint32_t word(const Block& b, uint32_t i) {
    return i == 0 ? b.w0 : (i == 1 ? b.w1 : (i == 2 ? b.w2 : b.w3));
}

// [0,1) on 24 bits: exact in fp32
__device__ __forceinline__ float u01(uint32_t w) { return __fmul_rn((float)(w >> 8), 0x1p-24f); }

// low + (high - low) * u01(w), the multiply and the add never fused.  Evaluated as float(w >> 8) * ((high - low) * 2^-24):
// the integer converts exactly, scaling the float32 span by 2^-24 is exact, so the single multiply rounds the same
// real number as (high - low) * u01(w) would -- the same bits with one multiply (the scaled span folds at compile
// time for constant bounds).
__device__ __forceinline__ float uniformf(float lo, float hi, uint32_t w) {
    return __fadd_rn(lo, __fmul_rn((float)(w >> 8), __fmul_rn(__fsub_rn(hi, lo), 0x1p-24f)));
}

}  // namespace gymcuda

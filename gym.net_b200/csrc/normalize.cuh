// Observation / reward normalisation of a batch, on device and in place.
//
// The reference has no such wrapper: its callers keep running statistics by hand around Env.Step
// (examples/ReinforcementLearning/.../PlaySessions/BasePlaySession.cs:58-69).  What every learner then adds is the
// VecNormalize recipe, restated here for a whole batch per call:
//   observations: running mean / variance of every component over all envs and calls; obs <- clip((obs - mean) / sqrt(var + eps))
//   rewards:      ret <- ret * gamma + reward per env; running variance of ret; reward <- clip(reward / sqrt(var_ret + eps));
//                 ret <- 0 where the episode ended
// Two kernels per call: the statistics must include the batch before it is normalised (update), then every env reads the
// same totals (apply).  Sums are kept in double: sum, sum of squares and count give mean and variance without a merge step.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gymcuda {

constexpr int NORM_MAX_OD = 8;
constexpr int NORM_BLOCK = 256;
constexpr int NORM_VALUES = 2 * NORM_MAX_OD + 2;   // per-env contributions: obs_j, obs_j^2, ret, ret^2
constexpr int NORM_ACC = NORM_VALUES + 2;          // + the numbers of envs accumulated so far: [18] into the observation sums, [19] into the return sums

// acc (double): [0..7] sum obs_j, [8..15] sum obs_j^2, [16] sum ret, [17] sum ret^2, [18] count of observations, [19] count of returns
// (two counts: a call may carry only observations -- the batch Reset returns -- or only rewards, and must not dilute the other statistic)
struct NormArgs {
    float* obs;            // [n][od], may be null
    float* reward;         // [n], may be null
    const uint8_t* done;   // [n], may be null (no episode ends)
    float* ret;            // [n] discounted return per env
    double* acc;           // [NORM_ACC]
    int n, od;
    float gamma, eps, clip_obs, clip_reward;
};

__global__ void __launch_bounds__(NORM_BLOCK) norm_update_kernel(const NormArgs p) {
    const int i = blockIdx.x * NORM_BLOCK + threadIdx.x;
    double v[NORM_VALUES];
#pragma unroll
    for (int k = 0; k < NORM_VALUES; ++k) v[k] = 0.0;
    if (i < p.n) {
        if (p.obs) {
#pragma unroll
            for (int j = 0; j < NORM_MAX_OD; ++j)
                if (j < p.od) { const double x = (double)p.obs[(size_t)i * p.od + j]; v[j] = x; v[NORM_MAX_OD + j] = x * x; }
        }
        if (p.reward) {
            const float r = p.ret[i] * p.gamma + p.reward[i];
            v[2 * NORM_MAX_OD] = (double)r;
            v[2 * NORM_MAX_OD + 1] = (double)r * (double)r;
            p.ret[i] = (p.done && p.done[i]) ? 0.0f : r;
        }
    }
    // warp shuffle -> one row per warp in shared memory -> one thread per value adds the CTA's sum to the totals
    __shared__ double part[NORM_BLOCK / 32][NORM_VALUES];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NORM_VALUES; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) part[warp][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < NORM_VALUES) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NORM_BLOCK / 32; ++w) s += part[w][threadIdx.x];
        if (s != 0.0) atomicAdd(&p.acc[threadIdx.x], s);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (p.obs) atomicAdd(&p.acc[NORM_VALUES], (double)p.n);
        if (p.reward) atomicAdd(&p.acc[NORM_VALUES + 1], (double)p.n);
    }
}

__device__ __forceinline__ float norm_clip(float x, float c) { return x < -c ? -c : (x > c ? c : x); }

__global__ void __launch_bounds__(NORM_BLOCK) norm_apply_kernel(const NormArgs p) {
    const int i = blockIdx.x * NORM_BLOCK + threadIdx.x;
    if (i >= p.n) return;
    const double count = p.acc[NORM_VALUES], count_ret = p.acc[NORM_VALUES + 1];
    if (p.obs && count > 0.0) {   // nothing accumulated yet: leave the batch as it is
#pragma unroll
        for (int j = 0; j < NORM_MAX_OD; ++j)
            if (j < p.od) {
                const double mean = p.acc[j] / count;
                double var = p.acc[NORM_MAX_OD + j] / count - mean * mean;
                var = var > 0.0 ? var : 0.0;
                const size_t at = (size_t)i * p.od + j;
                p.obs[at] = norm_clip((float)(((double)p.obs[at] - mean) / sqrt(var + (double)p.eps)), p.clip_obs);
            }
    }
    if (p.reward && count_ret > 0.0) {
        const double mean = p.acc[2 * NORM_MAX_OD] / count_ret;
        double var = p.acc[2 * NORM_MAX_OD + 1] / count_ret - mean * mean;
        var = var > 0.0 ? var : 0.0;
        p.reward[i] = norm_clip((float)((double)p.reward[i] / sqrt(var + (double)p.eps)), p.clip_reward);
    }
}

}  // namespace gymcuda

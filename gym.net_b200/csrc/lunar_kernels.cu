// LunarLander kernels of libgymcuda: the generic step / reset / sample / ctor kernels (kernels.cuh) instantiated for
// the LunarLander traits (lunar.cuh over lunar_core.cuh), and their host launchers (lunar_launch.h).
// The fused k-step rollout of the other envs has no LunarLander instance: one lander step is a few hundred
// microseconds of dependent arithmetic, so a rollout is k step launches with the random policy sampled in the kernel
// (StepArgs::sample), each launch split over the two classes of the contact partition.
#include "lunar_launch.h"

#include "lunar.cuh"
#include "lunar_render.cuh"

namespace gymcuda {

static_assert(LUNAR_PAIRS_WORD == lunar::AUXD - 3, "kernels.cuh: aux word of the broad-phase pair list");
static_assert(LUNAR_STATE_DIM == LunarLander::SD && LUNAR_AUX_DIM == LunarLander::AUX, "lunar_launch.h: state layout");

template <bool C, bool P>
static void launch_step_t(bool auto_reset, bool limit, int grid, cudaStream_t s, const StepArgs& a) {
    using E = LunarLanderT<C, P>;
    if (auto_reset && limit) step_kernel<E, true, true><<<grid, STEP_BLOCK, 0, s>>>(a);
    else if (auto_reset) step_kernel<E, true, false><<<grid, STEP_BLOCK, 0, s>>>(a);
    else if (limit) step_kernel<E, false, true><<<grid, STEP_BLOCK, 0, s>>>(a);
    else step_kernel<E, false, false><<<grid, STEP_BLOCK, 0, s>>>(a);
}

template <bool C>
static void launch_step_trio(bool auto_reset, bool limit, int n, cudaStream_t s, const StepArgs& a) {
    using E = LunarLanderT<C, true, true>;
    const int warps = (n + TRIO_ENVS_PER_WARP - 1) / TRIO_ENVS_PER_WARP;
    const int grid = (warps * 32 + STEP_BLOCK - 1) / STEP_BLOCK;
    if (auto_reset && limit) step_kernel<E, true, true><<<grid, STEP_BLOCK, 0, s>>>(a);
    else if (auto_reset) step_kernel<E, true, false><<<grid, STEP_BLOCK, 0, s>>>(a);
    else if (limit) step_kernel<E, false, true><<<grid, STEP_BLOCK, 0, s>>>(a);
    else step_kernel<E, false, false><<<grid, STEP_BLOCK, 0, s>>>(a);
}

cudaError_t lunar_launch_step_trio(bool continuous, bool auto_reset, bool limit, cudaStream_t s, const StepArgs& a) {
    if (continuous) launch_step_trio<true>(auto_reset, limit, a.n, s, a); else launch_step_trio<false>(auto_reset, limit, a.n, s, a);
    return cudaGetLastError();
}

cudaError_t lunar_launch_step(bool continuous, bool has_pairs, bool auto_reset, bool limit, int grid, cudaStream_t s, const StepArgs& a) {
    if (continuous) { if (has_pairs) launch_step_t<true, true>(auto_reset, limit, grid, s, a); else launch_step_t<true, false>(auto_reset, limit, grid, s, a); }
    else { if (has_pairs) launch_step_t<false, true>(auto_reset, limit, grid, s, a); else launch_step_t<false, false>(auto_reset, limit, grid, s, a); }
    return cudaGetLastError();
}

cudaError_t lunar_launch_reset(bool continuous, int grid, cudaStream_t s, const ResetArgs& a) {
    if (continuous) reset_kernel<LunarLanderCont><<<grid, 128, 0, s>>>(a); else reset_kernel<LunarLander><<<grid, 128, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t lunar_launch_sample(bool continuous, int grid, cudaStream_t s, const SampleArgs& a) {
    if (continuous) sample_kernel<LunarLanderCont><<<grid, 128, 0, s>>>(a); else sample_kernel<LunarLander><<<grid, 128, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t lunar_launch_ctor(bool continuous, int grid, cudaStream_t s, const ResetArgs& a) {
    if (continuous) ctor_kernel<LunarLanderCont><<<grid, 128, 0, s>>>(a); else ctor_kernel<LunarLander><<<grid, 128, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t lunar_launch_render(cudaStream_t s, const RenderArgs& a) {
    const dim3 grid((unsigned)render_grid_x(a.width, a.height), (unsigned)a.count);
    render_lunar_kernel<<<grid, RENDER_BLOCK, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace gymcuda

#ifdef LUNAR_PHASE_CLOCKS
// timing probe only (see lunar_core.cuh): [2][10][65536] clock64 samples of the last launches
extern "C" int gymcuda_debug_lunar_phase(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, gymcuda::lunar::g_lunar_phase, sizeof(long long) * 2 * 10 * 65536);
}
#endif

// Host-side launchers of the LunarLander kernels.  They live in their own translation unit (lunar_kernels.cu): the
// rigid-body code is by far the largest part of the library and compiles in parallel with the rest.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace gymcuda {

constexpr int LUNAR_STATE_DIM = 80;   // float32 words per lander (lunar.cuh static_asserts these)
constexpr int LUNAR_AUX_DIM = 31;     // int32 words per lander in get_state / set_state (29 in HBM + episode step + episode ordinal)

// One env step.  has_pairs selects the kernel of the partition class (kernels.cuh "contact partition"); a.part / a.split
// restrict the launch to that class.  grid = blocks of STEP_BLOCK threads.
cudaError_t lunar_launch_step(bool continuous, bool has_pairs, bool auto_reset, bool limit, int grid, cudaStream_t stream, const StepArgs& a);
// The contact class stepped by three lanes per lander (lunar_core.cuh "TRIO"): sizes its own grid (ten landers per warp) from a.n.
cudaError_t lunar_launch_step_trio(bool continuous, bool auto_reset, bool limit, cudaStream_t stream, const StepArgs& a);
cudaError_t lunar_launch_reset(bool continuous, int grid, cudaStream_t stream, const ResetArgs& a);
cudaError_t lunar_launch_sample(bool continuous, int grid, cudaStream_t stream, const SampleArgs& a);
cudaError_t lunar_launch_ctor(bool continuous, int grid, cudaStream_t stream, const ResetArgs& a);
struct RenderArgs;
cudaError_t lunar_launch_render(cudaStream_t stream, const RenderArgs& a);   // LunarLanderEnv.Render (lunar_render.cuh)

}  // namespace gymcuda

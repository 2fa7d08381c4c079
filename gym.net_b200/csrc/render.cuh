// Headless batched Render(): what Env.Render draws, rasterised on the device for any subset of the batch.
//
//   LunarLanderEnv.Render   src/Gym.Environments/Envs/Aether/LunarLanderEnv.cs:776-890
//       black background (:790); the three lander polygons as 1-px OUTLINES in Color1 = (128, 102, 230) (:817-841, :201,243,278);
//       then the ground -- the quads between the terrain line and y = 0, which the reference calls _Sky (:544-556) -- FILLED
//       white (:844-855), so a part of the lander below the terrain line is painted over; the eleven moon edges as 1-px red
//       lines (:857-869); at both helipad ends a white 50-px pole and a yellow (204, 204, 0) chevron outline (:871-886).
//       Particles (:792-815) are not simulated by the engine (inert for the lander, lunar_core.cuh) and are not drawn.
//   CartPoleEnv.Render      src/Gym.Environments/Envs/Classic/CartPoleEnv.cs:69-135
//       white background, a 1-px black track at y = 300, the 50 x 30 black cart, the 10-px wide pole in (204, 153, 102)
//       rotated by theta about the axle, the axle disc of radius 5 in the same colour -- in that order (:118-131).
// World -> screen as the reference: x * SCALE, VIEWPORT_H - y * SCALE on a 600 x 400 canvas; a frame of another size is the same
// picture sampled on a coarser / finer grid (pixel centres mapped back to canvas coordinates), which is what an image-based
// agent that down-samples the 600 x 400 frame wants without ever materialising it.
// Coverage is decided at the pixel centre, a 1-px line is "distance to the segment <= 0.5 canvas pixels scaled to the output grid"
// -- no anti-aliasing: ImageSharp's edge blending (the reference's GraphicsOptions default) is NOT reproduced, so pixels on an
// edge differ from the reference's by construction; interior pixels and geometry are the reference's.  Parity unpinned (no .NET).
//
// One thread per four output pixels, blockIdx.y = frame; the ~45 segments of a frame are built once per CTA in shared memory.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "detmath.cuh"

namespace gymcuda {

struct RenderArgs {
    const void* state;        // env state (classic: Vec[n]; LunarLander: float[SD][n] field-major)
    const int32_t* env_ids;   // [count] or null = envs 0 .. count-1
    uint8_t* rgb;             // [count][height][width][3]
    int n, count, width, height;
};

constexpr int RENDER_BLOCK = 256;
constexpr float CANVAS_W = 600.0f, CANVAS_H = 400.0f;

struct Seg { float ax, ay, bx, by; };   // canvas coordinates

__device__ __forceinline__ float seg_dist2(const Seg& s, float px, float py) {
    const float dx = s.bx - s.ax, dy = s.by - s.ay;
    const float len2 = dx * dx + dy * dy;
    float t = len2 > 0.0f ? ((px - s.ax) * dx + (py - s.ay) * dy) / len2 : 0.0f;
    t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
    const float qx = s.ax + t * dx - px, qy = s.ay + t * dy - py;
    return qx * qx + qy * qy;
}

// Four consecutive pixels per thread: their 12 bytes leave as three 32-bit words (coalesced: 384 contiguous bytes per warp)
// when the frame's pixel count is a multiple of four, byte by byte otherwise.
struct Rgb { int r, g, b; };
template <class Shade>
__device__ __forceinline__ void render_quad(const RenderArgs& p, int frame, int quad, Shade shade) {
    const int total = p.width * p.height;
    const int first = quad * 4;
    if (first >= total) return;
    const float sx = CANVAS_W / (float)p.width, sy = CANVAS_H / (float)p.height;
    Rgb c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int pix = first + k < total ? first + k : total - 1;
        c[k] = shade(((float)(pix % p.width) + 0.5f) * sx, ((float)(pix / p.width) + 0.5f) * sy);
    }
    uint8_t* out = p.rgb + ((size_t)frame * total + first) * 3;
    if ((total & 3) == 0 && (reinterpret_cast<uintptr_t>(p.rgb) & 3) == 0) {
        uint32_t* w = reinterpret_cast<uint32_t*>(out);
        w[0] = (uint32_t)c[0].r | ((uint32_t)c[0].g << 8) | ((uint32_t)c[0].b << 16) | ((uint32_t)c[1].r << 24);
        w[1] = (uint32_t)c[1].g | ((uint32_t)c[1].b << 8) | ((uint32_t)c[2].r << 16) | ((uint32_t)c[2].g << 24);
        w[2] = (uint32_t)c[2].b | ((uint32_t)c[3].r << 8) | ((uint32_t)c[3].g << 16) | ((uint32_t)c[3].b << 24);
    } else {
        for (int k = 0; k < 4 && first + k < total; ++k) { out[3 * k] = (uint8_t)c[k].r; out[3 * k + 1] = (uint8_t)c[k].g; out[3 * k + 2] = (uint8_t)c[k].b; }
    }
}
__host__ __device__ inline int render_grid_x(int width, int height) { return ((width * height + 3) / 4 + RENDER_BLOCK - 1) / RENDER_BLOCK; }

// ---------------------------------------------------------------- CartPole
static __global__ void __launch_bounds__(RENDER_BLOCK) render_cartpole_kernel(const RenderArgs p) {
    const int frame = blockIdx.y;
    const int e = p.env_ids ? p.env_ids[frame] : frame;
    const float4 s = reinterpret_cast<const float4*>(p.state)[e];
    const float scale = CANVAS_W / (2.4f * 2.0f);               // :73-74 screen_width / world_width
    const float carty = 300.0f, polewidth = 10.0f, poleheight = scale * (2.0f * 0.5f), cartwidth = 50.0f, cartheight = 30.0f;
    const float center_x = s.x * scale + CANVAS_W / 2.0f;       // :111
    float sn, cs;
    sincosf_det(s.z, &sn, &cs);
    render_quad(p, frame, blockIdx.x * RENDER_BLOCK + threadIdx.x, [&](float cx, float cy) {
        int r = 255, g = 255, b = 255;                              // :121 white background
        if (cy >= carty && cy < carty + 1.0f) { r = g = b = 0; }    // :122 the track
        const float lx = cx - center_x;
        if (lx >= -cartwidth / 2 && lx < cartwidth / 2 && cy >= carty - cartheight / 2 && cy < carty + cartheight / 2) { r = g = b = 0; }   // :106, :116
        // pole: the rectangle (-5, carty - poleheight, 10, poleheight) rotated by theta about the axle (0, carty - 5) (:96-97, :117)
        const float pivot_y = carty - polewidth / 2;
        const float dx = lx, dy = cy - pivot_y;
        const float ux = cs * dx + sn * dy, uy = -sn * dx + cs * dy + pivot_y;   // inverse rotation (Matrix3x2.CreateRotation is clockwise on a y-down canvas)
        if (ux >= -polewidth / 2 && ux < polewidth / 2 && uy >= carty - poleheight && uy < carty) { r = 204; g = 153; b = 102; }
        if (dx * dx + dy * dy <= (polewidth / 2) * (polewidth / 2)) { r = 204; g = 153; b = 102; }   // :118 the axle disc
        return Rgb{r, g, b};
    });
}

}  // namespace gymcuda

// LunarLanderEnv.Render on the device (see render.cuh for what is drawn and how coverage is decided).
#pragma once
#include "render.cuh"
#include "lunar_core.cuh"

namespace gymcuda {

// ---------------------------------------------------------------- LunarLander
// geometry from lunar_core.cuh (SHAPES); state layout from lunar.cuh (load_lander): body i at words 7 i = (c.x, c.y, a, ...),
// terrain heights at words 21 + 8 + 4 MAXC .. + 10
static __global__ void __launch_bounds__(RENDER_BLOCK) render_lunar_kernel(const RenderArgs p) {
    constexpr int BODY_C0 = 0, TERRAIN0 = 21 + 8 + 4 * lunar::MAXC;
    __shared__ Seg lander[14];
    __shared__ Seg moon[11];
    __shared__ Seg flag[8];
    __shared__ float ty[11];
    const int frame = blockIdx.y;
    const int e = p.env_ids ? p.env_ids[frame] : frame;
    const float* st = reinterpret_cast<const float*>(p.state);
    const size_t n = (size_t)p.n;
    constexpr float S = 30.0f, W = 600.0f / 30.0f, H = 400.0f / 30.0f;
    if (threadIdx.x < 11) ty[threadIdx.x] = st[(size_t)(TERRAIN0 + (int)threadIdx.x) * n + e];
    if (threadIdx.x < 3) {   // one thread per body: its outline in canvas coordinates (:817-841)
        const int bi = (int)threadIdx.x;
        const float bcx = st[(size_t)(BODY_C0 + 7 * bi) * n + e], bcy = st[(size_t)(BODY_C0 + 7 * bi + 1) * n + e], ba = st[(size_t)(BODY_C0 + 7 * bi + 2) * n + e];
        float sn, cs;
        sincosf_det(ba, &sn, &cs);
        const float px = bcx - (cs * lunar::SHAPES[bi].centroid.x - sn * lunar::SHAPES[bi].centroid.y), py = bcy - (sn * lunar::SHAPES[bi].centroid.x + cs * lunar::SHAPES[bi].centroid.y);   // body origin = c - R * centroid
        const int base = bi == 0 ? 0 : (bi == 1 ? 6 : 10);
        const int cnt = lunar::SHAPES[bi].count;
        for (int k = 0; k < cnt; ++k) {
            const int k2 = k + 1 < cnt ? k + 1 : 0;
            const float x1 = (cs * lunar::SHAPES[bi].v[k].x - sn * lunar::SHAPES[bi].v[k].y) + px, y1 = (sn * lunar::SHAPES[bi].v[k].x + cs * lunar::SHAPES[bi].v[k].y) + py;
            const float x2 = (cs * lunar::SHAPES[bi].v[k2].x - sn * lunar::SHAPES[bi].v[k2].y) + px, y2 = (sn * lunar::SHAPES[bi].v[k2].x + cs * lunar::SHAPES[bi].v[k2].y) + py;
            lander[base + k] = Seg{x1 * S, CANVAS_H - y1 * S, x2 * S, CANVAS_H - y2 * S};
        }
    }
    __syncthreads();
    if (threadIdx.x < 10) {   // terrain edges (:545-557, :857-869)
        const float cw = W / 10.0f;
        const int i = (int)threadIdx.x;
        moon[i] = Seg{cw * (float)i * S, CANVAS_H - ty[i] * S, cw * (float)(i + 1) * S, CANVAS_H - ty[i + 1] * S};
    }
    if (threadIdx.x == 10) moon[10] = Seg{0.0f, CANVAS_H, W * S, CANVAS_H};   // the base edge (0,0)-(W,0) (:541)
    if (threadIdx.x >= 16 && threadIdx.x < 18) {   // helipad flags at chunk_x[4] and chunk_x[6], helipad_y = H / 4 (:514-522, :871-886)
        const int f = (int)threadIdx.x - 16;
        const float x1 = (W / 10.0f) * (f == 0 ? 4.0f : 6.0f) * S;
        const float y1 = CANVAS_H - (H / 4.0f) * S, y2 = y1 - 50.0f;
        flag[4 * f + 0] = Seg{x1, y1, x1, y2};                         // pole
        flag[4 * f + 1] = Seg{x1, y2, x1, y2 + 10.0f};                 // chevron
        flag[4 * f + 2] = Seg{x1, y2 + 10.0f, x1 + 25.0f, y2 + 5.0f};
        flag[4 * f + 3] = Seg{x1 + 25.0f, y2 + 5.0f, x1, y2};
    }
    __syncthreads();
    const float sx = CANVAS_W / (float)p.width, sy = CANVAS_H / (float)p.height;
    const float half = 0.5f * (sx > sy ? sx : sy);   // a 1-px line of the 600 x 400 canvas stays at least one OUTPUT pixel wide
    const float h2 = half * half;
    // Culling: the CTA's 1024 consecutive pixels span a few rows; a segment whose box (grown by the line's half width) misses
    // that band of canvas rows cannot colour any of them and is skipped by the whole CTA (conservative: same picture).
    __shared__ unsigned live_lander, live_moon, live_flag;
    if (threadIdx.x == 0) {
        const int total = p.width * p.height;
        const int first = blockIdx.x * RENDER_BLOCK * 4;
        const int last = first + RENDER_BLOCK * 4 - 1 < total - 1 ? first + RENDER_BLOCK * 4 - 1 : total - 1;
        const float y_lo = ((float)(first / p.width) + 0.5f) * sy - half, y_hi = ((float)(last / p.width) + 0.5f) * sy + half;
        unsigned ml = 0u, mm = 0u, mf = 0u;
        for (int k = 0; k < 14; ++k) { const float a = lander[k].ay, c = lander[k].by; if ((a < c ? a : c) <= y_hi && (a > c ? a : c) >= y_lo) ml |= 1u << k; }
        for (int k = 0; k < 11; ++k) { const float a = moon[k].ay, c = moon[k].by; if ((a < c ? a : c) <= y_hi && (a > c ? a : c) >= y_lo) mm |= 1u << k; }
        for (int k = 0; k < 8; ++k) { const float a = flag[k].ay, c = flag[k].by; if ((a < c ? a : c) <= y_hi && (a > c ? a : c) >= y_lo) mf |= 1u << k; }
        live_lander = ml; live_moon = mm; live_flag = mf;
    }
    __syncthreads();
    const unsigned ll = live_lander, lm = live_moon, lf = live_flag;
    render_quad(p, frame, blockIdx.x * RENDER_BLOCK + threadIdx.x, [&](float cx, float cy) {
        int r = 0, g = 0, b = 0;                                         // :790 space is black
#pragma unroll 1
        for (int k = 0; k < 14; ++k) if (((ll >> k) & 1u) && seg_dist2(lander[k], cx, cy) <= h2) { r = 128; g = 102; b = 230; }
        {   // the ground below the terrain line, white (:844-855)
            const float cwp = (W / 10.0f) * S;
            int i = (int)(cx / cwp);
            i = i < 0 ? 0 : (i > 9 ? 9 : i);
            const Seg m = moon[i];
            const float t = (cx - m.ax) / (m.bx - m.ax);
            const float yline = m.ay + t * (m.by - m.ay);
            if (cy >= yline) { r = g = b = 255; }
        }
#pragma unroll 1
        for (int k = 0; k < 11; ++k) if (((lm >> k) & 1u) && seg_dist2(moon[k], cx, cy) <= h2) { r = 255; g = 0; b = 0; }
#pragma unroll 1
        for (int f = 0; f < 2; ++f) {
            if (((lf >> (4 * f)) & 1u) && seg_dist2(flag[4 * f], cx, cy) <= h2) { r = g = b = 255; }
            for (int k = 1; k < 4; ++k) if (((lf >> (4 * f + k)) & 1u) && seg_dist2(flag[4 * f + k], cx, cy) <= h2) { r = 204; g = 204; b = 0; }
        }
        return Rgb{r, g, b};
    });
}

}  // namespace gymcuda

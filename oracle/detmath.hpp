// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// "detmath v1": the fp32 elementary functions of the engine's storage-precision arithmetic,
// written so that every operation is a single IEEE-754 binary32 (or binary64) operation with
// round-to-nearest-even: +, -, *, /, sqrt, fma, rint (binary64 only).  The same sequence of operations on the
// GPU (compiled with -fmad=false, explicit fmaf) gives the same bits, which is what lets the
// F32 mode of this oracle be compared BIT-FOR-BIT with the CUDA kernels on free-running rollouts.
// Compile this file with -ffp-contract=off.
//
// Accuracy (tests/test_oracle_detmath.py): sin/cos <= 2 ulp(fp32) for |x| <= 1e5.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace oracle { namespace det {

static const float PIO4_F       = 0.7853981852531433f;
static const float TWO_OVER_PI  = 0.6366197466850281f;
static const float PIO2_1       = 1.5707963705062866f;       // fl32(pi/2)
static const float PIO2_2       = -4.371138828673793e-08f;   // fl32(pi/2 - PIO2_1)
static const float PIO2_3       = -1.7151245100058819e-15f;  // fl32(pi/2 - PIO2_1 - PIO2_2)
static const double TWO_OVER_PI_D = 0.6366197723675814;
static const double PIO2_HI_D   = 1.5707963267948966;
static const double PIO2_LO_D   = 6.123233995736766e-17;

// minimax-style coefficients on [-pi/4, pi/4] (Cephes single-precision set)
static const float S1 = -1.6666654611e-1f, S2 = 8.3321608736e-3f, S3 = -1.9515295891e-4f;
static const float C1 = 4.166664568298827e-2f, C2 = -1.388731625493765e-3f, C3 = 2.443315711809948e-5f;

inline void sincosf_det(float x, float* s, float* c) {
    float ax = std::fabs(x);
    float r;
    int q;
    if (ax <= PIO4_F) {
        r = x; q = 0;
    } else if (ax <= 32768.0f) {
        // quadrant = x * 2/pi rounded ONCE to an integer (ties to even) by adding 1.5 * 2^23 inside the fma
        const float MAGIC = 12582912.0f;
        float t = std::fma(x, TWO_OVER_PI, MAGIC);
        float fq = t - MAGIC;
        r = std::fma(fq, -PIO2_1, x);
        r = std::fma(fq, -PIO2_2, r);
        r = std::fma(fq, -PIO2_3, r);
        q = (int)fq;
    } else if (ax <= 1.0e14f) {
        double dq = std::rint((double)x * TWO_OVER_PI_D);
        double dr = std::fma(dq, -PIO2_HI_D, (double)x);
        dr = std::fma(dq, -PIO2_LO_D, dr);
        r = (float)dr;
        q = (int)((long long)dq & 3);
    } else {
        *s = *c = std::numeric_limits<float>::quiet_NaN();
        return;
    }
    float r2 = r * r;
    float ps = std::fma(S3, r2, S2);
    ps = std::fma(ps, r2, S1);
    float sp = std::fma(r * r2, ps, r);
    float pc = std::fma(C3, r2, C2);
    pc = std::fma(pc, r2, C1);
    float cp = std::fma(r2 * r2, pc, std::fma(-0.5f, r2, 1.0f));
    switch (q & 3) {
        case 0: *s = sp;  *c = cp;  break;
        case 1: *s = cp;  *c = -sp; break;
        case 2: *s = -sp; *c = -cp; break;
        default: *s = -cp; *c = sp; break;
    }
}

inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

}}  // namespace oracle::det

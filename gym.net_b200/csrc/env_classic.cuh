// Per-env device dynamics of the classic-control family.  One thread owns one env instance.
//
// Each env is a traits struct used by the generic step / reset / rollout kernels (kernels.cuh):
//   SD, OD, AD, ACTN          state / observation / action widths, Discrete(n) or 0 for Box
//   Vec                       the storage vector type of one env's state in HBM (float4 / float2)
//   S                         register state
//   load/store                one coalesced vector load/store per env
//   reset(S&, Block)          Philox block -> initial state
//   step(S&, Act, sbd)        transition + reward + termination
//   obs(S, float*)            observation (OD floats)
//
// Arithmetic ("engine arithmetic"): float32 state and float32 math built from single IEEE
// operations (detmath.cuh), with the termination test refined in double wherever float32 rounding
// could flip it, so that `done` equals the reference's double-precision evaluation from the same
// stored state.  Compiled with -fmad=false; fused operations are explicit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "detmath.cuh"
#include "philox.cuh"

namespace gymcuda {

struct StepOut { float reward; unsigned done; unsigned did_reset = 0u; };   // done: 0 / 1 in a full register (a byte-sized bool costs PRMT packing around calls); did_reset: the env already replaced its state by a new episode (LunarLander's fused crash reset)

// constructor arguments of the env (LunarLanderEnv.cs:381); unused by the classic-control family
struct EnvParams { float gravity, wind_power, turbulence_power; int32_t use_wind; };

// ------------------------------------------------------------------------------------------------
// CartPole: src/Gym.Environments/Envs/Classic/CartPoleEnv.cs
// ------------------------------------------------------------------------------------------------
struct CartPole {
    static constexpr int SD = 4, OD = 4, AD = 1, ACTN = 2, DEFAULT_LIMIT = 0;
    static constexpr bool HAS_SBD = true;     // steps_beyond_done (CartPoleEnv.cs:41)
    static constexpr bool PREGEN_RESET = true;   // rollout kernel pre-generates the next initial state
    static constexpr bool ROLLOUT_CHUNK = true;  // rollout kernel runs unrolled 8-step chunks
    static constexpr int AUXW = 0;               // extra int32 words per env in HBM
    static constexpr bool REJECT_INVALID = false;  // Debug.Assert only (CartPoleEnv.cs:139)
    using Vec = float4;
    using Act = int32_t;
    struct S { float x, x_dot, theta, theta_dot; };

    // C# `const float` values (CartPoleEnv.cs:24-36), folded in float32 by the C# compiler.
    static constexpr float GRAVITY = 9.8f;
    static constexpr float FORCE_MAG = 10.0f;
    static constexpr float TAU = 0.02f;
    static constexpr float POLEMASS_LENGTH = 0.05000000074505806f;   // 0.1f * 0.5f
    static constexpr float X_THRESHOLD = 2.4f;
    static constexpr float THETA_THRESHOLD = 0.20943951606750488f;   // (float)(12*2*pi/360)
    // derived, rounded once to float32
    static constexpr float INV_TOTAL_MASS = 0.9090908765792847f;     // 1 / 1.100000023841858
    static constexpr float K0 = 0.6666666865348816f;                 // length * 4/3
    static constexpr float K1 = 0.04545454680919647f;                // length * masspole / total_mass
    static constexpr float PML_OVER_M = 0.04545454680919647f;        // polemass_length / total_mass

    __device__ static __forceinline__ S load(const void* base, const int32_t*, int, int i, const EnvParams&) {
        const float4 v = reinterpret_cast<const float4*>(base)[i];
        return S{v.x, v.y, v.z, v.w};
    }
    __device__ static __forceinline__ void store(void* base, int32_t*, int, int i, const S& s) {
        reinterpret_cast<float4*>(base)[i] = make_float4(s.x, s.x_dot, s.theta, s.theta_dot);
    }
    // CartPoleEnv.cs:65  state = uniform(-0.05, 0.05, 4)
    __device__ static __forceinline__ void reset(S& s, uint64_t seed, uint32_t gid, uint32_t ordinal, uint64_t, const EnvParams&) {
        const Block b = draw(seed, gid, (uint64_t)ordinal, STREAM_RESET);
        s.x = uniformf(-0.05f, 0.05f, b.w0);
        s.x_dot = uniformf(-0.05f, 0.05f, b.w1);
        s.theta = uniformf(-0.05f, 0.05f, b.w2);
        s.theta_dot = uniformf(-0.05f, 0.05f, b.w3);
    }
    __device__ static __forceinline__ bool valid(Act a) { return a == 0 || a == 1; }

    // The reference's own double-precision termination test (:154,:156,:167) from the float32 state; reached only
    // when a float32 position lands exactly on a threshold (see step)
    __device__ static __noinline__ unsigned done_f64(float x, float x_dot, float theta, float theta_dot) {
        const double xd = (double)x + (double)TAU * (double)x_dot;                     // :154
        const double thd = (double)theta + (double)TAU * (double)theta_dot;            // :156
        return (unsigned)((fabs(xd) > (double)X_THRESHOLD) | (fabs(thd) > (double)THETA_THRESHOLD));  // :167
    }

    // CartPoleEnv.cs:137-186.  Accelerations (:146-151) in float32.  The position updates (:154,:156) are ONE
    // float32 fma each: tau * x_dot is exact in double (24 x 24 bits), so fmaf(tau, x_dot, x) is the reference's
    // double sum e = x + tau * x_dot rounded once to float32 (the reference rounds it to double first: the two
    // differ only by double rounding, <= 1 ulp32, far inside 1e-5).  Termination (:167) compares against
    // float32-representable thresholds T, and rounding is monotonic: fl32(e) > T implies fl64(e) > T and
    // fl32(e) < T implies fl64(e) < T.  Only fl32(e) == T is undecided in float32 and goes to done_f64, so
    // `done` is exactly the reference's double-precision flag from the same float32 state.
    // SMALL: the caller guarantees small_ok(s): sincos is the bare polynomial, and the quotient :150 is the
    // six-instruction core of the IEEE division without its exponent-range check (div_inrange): with
    // |theta| <= pi/4 and |theta_dot| <= 70, den is in [0.62, 0.67] and num is 0 or in [2^-60, 2^8], where that
    // core IS the correctly rounded quotient -- both variants return the same bits.
    template <bool SMALL = false>
    __device__ static __forceinline__ StepOut step(S& s, Act a, int32_t& sbd, uint64_t, uint32_t, uint64_t) {
        const float force = (a == 1) ? FORCE_MAG : -FORCE_MAG;                         // :146
        float sn, cs;
        if (SMALL) sincos_poly(s.theta, &sn, &cs);
        else sincosf_det(s.theta, &sn, &cs);                                           // :147-148
        const float t1 = (POLEMASS_LENGTH * s.theta_dot) * s.theta_dot;
        const float temp = fmaf(t1, sn, force) * INV_TOTAL_MASS;                       // :149
        const float den = fmaf(-K1, cs * cs, K0);
        const float num = fmaf(GRAVITY, sn, -(cs * temp));
        const float thetaacc = SMALL ? div_inrange(num, den) : num / den;              // :150
        const float xacc = fmaf(-(PML_OVER_M * thetaacc), cs, temp);                   // :151
        const float nx = fmaf(TAU, s.x_dot, s.x);                                      // :154
        const float nth = fmaf(TAU, s.theta_dot, s.theta);                             // :156
        unsigned done = 0;
        if ((fabsf(nx) >= X_THRESHOLD) | (fabsf(nth) >= THETA_THRESHOLD)) {             // :167
            done = (unsigned)((fabsf(nx) > X_THRESHOLD) | (fabsf(nth) > THETA_THRESHOLD));
            if (!done) done = done_f64(s.x, s.x_dot, s.theta, s.theta_dot);            // exactly on a threshold
        }
        s.x_dot = fmaf(TAU, xacc, s.x_dot);                                            // :155
        s.theta_dot = fmaf(TAU, thetaacc, s.theta_dot);                                // :157
        s.x = nx;
        s.theta = nth;
        float reward = 1.0f;                                                           // :170,:174
        if (done) {
            if (sbd == -1) sbd = 0;                                                    // :173
            else { sbd += 1; reward = 0.0f; }                                          // :181-182
        }
        return StepOut{reward, done};
    }
    // rollout fast path: true when step<true> is valid for this state and, with auto-reset, for every later one.
    // A non-terminal state has |theta| <= 0.2095 and a reset state |theta|, |theta_dot| <= 0.05; a step changes
    // theta_dot by < 6 while |theta_dot| <= 70, and a state entered with |theta_dot| > 21 and |theta| <= 0.2095
    // terminates at once (theta moves by 0.02 * theta_dot), so from |theta_dot| <= 64 no later entry exceeds 70.
    static constexpr bool HAS_SMALL = true;
    __device__ static __forceinline__ bool small_ok(const S& s) {
        return (fabsf(s.theta) <= 0.7853981852531433f) & (fabsf(s.theta_dot) <= 64.0f);
    }
    __device__ static __forceinline__ void obs(const S& s, float* o) {
        o[0] = s.x; o[1] = s.x_dot; o[2] = s.theta; o[3] = s.theta_dot;                // :166,:185
    }
};

// ------------------------------------------------------------------------------------------------
// Pendulum-v1 (not in the reference, README.md:76; spec = upstream gym 0.26 pendulum.py)
// ------------------------------------------------------------------------------------------------
static __device__ __noinline__ float fmodf_cold(float a, float b) { return fmodf(a, b); }

// Python float `a % b` for b > 0 (upstream angle_normalize): fmod is exact, so the result is defined by IEEE
// alone.  For |a| <= 2^21 (every angle a pendulum reaches) fmod is computed without branches or conversions:
// the quotient is rounded to the nearest integer with the 1.5 * 2^23 magic number (it is floor(|a| / b) or one
// more), the remainder |a| - q * b is ONE fma -- exact whenever q is the true quotient, because the true
// remainder is representable -- and a negative remainder steps q down and recomputes.  Beyond: CUDA's fmodf.
template <bool IN_RANGE = false>   // IN_RANGE: the caller guarantees |a| <= 2^21
__device__ __forceinline__ float py_mod_2pi(float a) {
    constexpr float b = 6.2831854820251465f;          // fl32(2 pi)
    constexpr float INV_B = 0.15915493667125702f;     // fl32(1 / b): only steers the quotient estimate
    constexpr float MAGIC = 12582912.0f;
    const float ax = fabsf(a);
    float r;
    if (IN_RANGE || ax <= 2097152.0f) {
        float q = fmaf(ax, INV_B, MAGIC) - MAGIC;
        r = fmaf(-q, b, ax);
        if (r < 0.0f) { q = q - 1.0f; r = fmaf(-q, b, ax); }
    } else {
        r = fmodf_cold(ax, b);
    }
    float m = (a < 0.0f) ? -r : r;               // fmod carries the sign of a
    if (m != 0.0f) { if (m < 0.0f) m += b; } else m = 0.0f;
    return m;
}

struct Pendulum {
    static constexpr int SD = 2, OD = 3, AD = 1, ACTN = 0, DEFAULT_LIMIT = 200;
    static constexpr bool HAS_SBD = false;
    static constexpr bool PREGEN_RESET = true;
    static constexpr int AUXW = 0;
    static constexpr bool REJECT_INVALID = true;
    static constexpr bool HAS_SMALL = true;         // step<true>: |theta| known to be far inside the branch-free ranges
    static constexpr bool ROLLOUT_CHUNK = true;     // rollout kernel runs unrolled 8-step chunks
    static constexpr float ACT_LOW = -2.0f, ACT_HIGH = 2.0f;
    using Vec = float2;
    using Act = float;
    // sn, cs = sin/cos of th, carried in registers: the observation of step t and the dynamics of step t + 1
    // need the same pair, so a fused rollout evaluates sincos once per step instead of twice
    struct S { float th, thdot, sn, cs; };
    __device__ static __forceinline__ S load(const void* base, const int32_t*, int, int i, const EnvParams&) {
        const float2 v = reinterpret_cast<const float2*>(base)[i];
        S s{v.x, v.y, 0.0f, 1.0f};
        sincosf_det(s.th, &s.sn, &s.cs);
        return s;
    }
    __device__ static __forceinline__ void store(void* base, int32_t*, int, int i, const S& s) {
        reinterpret_cast<float2*>(base)[i] = make_float2(s.th, s.thdot);
    }
    __device__ static __forceinline__ void reset(S& s, uint64_t seed, uint32_t gid, uint32_t ordinal, uint64_t, const EnvParams&) {
        const Block b = draw(seed, gid, (uint64_t)ordinal, STREAM_RESET);
        constexpr float PI_F = 3.1415927410125732f;
        s.th = uniformf(-PI_F, PI_F, b.w0);
        s.thdot = uniformf(-1.0f, 1.0f, b.w1);
        sincosf_det(s.th, &s.sn, &s.cs);
    }
    __device__ static __forceinline__ bool valid(Act a) { return a == a; }
    // rollout fast path: theta moves by at most 8 * 0.05 per step, so from |theta| <= 30000 the eight steps of a
    // chunk stay inside the branch-free ranges of sincos (32768) and fmod (2^21); a reset lands in [-pi, pi]
    __device__ static __forceinline__ bool small_ok(const S& s) { return fabsf(s.th) <= 30000.0f; }
    template <bool SMALL = false>
    __device__ static __forceinline__ StepOut step(S& s, Act a, int32_t&, uint64_t, uint32_t, uint64_t) {
        constexpr float PI_F = 3.1415927410125732f;
        const float th = s.th, thdot = s.thdot;
        const float u = clampf(a, -2.0f, 2.0f);
        const float an = py_mod_2pi<SMALL>(th + PI_F) - PI_F;
        const float costs = (an * an + 0.1f * (thdot * thdot)) + 0.001f * (u * u);
        float newthdot = thdot + (15.0f * s.sn + 3.0f * u) * 0.05f;
        newthdot = clampf(newthdot, -8.0f, 8.0f);
        s.th = th + newthdot * 0.05f;
        s.thdot = newthdot;
        sincosf_det<SMALL>(s.th, &s.sn, &s.cs);
        return StepOut{-costs, 0u};
    }
    __device__ static __forceinline__ void obs(const S& s, float* o) { o[0] = s.cs; o[1] = s.sn; o[2] = s.thdot; }
};

// ------------------------------------------------------------------------------------------------
// MountainCar-v0 / MountainCarContinuous-v0 (not in the reference, README.md:74-75)
// ------------------------------------------------------------------------------------------------
template <bool CONTINUOUS>
struct MountainCarT {
    static constexpr int SD = 2, OD = 2, AD = 1, ACTN = CONTINUOUS ? 0 : 3;
    static constexpr int DEFAULT_LIMIT = CONTINUOUS ? 999 : 200;
    static constexpr bool HAS_SBD = false;
    static constexpr bool PREGEN_RESET = true;
    static constexpr int AUXW = 0;
    static constexpr bool REJECT_INVALID = true;
    static constexpr bool HAS_SMALL = true;         // step<true>: position known to be in range for the branch-free sincos
    static constexpr bool ROLLOUT_CHUNK = true;     // rollout kernel runs unrolled 8-step chunks
    static constexpr float ACT_LOW = -1.0f, ACT_HIGH = 1.0f;
    using Vec = float2;
    using Act = typename std::conditional<CONTINUOUS, float, int32_t>::type;
    struct S { float position, velocity; };
    __device__ static __forceinline__ S load(const void* base, const int32_t*, int, int i, const EnvParams&) {
        const float2 v = reinterpret_cast<const float2*>(base)[i];
        return S{v.x, v.y};
    }
    __device__ static __forceinline__ void store(void* base, int32_t*, int, int i, const S& s) {
        reinterpret_cast<float2*>(base)[i] = make_float2(s.position, s.velocity);
    }
    __device__ static __forceinline__ void reset(S& s, uint64_t seed, uint32_t gid, uint32_t ordinal, uint64_t, const EnvParams&) {
        const Block b = draw(seed, gid, (uint64_t)ordinal, STREAM_RESET);
        s.position = uniformf(-0.6f, -0.4f, b.w0);
        s.velocity = 0.0f;
    }
    __device__ static __forceinline__ bool valid(Act a) {
        if (CONTINUOUS) return a == a;
        return a >= 0 && a < 3;
    }
    // upstream double-precision step, used only to refine `done` next to a threshold
    __device__ static __noinline__ unsigned done_f64(float position0, float velocity0, Act a) {
        double position = (double)position0, velocity = (double)velocity0;
        if (CONTINUOUS) {
            double force = (double)a;
            force = force < -1.0 ? -1.0 : (force > 1.0 ? 1.0 : force);
            velocity += force * 0.0015 - 0.0025 * cos(3 * position);
        } else {
            velocity += ((int)a - 1) * 0.001 + cos(3 * position) * (-0.0025);
        }
        velocity = velocity < -0.07 ? -0.07 : (velocity > 0.07 ? 0.07 : velocity);
        position += velocity;
        position = position < -1.2 ? -1.2 : (position > 0.6 ? 0.6 : position);
        if (position == -1.2 && velocity < 0) velocity = 0;
        return (unsigned)(position >= (CONTINUOUS ? 0.45 : 0.5) && velocity >= 0.0);
    }
    // rollout fast path: every step clamps the position to [-1.2, 0.6] and a reset draws it from [-0.6, -0.4]
    __device__ static __forceinline__ bool small_ok(const S& s) { return fabsf(s.position) <= 10000.0f; }
    template <bool SMALL = false>
    __device__ static __forceinline__ StepOut step(S& s, Act a, int32_t&, uint64_t, uint32_t, uint64_t) {
        constexpr float MIN_POS = -1.2f, MAX_POS = 0.6f, MAX_SPEED = 0.07f;
        constexpr float GOAL = CONTINUOUS ? 0.45f : 0.5f;
        const float position = s.position, velocity = s.velocity;
        float sn, cs;
        sincosf_det<SMALL>(3.0f * position, &sn, &cs);
        float push;
        if (CONTINUOUS) push = clampf((float)a, -1.0f, 1.0f) * 0.0015f;
        else push = (float)((int)a - 1) * 0.001f;
        float nv = velocity + (push + cs * (-0.0025f));
        nv = clampf(nv, -MAX_SPEED, MAX_SPEED);
        float np = position + nv;
        np = clampf(np, MIN_POS, MAX_POS);
        if (np == MIN_POS && nv < 0.0f) nv = 0.0f;
        // everything about termination sits behind ONE rarely taken test: below goal - 1e-6 the float32 flag is 0
        // and the double evaluation (which differs from float32 by < 1e-7 in position) agrees
        unsigned done = 0;
        if (np >= GOAL - 1e-6f) {
            done = (unsigned)((np >= GOAL) & (nv >= 0.0f));
            if (fabsf(np - GOAL) <= 1e-6f || fabsf(nv) <= 1e-7f) done = done_f64(position, velocity, a);
        }
        s.position = np;
        s.velocity = nv;
        float reward;
        if (CONTINUOUS) {
            reward = done ? 100.0f : 0.0f;
            reward = reward - ((float)a * (float)a) * 0.1f;
        } else {
            reward = -1.0f;
        }
        return StepOut{reward, done};
    }
    __device__ static __forceinline__ void obs(const S& s, float* o) { o[0] = s.position; o[1] = s.velocity; }
};
using MountainCar = MountainCarT<false>;
using MountainCarCont = MountainCarT<true>;

// ------------------------------------------------------------------------------------------------
// Acrobot-v1 (not in the reference, README.md:73; upstream acrobot.py, "book" dynamics, RK4, dt 0.2)
// ------------------------------------------------------------------------------------------------
// upstream double-precision dynamics (used to refine `done` next to its threshold); the float32 engine arithmetic is
// acrobot_dsdt_f32 / acrobot_rk4_f32 below
template <class R> struct AcroMath;
template <> struct AcroMath<double> {
    __device__ static __forceinline__ void sc(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
    __device__ static __forceinline__ double cos_minus_half_pi(double x) { return cos(x - 3.14159265358979323846 / 2.0); }
};
// algebra of dsdt given sin/cos(theta2), cos(theta1 + theta2 - pi/2) and cos(theta1 - pi/2)
template <class R>
__device__ __forceinline__ void acrobot_dsdt_trig(const R s[4], R a, R s2, R c2, R cmh12, R cmh1, R out[4]) {
    const R m1 = 1, m2 = 1, l1 = 1, lc1 = R(0.5), lc2 = R(0.5), I1 = 1, I2 = 1, g = R(9.8);
    const R dtheta1 = s[2], dtheta2 = s[3];
    const R d1 = m1 * lc1 * lc1 + m2 * (l1 * l1 + lc2 * lc2 + 2 * l1 * lc2 * c2) + I1 + I2;
    const R d2 = m2 * (lc2 * lc2 + l1 * lc2 * c2) + I2;
    const R phi2 = m2 * lc2 * g * cmh12;
    const R phi1 = -m2 * l1 * lc2 * dtheta2 * dtheta2 * s2 - 2 * m2 * l1 * lc2 * dtheta2 * dtheta1 * s2 +
                   (m1 * lc1 + m2 * l1) * g * cmh1 + phi2;
    const R ddtheta2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * dtheta1 * dtheta1 * s2 - phi2) /
                       (m2 * lc2 * lc2 + I2 - d2 * d2 / d1);
    const R ddtheta1 = -(d2 * ddtheta2 + phi1) / d1;
    out[0] = dtheta1; out[1] = dtheta2; out[2] = ddtheta1; out[3] = ddtheta2;
}

template <class R>
__device__ __forceinline__ void acrobot_dsdt(const R s[4], R a, R out[4]) {
    R s2, c2;
    AcroMath<R>::sc(s[1], &s2, &c2);
    const R cmh12 = AcroMath<R>::cos_minus_half_pi(s[0] + s[1]);
    const R cmh1 = AcroMath<R>::cos_minus_half_pi(s[0]);
    acrobot_dsdt_trig<R>(s, a, s2, c2, cmh12, cmh1, out);
}

template <class R>
__device__ __forceinline__ R acro_wrap(R x, R m, R M) {
    const R diff = M - m;
    if (!(fabs(x) < R(1e6))) return x;   // non-finite / absurd input: leave as is (never loops forever)
    while (x > M) x = x - diff;
    while (x < m) x = x + diff;
    return x;
}

// sin/cos of the joint angles of one state: what the observation, the termination test and the first RK4
// stage of the NEXT step all need -- carried in registers by the float32 path
struct AcroTrig { float s1, c1, s2, c2, s12, c12; };

template <bool IN_RANGE = false>   // IN_RANGE: both angles known to be within the branch-free range of sincos
__device__ __forceinline__ AcroTrig acrobot_trig(const float v[4]) {
    AcroTrig t;
    sincosf_det<IN_RANGE>(v[0], &t.s1, &t.c1);
    sincosf_det<IN_RANGE>(v[1], &t.s2, &t.c2);
    // theta1 + theta2: angle-addition formulas (one fma + one multiply each) instead of a third sincos
    t.s12 = fmaf(t.s1, t.c2, t.c1 * t.s2);
    t.c12 = fmaf(t.c1, t.c2, -(t.s1 * t.s2));
    return t;
}

// Engine arithmetic of Acrobot (float32): the same book dynamics with the constants folded (m1 = m2 = l1 = 1,
// lc = 0.5, I = 1, g = 9.8: d1 = 3.5 + cos t2, d2 = 1.25 + cos t2 / 2, m2 lc2 g = 4.9, (m1 lc1 + m2 l1) g = 14.7),
// explicit fma, ONE reciprocal of d1 instead of three divisions by it, cos(x - pi/2) taken as sin x.  The two
// quotients have d1 in [2.5, 4.5] and 1.25 - d2^2/d1 in [0.56, 1.03]: div_inrange is the IEEE quotient for every
// state whose velocities are finite and below ~1e15 (every state the clamps of the previous step can produce).
__device__ __forceinline__ void acrobot_dsdt_f32(const float s[4], float a, float s2, float c2, float sh12, float sh1, float out[4]) {
    const float dth1 = s[2], dth2 = s[3];
    const float d1 = c2 + 3.5f;
    const float d2 = fmaf(0.5f, c2, 1.25f);
    const float phi2 = 4.9f * sh12;
    const float phi1 = fmaf(14.7f, sh1, phi2) - (s2 * dth2) * fmaf(0.5f, dth2, dth1);
    const float r1 = div_inrange(1.0f, d1);
    const float e = d2 * r1;
    const float num = (a - phi2) + fmaf(e, phi1, -((0.5f * s2) * (dth1 * dth1)));
    const float den = fmaf(-d2, e, 1.25f);
    const float ddth2 = div_inrange(num, den);
    const float ddth1 = -(fmaf(d2, ddth2, phi1) * r1);
    out[0] = dth1; out[1] = dth2; out[2] = ddth1; out[3] = ddth2;
}

// one classical RK4 step over dt = 0.2 in engine arithmetic, wrap and clamp; t0 = trig of the current state
template <bool IN_RANGE = false>
__device__ __forceinline__ void acrobot_rk4_f32(float s[4], int action, const AcroTrig& t0) {
    constexpr float PI = 3.14159265358979323846f;
    const float a = (float)(action - 1);
    float k1[4], k2[4], k3[4], k4[4], y[4];
    acrobot_dsdt_f32(s, a, t0.s2, t0.c2, t0.s12, t0.s1, k1);
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = fmaf(0.1f, k1[i], s[i]);
    AcroTrig t = acrobot_trig<IN_RANGE>(y);
    acrobot_dsdt_f32(y, a, t.s2, t.c2, t.s12, t.s1, k2);
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = fmaf(0.1f, k2[i], s[i]);
    t = acrobot_trig<IN_RANGE>(y);
    acrobot_dsdt_f32(y, a, t.s2, t.c2, t.s12, t.s1, k3);
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = fmaf(0.2f, k3[i], s[i]);
    t = acrobot_trig<IN_RANGE>(y);
    acrobot_dsdt_f32(y, a, t.s2, t.c2, t.s12, t.s1, k4);
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = fmaf(0.2f / 6.0f, fmaf(2.0f, k2[i] + k3[i], k1[i] + k4[i]), s[i]);
    y[0] = acro_wrap<float>(y[0], -PI, PI);
    y[1] = acro_wrap<float>(y[1], -PI, PI);
    constexpr float MV1 = 4 * PI, MV2 = 9 * PI;
    y[2] = y[2] < -MV1 ? -MV1 : (y[2] > MV1 ? MV1 : y[2]);
    y[3] = y[3] < -MV2 ? -MV2 : (y[3] > MV2 ? MV2 : y[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = y[i];
}

// one classical RK4 step of dsdt over dt = 0.2, wrap and clamp (upstream acrobot.py); s updated in place
template <class R, class Stage1>
__device__ __forceinline__ void acrobot_rk4(R s[4], int action, Stage1 stage1) {
    const R PI = R(3.14159265358979323846);
    const R dt = R(0.2);
    const R a = (R)(action - 1);
    R k1[4], k2[4], k3[4], k4[4], y[4];
    stage1(s, a, k1);
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = s[i] + dt / 2 * k1[i];
    acrobot_dsdt<R>(y, a, k2);
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = s[i] + dt / 2 * k2[i];
    acrobot_dsdt<R>(y, a, k3);
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = s[i] + dt * k3[i];
    acrobot_dsdt<R>(y, a, k4);
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = s[i] + dt / R(6.0) * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
    y[0] = acro_wrap<R>(y[0], -PI, PI);
    y[1] = acro_wrap<R>(y[1], -PI, PI);
    const R MV1 = 4 * PI, MV2 = 9 * PI;
    y[2] = y[2] < -MV1 ? -MV1 : (y[2] > MV1 ? MV1 : y[2]);
    y[3] = y[3] < -MV2 ? -MV2 : (y[3] > MV2 ? MV2 : y[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = y[i];
}

// double-precision evaluation of the termination value  -cos(th1) - cos(th2 + th1)  of the new state
__device__ __forceinline__ double acrobot_integrate_f64(double s[4], int action) {
    acrobot_rk4<double>(s, action, [](const double* st, double a, double* out) { acrobot_dsdt<double>(st, a, out); });
    return -cos(s[0]) - cos(s[1] + s[0]);
}

static __device__ __noinline__ unsigned acrobot_done_f64(float s0, float s1, float s2, float s3, int action) {
    double sd[4] = {(double)s0, (double)s1, (double)s2, (double)s3};
    return (unsigned)(acrobot_integrate_f64(sd, action) > 1.0);
}

// Engine arithmetic v3: beyond (9, 18) rad/s -- towards the corners of the velocity clamp box (4 pi, 9 pi), where one RK4
// step of 0.2 s changes the velocities by tens of rad/s and amplifies float32 rounding about a hundredfold (up to 2.5e-4
// relative, against the 1e-5 of north_star) -- the step is the upstream double-precision RK4, rounded to float32 once.
// Out of line and rare: random-policy episodes stay below (6, 12) rad/s, so rollouts never take it.
constexpr float ACROBOT_F32_MAX_V1 = 9.0f, ACROBOT_F32_MAX_V2 = 18.0f;
struct AcroFast { float v0, v1, v2, v3; unsigned done; };
static __device__ __noinline__ AcroFast acrobot_step_f64(float s0, float s1, float s2, float s3, int action) {
    double sd[4] = {(double)s0, (double)s1, (double)s2, (double)s3};
    const unsigned done = (unsigned)(acrobot_integrate_f64(sd, action) > 1.0);
    return AcroFast{(float)sd[0], (float)sd[1], (float)sd[2], (float)sd[3], done};
}

struct Acrobot {
    static constexpr int SD = 4, OD = 6, AD = 1, ACTN = 3, DEFAULT_LIMIT = 500;
    static constexpr bool HAS_SBD = false;
    static constexpr bool PREGEN_RESET = true;
    static constexpr int AUXW = 0;
    static constexpr bool REJECT_INVALID = true;
    static constexpr bool HAS_SMALL = true;         // step<true>: angles and velocities bounded, sincos without its cold path
    static constexpr bool ROLLOUT_CHUNK = false;    // one RK4 step is ~450 instructions: 8 unrolled copies run 1.4x slower (measured; instruction cache)
    using Vec = float4;
    using Act = int32_t;
    struct S { float v[4]; AcroTrig t; };
    __device__ static __forceinline__ S load(const void* base, const int32_t*, int, int i, const EnvParams&) {
        const float4 v = reinterpret_cast<const float4*>(base)[i];
        S s;
        s.v[0] = v.x; s.v[1] = v.y; s.v[2] = v.z; s.v[3] = v.w;
        s.t = acrobot_trig(s.v);
        return s;
    }
    __device__ static __forceinline__ void store(void* base, int32_t*, int, int i, const S& s) {
        reinterpret_cast<float4*>(base)[i] = make_float4(s.v[0], s.v[1], s.v[2], s.v[3]);
    }
    __device__ static __forceinline__ void reset(S& s, uint64_t seed, uint32_t gid, uint32_t ordinal, uint64_t, const EnvParams&) {
        const Block b = draw(seed, gid, (uint64_t)ordinal, STREAM_RESET);
        s.v[0] = uniformf(-0.1f, 0.1f, b.w0);
        s.v[1] = uniformf(-0.1f, 0.1f, b.w1);
        s.v[2] = uniformf(-0.1f, 0.1f, b.w2);
        s.v[3] = uniformf(-0.1f, 0.1f, b.w3);
        s.t = acrobot_trig(s.v);
    }
    __device__ static __forceinline__ bool valid(Act a) { return a >= 0 && a < 3; }
    // fast path of one step: the RK4 stages evaluate sincos at theta + {0.1, 0.1, 0.2} * velocity-like terms; with
    // |theta| <= 1000 and |dtheta| <= 1e4 every argument stays below 32768 (a step's own outputs are wrapped to
    // [-pi, pi] and clamped to 4 pi / 9 pi, so only SetState can leave this range)
    __device__ static __forceinline__ bool small_ok(const S& s) {
        return (fabsf(s.v[0]) <= 1000.0f) & (fabsf(s.v[1]) <= 1000.0f) & (fabsf(s.v[2]) <= 1.0e4f) & (fabsf(s.v[3]) <= 1.0e4f);
    }
    template <bool SMALL = false>
    __device__ static __forceinline__ StepOut step(S& s, Act a, int32_t&, uint64_t, uint32_t, uint64_t) {
        const float o0 = s.v[0], o1 = s.v[1], o2 = s.v[2], o3 = s.v[3];
        if (fabsf(o2) > ACROBOT_F32_MAX_V1 || fabsf(o3) > ACROBOT_F32_MAX_V2) {   // fast links: double-precision step (rare)
            const AcroFast f = acrobot_step_f64(o0, o1, o2, o3, (int)a);
            s.v[0] = f.v0; s.v[1] = f.v1; s.v[2] = f.v2; s.v[3] = f.v3;
            s.t = acrobot_trig<false>(s.v);
            return StepOut{f.done ? 0.0f : -1.0f, f.done};
        }
        const AcroTrig t0 = s.t;   // the first RK4 stage reuses the trig of the current state
        acrobot_rk4_f32<SMALL>(s.v, (int)a, t0);
        s.t = acrobot_trig<SMALL>(s.v);
        const float v = -s.t.c1 - s.t.c12;
        unsigned done = (unsigned)(v > 1.0f);
        if (fabsf(v - 1.0f) <= 2e-5f) done = acrobot_done_f64(o0, o1, o2, o3, (int)a);
        return StepOut{done ? 0.0f : -1.0f, done};
    }
    __device__ static __forceinline__ void obs(const S& s, float* o) {
        o[0] = s.t.c1; o[1] = s.t.s1; o[2] = s.t.c2; o[3] = s.t.s2; o[4] = s.v[2]; o[5] = s.v[3];
    }
};

}  // namespace gymcuda

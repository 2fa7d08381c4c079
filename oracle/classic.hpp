// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Scalar per-instance dynamics of the classic-control environments, in two arithmetics:
//   *_f64 : the reference's arithmetic (double; libm sin/cos; float32-rounded constants)
//   *_f32 : the engine's storage-precision arithmetic (detmath v1) -- bit-identical to the kernels
//
// CartPole follows the reference line by line:
//   src/Gym.Environments/Envs/Classic/CartPoleEnv.cs:24-36   constants (C# `const float`)
//   src/Gym.Environments/Envs/Classic/CartPoleEnv.cs:137-186 Step
//   src/Gym.Environments/Envs/Classic/CartPoleEnv.cs:63-67   Reset
// Pendulum / MountainCar / MountainCarContinuous / Acrobot do NOT exist in the reference
// (unchecked roadmap items, README.md:73-76).  Their spec is upstream openai/gym 0.26
// classic_control, restated from the published algorithm: PARITY UNPINNED BY THE REFERENCE.
#pragma once
#include <cmath>
#include <cstdint>
#include "detmath.hpp"

namespace oracle {

// ---------------------------------------------------------------- CartPole constants
// C# folds `const float` expressions in float32 (CartPoleEnv.cs:24-36); at use sites they are
// promoted to double.  These are the resulting float32 values.
namespace cp {
static const float GRAVITY = 9.8f;
static const float MASSPOLE = 0.1f;
static const float TOTAL_MASS = 0.1f + 1.0f;          // folded in float32 -> 1.10000002384185791015625
static const float LENGTH = 0.5f;
static const float POLEMASS_LENGTH = 0.1f * 0.5f;     // folded in float32
static const float FORCE_MAG = 10.0f;
static const float TAU = 0.02f;
static const float THETA_THRESHOLD = (float)(12 * 2 * 3.14159265358979323846 / 360);  // 0.20943951606750488
static const float X_THRESHOLD = 2.4f;
}

struct StepOut { float reward; uint8_t done; };

// CartPoleEnv.cs:137-186.  s = (x, x_dot, theta, theta_dot) in double (NDArray of doubles, :141-144).
inline StepOut cartpole_step_f64(double s[4], int action, int32_t* steps_beyond_done) {
    using namespace cp;
    double x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
    float force = action == 1 ? FORCE_MAG : -FORCE_MAG;                                   // :146
    double costheta = std::cos(theta);                                                    // :147
    double sintheta = std::sin(theta);                                                    // :148
    double temp = (force + POLEMASS_LENGTH * theta_dot * theta_dot * sintheta) / TOTAL_MASS;          // :149
    double thetaacc = (GRAVITY * sintheta - costheta * temp) /
                      (LENGTH * (4.0 / 3.0 - MASSPOLE * costheta * costheta / TOTAL_MASS));           // :150
    double xacc = temp - POLEMASS_LENGTH * thetaacc * costheta / TOTAL_MASS;                          // :151
    // kinematics_integrator == "euler" (:32, :153-157); the semi-implicit branch (:158-164) is dead code
    x = x + TAU * x_dot;
    x_dot = x_dot + TAU * xacc;
    theta = theta + TAU * theta_dot;
    theta_dot = theta_dot + TAU * thetaacc;
    s[0] = x; s[1] = x_dot; s[2] = theta; s[3] = theta_dot;                               // :166
    bool done = x < -X_THRESHOLD || x > X_THRESHOLD || theta < -THETA_THRESHOLD || theta > THETA_THRESHOLD;  // :167
    float reward;
    if (!done) {
        reward = 1.0f;                                                                    // :170
    } else if (*steps_beyond_done == -1) {
        *steps_beyond_done = 0;                                                           // :173
        reward = 1.0f;
    } else {
        *steps_beyond_done += 1;                                                          // :181
        reward = 0.0f;
    }
    return StepOut{reward, (uint8_t)done};
}

// Engine arithmetic (restated from DESIGN.md section 5, not from the reference): accelerations in fp32 with
// the divisions by total_mass folded into float32-rounded reciprocals and explicit fma.  Positions
// (:154,:156) are one float32 fma each = the reference's double sum rounded once to float32 (tau * x_dot
// is exact in double).  `done` (:167) is the reference's own double-precision test from the same
// float32 state, evaluated here unconditionally; the kernel decides it in float32 and only falls back
// to double when a new position equals a threshold exactly -- the same flag, since the thresholds are
// float32 values and rounding is monotonic.
inline StepOut cartpole_step_f32(float s[4], int action, int32_t* steps_beyond_done) {
    using namespace cp;
    const float INV_TOTAL_MASS = 0.9090908765792847f;   // fl32(1 / total_mass)
    const float K0 = 0.6666666865348816f;               // fl32(length * 4/3)
    const float K1 = 0.04545454680919647f;              // fl32(length * masspole / total_mass)
    const float PML_OVER_M = 0.04545454680919647f;      // fl32(polemass_length / total_mass)
    float x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
    float force = action == 1 ? FORCE_MAG : -FORCE_MAG;
    float sn, cs;
    det::sincosf_det(theta, &sn, &cs);
    float t1 = (POLEMASS_LENGTH * theta_dot) * theta_dot;
    float temp = std::fma(t1, sn, force) * INV_TOTAL_MASS;
    float den = std::fma(-K1, cs * cs, K0);
    float num = std::fma(GRAVITY, sn, -(cs * temp));
    float thetaacc = num / den;
    float xacc = std::fma(-(PML_OVER_M * thetaacc), cs, temp);
    double xd = (double)x + (double)TAU * (double)x_dot;
    double thd = (double)theta + (double)TAU * (double)theta_dot;
    s[1] = std::fma(TAU, xacc, x_dot);
    s[3] = std::fma(TAU, thetaacc, theta_dot);
    s[0] = std::fma(TAU, x_dot, x);
    s[2] = std::fma(TAU, theta_dot, theta);
    bool done = std::fabs(xd) > (double)X_THRESHOLD || std::fabs(thd) > (double)THETA_THRESHOLD;
    float reward = 1.0f;
    if (done) {
        if (*steps_beyond_done == -1) *steps_beyond_done = 0;
        else { *steps_beyond_done += 1; reward = 0.0f; }
    }
    return StepOut{reward, (uint8_t)done};
}

// ---------------------------------------------------------------- Pendulum-v1 (upstream spec)
namespace pd {
static const double G = 10.0, M = 1.0, L = 1.0, DT = 0.05, MAX_SPEED = 8.0, MAX_TORQUE = 2.0;
static const double PI = 3.14159265358979323846;
}

inline double py_mod(double a, double b) {   // Python float %, b > 0
    double m = std::fmod(a, b);
    if (m != 0.0) { if (m < 0.0) m += b; } else m = 0.0;
    return m;
}
inline float py_modf32(float a, float b) {
    float m = std::fmod(a, b);               // fmod is exact, hence deterministic
    if (m != 0.0f) { if (m < 0.0f) m += b; } else m = 0.0f;
    return m;
}

// upstream pendulum.py step(): u clipped, cost from the OLD state, thdot clipped before th update
inline StepOut pendulum_step_f64(double s[2], float action) {
    using namespace pd;
    double th = s[0], thdot = s[1];
    double u = (double)action;
    u = u < -MAX_TORQUE ? -MAX_TORQUE : (u > MAX_TORQUE ? MAX_TORQUE : u);
    double an = py_mod(th + PI, 2 * PI) - PI;
    double costs = an * an + 0.1 * (thdot * thdot) + 0.001 * (u * u);
    double newthdot = thdot + (3 * G / (2 * L) * std::sin(th) + 3.0 / (M * L * L) * u) * DT;
    newthdot = newthdot < -MAX_SPEED ? -MAX_SPEED : (newthdot > MAX_SPEED ? MAX_SPEED : newthdot);
    double newth = th + newthdot * DT;
    s[0] = newth; s[1] = newthdot;
    return StepOut{(float)(-costs), 0};
}

inline StepOut pendulum_step_f32(float s[2], float action) {
    const float PI_F = 3.1415927410125732f, TWO_PI_F = 6.2831854820251465f;
    float th = s[0], thdot = s[1];
    float u = det::clampf(action, -2.0f, 2.0f);
    float an = py_modf32(th + PI_F, TWO_PI_F) - PI_F;
    float costs = (an * an + 0.1f * (thdot * thdot)) + 0.001f * (u * u);
    float sn, cs;
    det::sincosf_det(th, &sn, &cs);
    float newthdot = thdot + (15.0f * sn + 3.0f * u) * 0.05f;
    newthdot = det::clampf(newthdot, -8.0f, 8.0f);
    float newth = th + newthdot * 0.05f;
    s[0] = newth; s[1] = newthdot;
    return StepOut{-costs, 0};
}

// ---------------------------------------------------------------- MountainCar-v0 / Continuous-v0
namespace mc {
static const double MIN_POS = -1.2, MAX_POS = 0.6, MAX_SPEED = 0.07;
static const double GOAL_DISCRETE = 0.5, GOAL_CONT = 0.45;
static const double FORCE = 0.001, GRAVITY = 0.0025, POWER = 0.0015;
}

inline StepOut mountaincar_step_f64(double s[2], int action) {
    using namespace mc;
    double position = s[0], velocity = s[1];
    velocity += (action - 1) * FORCE + std::cos(3 * position) * (-GRAVITY);
    velocity = velocity < -MAX_SPEED ? -MAX_SPEED : (velocity > MAX_SPEED ? MAX_SPEED : velocity);
    position += velocity;
    position = position < MIN_POS ? MIN_POS : (position > MAX_POS ? MAX_POS : position);
    if (position == MIN_POS && velocity < 0) velocity = 0;
    bool done = position >= GOAL_DISCRETE && velocity >= 0.0;
    s[0] = position; s[1] = velocity;
    return StepOut{-1.0f, (uint8_t)done};
}

inline StepOut mountaincar_cont_step_f64(double s[2], float action) {
    using namespace mc;
    double position = s[0], velocity = s[1];
    double a = (double)action;
    double force = a < -1.0 ? -1.0 : (a > 1.0 ? 1.0 : a);
    velocity += force * POWER - 0.0025 * std::cos(3 * position);
    if (velocity > MAX_SPEED) velocity = MAX_SPEED;
    if (velocity < -MAX_SPEED) velocity = -MAX_SPEED;
    position += velocity;
    if (position > MAX_POS) position = MAX_POS;
    if (position < MIN_POS) position = MIN_POS;
    if (position == MIN_POS && velocity < 0) velocity = 0;
    bool done = position >= GOAL_CONT && velocity >= 0.0;
    double reward = 0;
    if (done) reward = 100.0;
    reward -= (a * a) * 0.1;
    s[0] = position; s[1] = velocity;
    return StepOut{(float)reward, (uint8_t)done};
}

// fp32 path with double refinement of `done` when the fp32 values are within rounding distance
// of a threshold (so the flag equals the double evaluation from the same float32 state).
inline StepOut mountaincar_any_step_f32(float s[2], bool continuous, int iaction, float faction) {
    const float MIN_POS = -1.2f, MAX_POS = 0.6f, MAX_SPEED = 0.07f;
    const float goal = continuous ? 0.45f : 0.5f;
    float position = s[0], velocity = s[1];
    float sn, cs;
    det::sincosf_det(3.0f * position, &sn, &cs);
    float push;
    float fclip = 0.0f;
    if (continuous) { fclip = det::clampf(faction, -1.0f, 1.0f); push = fclip * 0.0015f; }
    else push = (float)(iaction - 1) * 0.001f;
    float nv = velocity + (push + cs * (-0.0025f));
    nv = det::clampf(nv, -MAX_SPEED, MAX_SPEED);
    float np = position + nv;
    np = det::clampf(np, MIN_POS, MAX_POS);
    if (np == MIN_POS && nv < 0.0f) nv = 0.0f;
    bool done = false;
    if (np >= goal - 1e-6f) {   // below: not done in float32, and the double evaluation (< 1e-7 away) agrees
        done = np >= goal && nv >= 0.0f;
        if (std::fabs(np - goal) <= 1e-6f || std::fabs(nv) <= 1e-7f) {
            double sd[2] = {(double)position, (double)velocity};
            StepOut r = continuous ? mountaincar_cont_step_f64(sd, faction) : mountaincar_step_f64(sd, iaction);
            done = r.done != 0;
        }
    }
    s[0] = np; s[1] = nv;
    float reward;
    if (continuous) {
        reward = done ? 100.0f : 0.0f;
        reward = reward - (faction * faction) * 0.1f;
    } else reward = -1.0f;
    return StepOut{reward, (uint8_t)done};
}

// ---------------------------------------------------------------- Acrobot-v1 (book dynamics, RK4, dt 0.2)
template <class R> struct AcroMath;
template <> struct AcroMath<double> {
    static void sincos(double x, double* s, double* c) { *s = std::sin(x); *c = std::cos(x); }
    // upstream evaluates cos(theta - pi/2) literally
    static double cos_minus_half_pi(double x) { return std::cos(x - 3.14159265358979323846 / 2.0); }
};

template <class R>
inline void acrobot_dsdt(const R s[4], R a, R out[4]) {
    const R m1 = 1, m2 = 1, l1 = 1, lc1 = R(0.5), lc2 = R(0.5), I1 = 1, I2 = 1, g = R(9.8);
    R theta1 = s[0], theta2 = s[1], dtheta1 = s[2], dtheta2 = s[3];
    R s2, c2;
    AcroMath<R>::sincos(theta2, &s2, &c2);
    R d1 = m1 * lc1 * lc1 + m2 * (l1 * l1 + lc2 * lc2 + 2 * l1 * lc2 * c2) + I1 + I2;
    R d2 = m2 * (lc2 * lc2 + l1 * lc2 * c2) + I2;
    R phi2 = m2 * lc2 * g * AcroMath<R>::cos_minus_half_pi(theta1 + theta2);
    R phi1 = -m2 * l1 * lc2 * dtheta2 * dtheta2 * s2 - 2 * m2 * l1 * lc2 * dtheta2 * dtheta1 * s2 +
             (m1 * lc1 + m2 * l1) * g * AcroMath<R>::cos_minus_half_pi(theta1) + phi2;
    R ddtheta2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * dtheta1 * dtheta1 * s2 - phi2) /
                 (m2 * lc2 * lc2 + I2 - d2 * d2 / d1);
    R ddtheta1 = -(d2 * ddtheta2 + phi1) / d1;
    out[0] = dtheta1; out[1] = dtheta2; out[2] = ddtheta1; out[3] = ddtheta2;
}

template <class R>
inline R acro_wrap(R x, R m, R M) {
    R diff = M - m;
    if (!(std::fabs(x) < R(1e6))) return x;   // non-finite / absurd input: leave as is (never loops forever)
    while (x > M) x = x - diff;
    while (x < m) x = x + diff;
    return x;
}

// returns the termination value  -cos(th1) - cos(th2 + th1)  of the NEW state
template <class R>
inline R acrobot_integrate(R s[4], int action) {
    const R PI = R(3.14159265358979323846);
    const R dt = R(0.2);
    R a = (R)(action - 1);
    R k1[4], k2[4], k3[4], k4[4], y[4];
    acrobot_dsdt<R>(s, a, k1);
    for (int i = 0; i < 4; ++i) y[i] = s[i] + dt / 2 * k1[i];
    acrobot_dsdt<R>(y, a, k2);
    for (int i = 0; i < 4; ++i) y[i] = s[i] + dt / 2 * k2[i];
    acrobot_dsdt<R>(y, a, k3);
    for (int i = 0; i < 4; ++i) y[i] = s[i] + dt * k3[i];
    acrobot_dsdt<R>(y, a, k4);
    for (int i = 0; i < 4; ++i) y[i] = s[i] + dt / R(6.0) * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
    y[0] = acro_wrap<R>(y[0], -PI, PI);
    y[1] = acro_wrap<R>(y[1], -PI, PI);
    const R MV1 = 4 * PI, MV2 = 9 * PI;
    y[2] = y[2] < -MV1 ? -MV1 : (y[2] > MV1 ? MV1 : y[2]);
    y[3] = y[3] < -MV2 ? -MV2 : (y[3] > MV2 ? MV2 : y[3]);
    for (int i = 0; i < 4; ++i) s[i] = y[i];
    R s1, c1, s12, c12;
    AcroMath<R>::sincos(y[0], &s1, &c1);
    AcroMath<R>::sincos(y[1] + y[0], &s12, &c12);
    return -c1 - c12;
}

inline StepOut acrobot_step_f64(double s[4], int action) {
    double v = acrobot_integrate<double>(s, action);
    bool done = v > 1.0;
    return StepOut{done ? 0.0f : -1.0f, (uint8_t)done};
}

// Engine arithmetic of Acrobot (float32; restated from DESIGN.md, not from upstream): the same book dynamics
// with the constants folded (m1 = m2 = l1 = 1, lc = 0.5, I = 1, g = 9.8: d1 = 3.5 + cos t2, d2 = 1.25 + cos t2 / 2,
// m2 lc2 g = 4.9, (m1 lc1 + m2 l1) g = 14.7), explicit fma, ONE reciprocal of d1 instead of three divisions by it,
// cos(x - pi/2) taken as sin x, and sin/cos(theta1 + theta2) from the angle-addition formulas.
// s2, c2 = sin/cos(theta2), sh12 = sin(theta1 + theta2), sh1 = sin(theta1).
inline void acrobot_dsdt_f32(const float s[4], float a, float s2, float c2, float sh12, float sh1, float out[4]) {
    const float dth1 = s[2], dth2 = s[3];
    const float d1 = c2 + 3.5f;
    const float d2 = std::fma(0.5f, c2, 1.25f);
    const float phi2 = 4.9f * sh12;
    const float phi1 = std::fma(14.7f, sh1, phi2) - (s2 * dth2) * std::fma(0.5f, dth2, dth1);
    const float r1 = 1.0f / d1;
    const float e = d2 * r1;
    const float num = (a - phi2) + std::fma(e, phi1, -((0.5f * s2) * (dth1 * dth1)));
    const float den = std::fma(-d2, e, 1.25f);
    const float ddth2 = num / den;
    const float ddth1 = -(std::fma(d2, ddth2, phi1) * r1);
    out[0] = dth1; out[1] = dth2; out[2] = ddth1; out[3] = ddth2;
}

// sin / cos of theta1 + theta2 from the two angles' own sin / cos (angle-addition formulas, one fma + one
// multiply each) instead of a third sincos evaluation
inline void acrobot_trig_f32(const float y[4], float* s1, float* c1, float* s2, float* c2, float* s12, float* c12) {
    det::sincosf_det(y[0], s1, c1);
    det::sincosf_det(y[1], s2, c2);
    *s12 = std::fma(*s1, *c2, *c1 * *s2);
    *c12 = std::fma(*c1, *c2, -(*s1 * *s2));
}

inline void acrobot_dsdt_f32(const float y[4], float a, float out[4]) {
    float s2, c2, s12, c12, s1, c1;
    acrobot_trig_f32(y, &s1, &c1, &s2, &c2, &s12, &c12);
    acrobot_dsdt_f32(y, a, s2, c2, s12, s1, out);
}

// one classical RK4 step over dt = 0.2 in engine arithmetic, wrap and clamp; returns -cos(th1) - cos(th2 + th1)
inline float acrobot_integrate_f32(float s[4], int action) {
    const float PI = 3.14159265358979323846f;
    const float a = (float)(action - 1);
    float k1[4], k2[4], k3[4], k4[4], y[4];
    acrobot_dsdt_f32(s, a, k1);
    for (int i = 0; i < 4; ++i) y[i] = std::fma(0.1f, k1[i], s[i]);
    acrobot_dsdt_f32(y, a, k2);
    for (int i = 0; i < 4; ++i) y[i] = std::fma(0.1f, k2[i], s[i]);
    acrobot_dsdt_f32(y, a, k3);
    for (int i = 0; i < 4; ++i) y[i] = std::fma(0.2f, k3[i], s[i]);
    acrobot_dsdt_f32(y, a, k4);
    for (int i = 0; i < 4; ++i) y[i] = std::fma(0.2f / 6.0f, std::fma(2.0f, k2[i] + k3[i], k1[i] + k4[i]), s[i]);
    y[0] = acro_wrap<float>(y[0], -PI, PI);
    y[1] = acro_wrap<float>(y[1], -PI, PI);
    const float MV1 = 4 * PI, MV2 = 9 * PI;
    y[2] = y[2] < -MV1 ? -MV1 : (y[2] > MV1 ? MV1 : y[2]);
    y[3] = y[3] < -MV2 ? -MV2 : (y[3] > MV2 ? MV2 : y[3]);
    for (int i = 0; i < 4; ++i) s[i] = y[i];
    float s1, c1, s2, c2, s12, c12;
    acrobot_trig_f32(y, &s1, &c1, &s2, &c2, &s12, &c12);
    return -c1 - c12;
}

// Engine arithmetic v3: a state whose link velocities are beyond (9, 18) rad/s -- towards the corners of the clamp box
// (4 pi, 9 pi), where one RK4 step of 0.2 s amplifies float32 rounding about a hundredfold -- is stepped with the
// upstream double-precision RK4 and rounded to float32 once.
static const float ACROBOT_F32_MAX_V1 = 9.0f, ACROBOT_F32_MAX_V2 = 18.0f;

inline StepOut acrobot_step_f32(float s[4], int action) {
    double sd[4] = {s[0], s[1], s[2], s[3]};
    if (std::fabs(s[2]) > ACROBOT_F32_MAX_V1 || std::fabs(s[3]) > ACROBOT_F32_MAX_V2) {
        const bool done = acrobot_integrate<double>(sd, action) > 1.0;
        for (int i = 0; i < 4; ++i) s[i] = (float)sd[i];
        return StepOut{done ? 0.0f : -1.0f, (uint8_t)done};
    }
    float v = acrobot_integrate_f32(s, action);
    bool done = v > 1.0f;
    if (std::fabs(v - 1.0f) <= 2e-5f) done = acrobot_integrate<double>(sd, action) > 1.0;
    return StepOut{done ? 0.0f : -1.0f, (uint8_t)done};
}

}  // namespace oracle

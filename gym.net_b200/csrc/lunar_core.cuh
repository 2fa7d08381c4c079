// LunarLander on the GPU: one thread owns one lander (3 rigid bodies, 2 revolute joints, 11 static
// terrain edges) and runs the whole World.Step of the reference for it.
//
// Restates Gym.Environments.Envs.Aether.LunarLanderEnv:
//   src/Gym.Environments/Envs/Aether/LunarLanderEnv.cs:155-164   constants
//   :181-302  LunarLanderBody (fuselage polygon, corner-anchored leg boxes, revolute joints)
//   :305-346  ContactDetector ("last BeginContact wins" flags)
//   :489-572  Reset (terrain, initial force, zero step)
//   :574-774  Step (engines, World.Step(1/50, 180 velocity / 60 position iterations), obs, reward, done)
// and, because the reference delegates all rigid-body arithmetic to the un-vendored NuGet package
// Aether.Physics2D 1.6.1 (src/Gym.Environments/Gym.Environments.csproj:33), the published algorithm of
// its Box2D-2.3 / Farseer-3.5 lineage specialised to this fixed topology: edge-vs-polygon manifolds
// (b2CollideEdgeAndPolygon), sequential-impulse contact solver with block solver and warm starting,
// revolute joint with motor + limit, island integration and sleeping.  Deviations are listed in
// DESIGN.md ("LunarLander"): no TOI sub-stepping, inert particles skipped, fixed constraint order,
// at most MAXC touching manifolds, float32 sin/cos (detmath).
//
// This kernel is bound by dependent float32 arithmetic (180 + 60 sequential solver iterations, about
// 1e5 flop against ~0.8 KB of state per env step), not by HBM.
//
// HBM layout: field-major structure of arrays, word f of env i at state[f * n + i] (float) and
// aux[f * n + i] (int32), so every load/store of the 32 lanes of a warp is one coalesced 128 B line.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

#include "detmath.cuh"
#include "philox.cuh"

namespace gymcuda { namespace lunar {

// ---------------------------------------------------------------- constants
constexpr float SCALE = 30.0f;                       // :156
constexpr float MAIN_ENGINE_POWER = 13.0f;           // :157
constexpr float SIDE_ENGINE_POWER = 0.6f;            // :158
constexpr float INITIAL_RANDOM = 1000.0f;            // :159
constexpr float SIDE_ENGINE_HEIGHT = 14.0f;          // :160
constexpr float SIDE_ENGINE_AWAY = 12.0f;            // :161
constexpr float VIEW_W = 600.0f / 30.0f;             // VIEWPORT_W / SCALE
constexpr float VIEW_H = 400.0f / 30.0f;             // VIEWPORT_H / SCALE
constexpr float DT = 1.0f / 50.0f;                   // :721 (FPS = 50)
constexpr float LEG_AWAY = 20.0f, LEG_DOWN = 18.0f;  // :186-187
constexpr int CHUNKS = 11;                           // :503
constexpr int NUM_EDGES = 11;                        // 10 terrain edges (:545-557) + the base edge (:541)
constexpr int BASE_EDGE = 10;

// Box2D-2.3 / Farseer-3.5 lineage settings
constexpr float LINEAR_SLOP = 0.005f;
constexpr float LINEAR_SLOP_SQ_MAX = 2.5000001187436283e-05f;   // largest float32 x with sqrtf(x) <= LINEAR_SLOP (joint_position)
constexpr float ANGULAR_SLOP = 0.03490658849477768f;          // (2.0f / 180.0f * Pi) folded in float32 like the C# / C++ constant expression (NOT the double product rounded once: 0.034906584769…)
constexpr float POLYGON_RADIUS = 0.01f;                      // 2 * linearSlop
constexpr float BAUMGARTE = 0.2f;
constexpr float MAX_LINEAR_CORRECTION = 0.2f;
constexpr float MAX_ANGULAR_CORRECTION = 0.13962635397911072f;  // (8.0f / 180.0f * Pi), float32 folding
constexpr float MAX_TRANSLATION = 2.0f;
constexpr float MAX_ROTATION = 1.5707963705062866f;          // 0.5*pi
constexpr float VELOCITY_THRESHOLD = 1.0f;
constexpr float TIME_TO_SLEEP = 0.5f;
constexpr float LINEAR_SLEEP_TOL = 0.01f;
constexpr float ANGULAR_SLEEP_TOL = 0.03490658849477768f;
// (the two macros exist for timing probes only -- tools/lunar_split_probe.py builds variants with 1 iteration to see where a
// step's time goes; the product is always built with the reference's counts)
#ifndef LUNAR_VEL_ITERS
#define LUNAR_VEL_ITERS 180
#endif
#ifndef LUNAR_POS_ITERS
#define LUNAR_POS_ITERS 60
#endif
constexpr int VELOCITY_ITERATIONS = LUNAR_VEL_ITERS, POSITION_ITERATIONS = LUNAR_POS_ITERS;   // :723-724
constexpr float DEFAULT_FRICTION = 0.2f;

constexpr int MAXC = 6;   // touching manifolds kept per lander
constexpr int MAXP = 12;  // broad-phase pairs (contacts that exist, touching or not) kept per lander, in creation order
constexpr float AABB_EXTENSION = 0.1f, AABB_MULTIPLIER = 2.0f;   // Settings.AABBExtension / AABBMultiplier

// ---------------------------------------------------------------- small vector algebra (b2Math)
struct V2 { float x, y; };
__device__ __forceinline__ V2 mk(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ V2 operator+(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ V2 operator-(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ V2 operator-(V2 a) { return mk(-a.x, -a.y); }
__device__ __forceinline__ V2 operator*(float s, V2 a) { return mk(s * a.x, s * a.y); }
__device__ __forceinline__ float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ V2 cross_sv(float s, V2 a) { return mk(-s * a.y, s * a.x); }
__device__ __forceinline__ V2 cross_vs(V2 a, float s) { return mk(s * a.y, -s * a.x); }
struct Rot { float s, c; };
__device__ __forceinline__ Rot rot(float a) { Rot q; sincosf_det(a, &q.s, &q.c); return q; }
__device__ __forceinline__ V2 rmul(Rot q, V2 v) { return mk(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
__device__ __forceinline__ V2 rmulT(Rot q, V2 v) { return mk(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ float minf(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float maxf(float a, float b) { return a > b ? a : b; }

// ---------------------------------------------------------------- shapes (:189, :262-266), mass data
// b2PolygonShape::Set order (gift-wrapped hull, CCW from the right-most lowest vertex) and
// b2PolygonShape::ComputeMass evaluated once in float32 (tests/golden/lunar_mass_data.py re-derives the numbers).
struct Shape { int count; V2 v[6]; V2 n[6]; V2 centroid; float mass, inv_mass, inertia, inv_inertia, friction; };
__constant__ Shape SHAPES[3] = {
    {6,
     {{0.5666666626930237f, -0.3333333432674408f}, {0.5666666626930237f, 0.0f}, {0.46666666865348816f, 0.5666666626930237f},
      {-0.46666666865348816f, 0.5666666626930237f}, {-0.5666666626930237f, 0.0f}, {-0.5666666626930237f, -0.3333333432674408f}},
     {{1.0f, 0.0f}, {0.9847835898399353f, 0.17378532886505127f}, {0.0f, 1.0f}, {-0.9847835898399353f, 0.17378532886505127f},
      {-1.0f, 0.0f}, {0.0f, -1.0f}},
     {0.0f, 0.10130719095468521f}, 4.816666603088379f, 0.20761245489120483f, 0.7838807106018066f, 1.275704264640808f, 0.1f},
    {4,
     {{0.06666667014360428f, 0.0f}, {0.06666667014360428f, 0.2666666805744171f}, {0.0f, 0.2666666805744171f}, {0.0f, 0.0f}, {0, 0}, {0, 0}},
     {{1.0f, 0.0f}, {0.0f, 1.0f}, {-1.0f, 0.0f}, {0.0f, -1.0f}, {0, 0}, {0, 0}},
     {0.03333333507180214f, 0.13333334028720856f}, 0.017777780070900917f, 56.24999237060547f, 0.00011193420505151153f, 8933.8193359375f, 0.2f},
    {4,
     {{0.06666667014360428f, 0.0f}, {0.06666667014360428f, 0.2666666805744171f}, {0.0f, 0.2666666805744171f}, {0.0f, 0.0f}, {0, 0}, {0, 0}},
     {{1.0f, 0.0f}, {0.0f, 1.0f}, {-1.0f, 0.0f}, {0.0f, -1.0f}, {0, 0}, {0, 0}},
     {0.03333333507180214f, 0.13333334028720856f}, 0.017777780070900917f, 56.24999237060547f, 0.00011193420505151153f, 8933.8193359375f, 0.2f},
};

// revolute joints fuselage <-> leg (:271-277): anchorA = (0,0), anchorB = (+-LEG_AWAY/S, LEG_DOWN/S),
// referenceAngle = leg.Rotation - fuselage.Rotation at creation = -+0.05 (:259)
struct JointDef { V2 anchor_b; float ref_angle, motor_speed, lower, upper; };
__constant__ JointDef JOINTS[2] = {
    {{-0.6666666865348816f, 0.6000000238418579f}, -0.05f, -0.3f, 0.3999999761581421f, 0.9f},      // leg 0: (0.9f - 0.5f, 0.9f) evaluated in float32 like the C# (:276-277): NOT 0.4f
    {{0.6666666865348816f, 0.6000000238418579f}, 0.05f, 0.3f, -0.9f, -0.3999999761581421f},       // leg 1: (-0.9f, -0.9f + 0.5f)
};
constexpr float MAX_MOTOR_TORQUE = 40.0f;   // LEG_SPRING_TORQUE (:188, :274)
enum { LIMIT_INACTIVE = 0, LIMIT_AT_LOWER = 1, LIMIT_AT_UPPER = 2, LIMIT_EQUAL = 3 };

// ---------------------------------------------------------------- persistent state of one lander
struct Body { V2 c; float a; V2 v; float w; float sleep_time; };
struct Joint { float ix, iy, iz, motor; int32_t limit_state; };
struct ContactSlot { int32_t pair; uint32_t key[2]; float ni[2], ti[2]; };   // pair = body * 16 + edge, -1 = free
constexpr uint32_t NO_KEY = 0xffffffffu;

enum : int32_t {
    F_GAME_OVER = 1, F_LEG0 = 2, F_LEG1 = 4, F_FUSELAGE = 8, F_AWAKE = 16, F_FIRST_STEP = 32, F_CONTINUOUS = 64
};

struct Lander {
    Body b[3];
    Joint j[2];
    ContactSlot c[MAXC];
    float terrain[CHUNKS];     // smooth_y (:523-536)
    float prev_shaping;        // float.MinValue sentinel after Reset (:498)
    V2 force;                  // force accumulator of the fuselage (ApplyForce in Reset, :496)
    float torque;
    uint32_t touch[3];         // bit e of touch[body]: polygon `body` was touching edge e after the last Collide
    float fat[3][4];           // broad-phase proxy box of each polygon (lo.x, lo.y, hi.x, hi.y): replaced only when the swept tight box leaves it
    uint32_t pairs[3];         // the contacts that exist (fat boxes overlap), in creation order: byte k = body * 16 + edge, 0xff = none
    int32_t flags;
    int32_t wind_idx, torque_idx;
    float gravity, wind_power, turbulence_power;
    int32_t use_wind;
    float obs[8];
};

constexpr int STATE_DIM = 21 + 8 + 4 * MAXC + CHUNKS + 1 + 3 + 12;   // 80
constexpr int AUX_DIM = 3 + 1 + 2 + 3 * MAXC + 2 + 3 + 2;              // touch[3], flags, limit[2], slots, wind idx, pairs[3], ep_t, episode

__device__ __forceinline__ void zero_lander(Lander& L) {
    for (int i = 0; i < 3; ++i) { L.b[i].c = mk(0.0f, 0.0f); L.b[i].a = 0.0f; L.b[i].v = mk(0.0f, 0.0f); L.b[i].w = 0.0f; L.b[i].sleep_time = 0.0f; L.touch[i] = 0u; }
    for (int i = 0; i < 2; ++i) { L.j[i].ix = L.j[i].iy = L.j[i].iz = L.j[i].motor = 0.0f; L.j[i].limit_state = 0; }
    for (int s = 0; s < MAXC; ++s) { L.c[s].pair = 0; L.c[s].key[0] = L.c[s].key[1] = 0u; L.c[s].ni[0] = L.c[s].ni[1] = L.c[s].ti[0] = L.c[s].ti[1] = 0.0f; }
    for (int i = 0; i < CHUNKS; ++i) L.terrain[i] = 0.0f;
    L.prev_shaping = 0.0f; L.force = mk(0.0f, 0.0f); L.torque = 0.0f; L.flags = 0;
    L.wind_idx = 0; L.torque_idx = 0; L.gravity = 0.0f; L.wind_power = 0.0f; L.turbulence_power = 0.0f; L.use_wind = 0;
    for (int i = 0; i < 8; ++i) L.obs[i] = 0.0f;
    for (int i = 0; i < 3; ++i) { L.pairs[i] = 0xffffffffu; for (int j = 0; j < 4; ++j) L.fat[i][j] = 0.0f; }
}

// ---------------------------------------------------------------- geometry helpers
__device__ __forceinline__ void edge_points(const Lander& L, int e, V2* v1, V2* v2) {
    if (e == BASE_EDGE) { *v1 = mk(0.0f, 0.0f); *v2 = mk(VIEW_W, 0.0f); return; }        // :541
    const float cw = VIEW_W / (float)(CHUNKS - 1);                                        // :512
    *v1 = mk(cw * (float)e, L.terrain[e]);                                                // :547-548
    *v2 = mk(cw * (float)(e + 1), L.terrain[e + 1]);
}
__device__ __forceinline__ float edge_friction(int e) { return e == BASE_EDGE ? DEFAULT_FRICTION : 0.1f; }   // :550

// ---------------------------------------------------------------- narrow phase: b2CollideEdgeAndPolygon
struct ClipVertex { V2 v; uint32_t key; };
__device__ __forceinline__ uint32_t make_key(int ia, int ib, int ta, int tb) { return (uint32_t)ia | ((uint32_t)ib << 8) | ((uint32_t)ta << 16) | ((uint32_t)tb << 24); }
enum { FEAT_VERTEX = 0, FEAT_FACE = 1 };
enum { MF_FACE_A = 1, MF_FACE_B = 2 };

struct Manifold {
    int type, count;
    V2 local_normal, local_point;   // faceA: in the edge (world) frame; faceB: in the polygon's local frame
    V2 lp[2];                       // faceA: polygon-local; faceB: world
    uint32_t key[2];
};

__device__ __forceinline__ int clip_segment(ClipVertex out[2], const ClipVertex in[2], V2 normal, float offset, int vertex_index_a) {
    int n = 0;
    const float d0 = dot(normal, in[0].v) - offset;
    const float d1 = dot(normal, in[1].v) - offset;
    if (d0 <= 0.0f) out[n++] = in[0];
    if (d1 <= 0.0f) out[n++] = in[1];
    if (d0 * d1 < 0.0f) {
        const float interp = d0 / (d0 - d1);
        out[n].v = in[0].v + interp * (in[1].v - in[0].v);
        out[n].key = make_key(vertex_index_a, (int)((in[0].key >> 8) & 0xff), FEAT_VERTEX, FEAT_FACE);
        ++n;
    }
    return n;
}

// The first exit of b2CollideEdgeAndPolygon -- ComputeEdgeSeparation: the polygon's deepest vertex is further than the
// contact radius from the edge's line -- evaluated INLINE with exactly the operations of collide_edge_polygon below (same
// bits: every operation is a single IEEE operation, nothing is contracted).  A lander near the ground holds six to nine
// broad-phase pairs (fat boxes overlap) of which one or two touch: only those pay the out-of-line narrow phase, with its
// vertex arrays in local memory (round 2, per-warp clocks: Collide was 108 k of a contact warp's 840 k cycles).
__device__ __forceinline__ bool edge_polygon_apart(V2 v1, V2 v2, int body, V2 p, Rot q) {
    const V2 centroid = rmul(q, SHAPES[body].centroid) + p;
    V2 edge1 = v2 - v1;
    {
        const float len = sqrtf(edge1.x * edge1.x + edge1.y * edge1.y);
        const float inv = 1.0f / len;
        edge1 = mk(edge1.x * inv, edge1.y * inv);
    }
    const V2 normal1 = mk(edge1.y, -edge1.x);
    const float offset1 = dot(normal1, centroid - v1);
    const V2 normal = offset1 >= 0.0f ? normal1 : -normal1;
    float edge_sep = 3.4028234663852886e38f;
    const int count = SHAPES[body].count;
#pragma unroll
    for (int i = 0; i < 6; ++i)
        if (i < count) { const float s = dot(normal, (rmul(q, SHAPES[body].v[i]) + p) - v1); if (s < edge_sep) edge_sep = s; }
    return edge_sep > 2.0f * POLYGON_RADIUS;
}

// Edge A (static, identity transform, no adjacent vertices) vs polygon B with transform (p, q).
static __device__ __noinline__ void collide_edge_polygon(Manifold* m, V2 v1, V2 v2, const Shape& sh, V2 p, Rot q) {
    m->count = 0;
    m->type = 0;
    const V2 centroid = rmul(q, sh.centroid) + p;
    V2 edge1 = v2 - v1;
    {   // b2Vec2::Normalize
        const float len = sqrtf(edge1.x * edge1.x + edge1.y * edge1.y);
        const float inv = 1.0f / len;
        edge1 = mk(edge1.x * inv, edge1.y * inv);
    }
    const V2 normal1 = mk(edge1.y, -edge1.x);
    const float offset1 = dot(normal1, centroid - v1);
    const bool front = offset1 >= 0.0f;
    V2 normal, lower, upper;
    if (front) { normal = normal1; lower = -normal1; upper = -normal1; }
    else { normal = -normal1; lower = normal1; upper = normal1; }
    V2 vb[6], nb[6];
    for (int i = 0; i < sh.count; ++i) { vb[i] = rmul(q, sh.v[i]) + p; nb[i] = rmul(q, sh.n[i]); }
    const float radius = 2.0f * POLYGON_RADIUS;

    // ComputeEdgeSeparation
    float edge_sep = 3.4028234663852886e38f;
    for (int i = 0; i < sh.count; ++i) { const float s = dot(normal, vb[i] - v1); if (s < edge_sep) edge_sep = s; }
    if (edge_sep > radius) return;

    // ComputePolygonSeparation
    int poly_index = -1;
    float poly_sep = -3.4028234663852886e38f;
    bool poly_valid = false;
    {
        const V2 perp = mk(-normal.y, normal.x);
        for (int i = 0; i < sh.count; ++i) {
            const V2 n = -nb[i];
            const float s1 = dot(n, vb[i] - v1);
            const float s2 = dot(n, vb[i] - v2);
            const float s = minf(s1, s2);
            if (s > radius) { poly_valid = true; poly_index = i; poly_sep = s; break; }   // no collision
            if (dot(n, perp) >= 0.0f) { if (dot(n - upper, normal) < -ANGULAR_SLOP) continue; }
            else { if (dot(n - lower, normal) < -ANGULAR_SLOP) continue; }
            if (s > poly_sep) { poly_valid = true; poly_index = i; poly_sep = s; }
        }
    }
    if (poly_valid && poly_sep > radius) return;

    const float k_relative_tol = 0.98f, k_absolute_tol = 0.001f;
    const bool primary_is_edge = !poly_valid || !(poly_sep > k_relative_tol * edge_sep + k_absolute_tol);

    ClipVertex ie[2];
    int rf_i1, rf_i2;
    V2 rf_v1, rf_v2, rf_normal;
    if (primary_is_edge) {
        m->type = MF_FACE_A;
        int best = 0;
        float best_value = dot(normal, nb[0]);
        for (int i = 1; i < sh.count; ++i) { const float v = dot(normal, nb[i]); if (v < best_value) { best_value = v; best = i; } }
        const int i1 = best, i2 = i1 + 1 < sh.count ? i1 + 1 : 0;
        ie[0].v = vb[i1]; ie[0].key = make_key(0, i1, FEAT_FACE, FEAT_VERTEX);
        ie[1].v = vb[i2]; ie[1].key = make_key(0, i2, FEAT_FACE, FEAT_VERTEX);
        if (front) { rf_i1 = 0; rf_i2 = 1; rf_v1 = v1; rf_v2 = v2; rf_normal = normal1; }
        else { rf_i1 = 1; rf_i2 = 0; rf_v1 = v2; rf_v2 = v1; rf_normal = -normal1; }
    } else {
        m->type = MF_FACE_B;
        ie[0].v = v1; ie[0].key = make_key(0, poly_index, FEAT_VERTEX, FEAT_FACE);
        ie[1].v = v2; ie[1].key = make_key(0, poly_index, FEAT_VERTEX, FEAT_FACE);
        rf_i1 = poly_index; rf_i2 = rf_i1 + 1 < sh.count ? rf_i1 + 1 : 0;
        rf_v1 = vb[rf_i1]; rf_v2 = vb[rf_i2]; rf_normal = nb[rf_i1];
    }
    const V2 side1 = mk(rf_normal.y, -rf_normal.x);
    const V2 side2 = -side1;
    const float side_off1 = dot(side1, rf_v1);
    const float side_off2 = dot(side2, rf_v2);
    ClipVertex c1[2], c2[2];
    if (clip_segment(c1, ie, side1, side_off1, rf_i1) < 2) return;
    if (clip_segment(c2, c1, side2, side_off2, rf_i2) < 2) return;
    if (primary_is_edge) { m->local_normal = rf_normal; m->local_point = rf_v1; }
    else { m->local_normal = sh.n[rf_i1]; m->local_point = sh.v[rf_i1]; }
    int count = 0;
    for (int i = 0; i < 2; ++i) {
        const float separation = dot(rf_normal, c2[i].v - rf_v1);
        if (separation <= radius) {
            if (primary_is_edge) {
                m->lp[count] = rmulT(q, c2[i].v - p);
                m->key[count] = c2[i].key;
            } else {
                m->lp[count] = c2[i].v;
                const uint32_t k = c2[i].key;   // swap A and B features
                m->key[count] = make_key((int)((k >> 8) & 0xff), (int)(k & 0xff), (int)((k >> 24) & 0xff), (int)((k >> 16) & 0xff));
            }
            ++count;
        }
    }
    m->count = count;
}

// ---------------------------------------------------------------- ContactDetector (:305-346)
__device__ __forceinline__ void begin_contact(Lander& L, int body) {
    L.flags &= ~(F_FUSELAGE | F_LEG0 | F_LEG1);          // :316, :324 clear every flag ...
    if (body == 0) L.flags |= F_FUSELAGE;                // :317-320 ... then set the bodies of THIS contact
    if (body == 1) L.flags |= F_LEG0;
    if (body == 2) L.flags |= F_LEG1;
}
__device__ __forceinline__ void end_contact(Lander& L, int body) {
    if (body == 1) L.flags &= ~F_LEG0;                   // :337-343 legs only; the fuselage flag is never cleared
    if (body == 2) L.flags &= ~F_LEG1;
}

// ---------------------------------------------------------------- solver work structures
struct VelPoint { V2 rb; float normal_impulse, tangent_impulse, normal_mass, tangent_mass, velocity_bias; };
struct ActiveContact {
    int body, edge, count;
    Manifold m;
    float friction;
    V2 normal;
    VelPoint p[2];
    float k11, k12, k22, nm11, nm12, nm21, nm22;   // K and its inverse (block solver)
};

// effective-mass matrix of a revolute joint plus everything b2Mat33::Solve33 / Solve22 compute from the matrix
// alone -- cross(ey, ez) and the two reciprocal determinants -- evaluated once per step instead of in each of the
// 180 velocity iterations (same operations on the same operands: same bits)
struct JointWork {
    V2 ra, rb;
    float m_exx, m_eyx, m_ezx, m_eyy, m_ezy, m_ezz, motor_mass;
    float c1x, c1y, c1z, inv_det33, inv_det22;
};

__device__ __forceinline__ float inv_det22(float a11, float a12, float a21, float a22) {
    float det = a11 * a22 - a12 * a21;
    if (det != 0.0f) det = 1.0f / det;
    return det;
}
// b2Mat22::Solve / b2Mat33::Solve22 with the reciprocal determinant given
__device__ __forceinline__ V2 solve22_pre(float a11, float a12, float a21, float a22, float det, V2 b) {
    return mk(det * (a22 * b.x - a12 * b.y), det * (a11 * b.y - a21 * b.x));
}
__device__ __forceinline__ V2 solve22(float a11, float a12, float a21, float a22, V2 b) {
    return solve22_pre(a11, a12, a21, a22, inv_det22(a11, a12, a21, a22), b);
}

// Body-indexed access with a run-time body index, written as selects over compile-time indices so that
// the per-body arrays stay in registers (a dynamically indexed array would be demoted to local memory).
#define GETB(arr, B) ((B) == 0 ? arr[0] : ((B) == 1 ? arr[1] : arr[2]))
#define SETB(arr, B, val) do { if ((B) == 0) arr[0] = (val); else if ((B) == 1) arr[1] = (val); else arr[2] = (val); } while (0)

// ---------------------------------------------------------------- World.Step(dt) for one lander
__device__ __forceinline__ void set_awake(Lander& L, bool awake) {
    if (awake) {
        if (!(L.flags & F_AWAKE)) { L.flags |= F_AWAKE; for (int i = 0; i < 3; ++i) L.b[i].sleep_time = 0.0f; }
    } else {
        L.flags &= ~F_AWAKE;
        for (int i = 0; i < 3; ++i) { L.b[i].sleep_time = 0.0f; L.b[i].v = mk(0.0f, 0.0f); L.b[i].w = 0.0f; }
        L.force = mk(0.0f, 0.0f); L.torque = 0.0f;
    }
}

// ---------------------------------------------------------------- broad phase (b2BroadPhase / b2Fixture::Synchronize)
// A contact EXISTS while the fat boxes of its two fixtures overlap -- created by FindNewContacts (appended to the world's
// contact list, pushed on the FRONT of the body's contact list), destroyed by Collide when they stop overlapping -- and is
// TOUCHING while its manifold has points.  The order of creation is state: Collide fires BeginContact / EndContact in
// creation order ("last BeginContact wins", LunarLanderEnv.cs:316-329) and the island takes a body's contacts newest first.
struct Box { float lx, ly, hx, hy; };
__device__ __forceinline__ Box body_box(const Shape& sh, V2 p, Rot q) {   // b2PolygonShape::ComputeAABB
    V2 lo = rmul(q, sh.v[0]) + p, hi = lo;
    for (int i = 1; i < sh.count; ++i) {
        const V2 w = rmul(q, sh.v[i]) + p;
        lo = mk(minf(lo.x, w.x), minf(lo.y, w.y)); hi = mk(maxf(hi.x, w.x), maxf(hi.y, w.y));
    }
    return Box{lo.x - POLYGON_RADIUS, lo.y - POLYGON_RADIUS, hi.x + POLYGON_RADIUS, hi.y + POLYGON_RADIUS};
}
__device__ __forceinline__ Box edge_fat_box(const Lander& L, int e) {     // b2EdgeShape::ComputeAABB grown once by aabbExtension (static: never moves)
    V2 v1, v2;
    edge_points(L, e, &v1, &v2);
    const float lx = minf(v1.x, v2.x) - POLYGON_RADIUS, ly = minf(v1.y, v2.y) - POLYGON_RADIUS;
    const float hx = maxf(v1.x, v2.x) + POLYGON_RADIUS, hy = maxf(v1.y, v2.y) + POLYGON_RADIUS;
    return Box{lx - AABB_EXTENSION, ly - AABB_EXTENSION, hx + AABB_EXTENSION, hy + AABB_EXTENSION};
}
__device__ __forceinline__ bool boxes_overlap(const Box& a, const Box& b) {
    if (b.lx - a.hx > 0.0f || b.ly - a.hy > 0.0f) return false;
    if (a.lx - b.hx > 0.0f || a.ly - b.hy > 0.0f) return false;
    return true;
}
__device__ __forceinline__ Box fat_of(const Lander& L, int body) { return Box{L.fat[body][0], L.fat[body][1], L.fat[body][2], L.fat[body][3]}; }
__device__ __forceinline__ uint32_t pair_at(const Lander& L, int k) { return (L.pairs[k >> 2] >> (8 * (k & 3))) & 0xffu; }
__device__ __forceinline__ void pair_put(Lander& L, int k, uint32_t v) { L.pairs[k >> 2] = (L.pairs[k >> 2] & ~(0xffu << (8 * (k & 3)))) | (v << (8 * (k & 3))); }
__device__ __forceinline__ int pair_count(const Lander& L) { int n = 0; while (n < MAXP && pair_at(L, n) != 0xffu) ++n; return n; }

// ---------------------------------------------------------------- register-resident constraint rows
// The 180 velocity and up to 60 position iterations are one long dependent float32 chain per lander, and a launch lasts as
// long as its slowest lander.  The first ROWS manifolds of the island (in island order) keep every operand of that chain in
// registers; the iteration loops are the SAME loops for every lane of a warp -- two joints, then row 0, row 1, ... each
// skipped by the lanes that have fewer rows -- so a warp that mixes landers without contact, with one manifold and with four
// executes the joints once per iteration, not once per kind of lander (ncu, round 2: the three separate loops of the first
// version ran back to back in almost every warp, 6.5 of 32 lanes active).  Same operations in the same order as plain
// per-contact loops: same bits.
struct VelRow {
    V2 normal, rb0, rb1;
    float ni0, ni1, ti0, ti1, nm0, nm1, tm0, tm1, friction;
    float k11, k12, k22, i11, i12, i21, i22;
    float inv_mass, inv_inertia;
    int count, body;   // velocity-constraint points of the manifold, the body it acts on
};
struct PosRow { V2 local_normal, local_point, lp0, lp1, centroid; float inv_mass, inv_inertia; int type, count, body; };

__device__ __forceinline__ void friction_point(V2& vB, float& wB, float mB, float iB, V2 tangent, V2 rb, float tangent_mass, float friction, float normal_impulse, float& tangent_impulse) {
    const V2 dv = vB + cross_sv(wB, rb);
    const float vt = dot(dv, tangent);
    float lambda = tangent_mass * (-vt);
    const float max_friction = friction * normal_impulse;
    const float new_impulse = clampf(tangent_impulse + lambda, -max_friction, max_friction);
    lambda = new_impulse - tangent_impulse;
    tangent_impulse = new_impulse;
    const V2 P = lambda * tangent;
    vB = vB + mB * P;
    wB = wB + iB * cross(rb, P);
}

// b2ContactSolver::SolveVelocityConstraints for one manifold of a body against the static ground (friction first, then the
// normal row or the 2-point block LCP by enumeration of its four cases)
__device__ __forceinline__ void solve_velocity_row(VelRow& r, V2& vB, float& wB, float mB, float iB) {
    const V2 normal = r.normal;
    const V2 tangent = cross_vs(normal, 1.0f);
    friction_point(vB, wB, mB, iB, tangent, r.rb0, r.tm0, r.friction, r.ni0, r.ti0);
    if (r.count == 2) friction_point(vB, wB, mB, iB, tangent, r.rb1, r.tm1, r.friction, r.ni1, r.ti1);
    if (r.count == 1) {
        const V2 dv = vB + cross_sv(wB, r.rb0);
        const float vn = dot(dv, normal);
        float lambda = -r.nm0 * vn;   // velocity bias 0: restitution 0 (:242, :268)
        const float new_impulse = maxf(r.ni0 + lambda, 0.0f);
        lambda = new_impulse - r.ni0;
        r.ni0 = new_impulse;
        const V2 P = lambda * normal;
        vB = vB + mB * P;
        wB = wB + iB * cross(r.rb0, P);
    } else {
        const V2 aa = mk(r.ni0, r.ni1);
        const V2 dv1 = vB + cross_sv(wB, r.rb0);
        const V2 dv2 = vB + cross_sv(wB, r.rb1);
        const float vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
        V2 b = mk(vn1, vn2);
        b = b - mk(r.k11 * aa.x + r.k12 * aa.y, r.k12 * aa.x + r.k22 * aa.y);
        // the four candidates, evaluated without branches, first valid one taken
        const V2 x1 = -mk(r.i11 * b.x + r.i12 * b.y, r.i21 * b.x + r.i22 * b.y);
        const bool ok1 = x1.x >= 0.0f && x1.y >= 0.0f;
        const float x2 = -r.nm0 * b.x;
        const bool ok2 = x2 >= 0.0f && (r.k12 * x2 + b.y) >= 0.0f;
        const float y3 = -r.nm1 * b.y;
        const bool ok3 = y3 >= 0.0f && (r.k12 * y3 + b.x) >= 0.0f;
        const bool ok4 = b.x >= 0.0f && b.y >= 0.0f;
        V2 x = mk(0.0f, 0.0f);
        if (ok3) x = mk(0.0f, y3);
        if (ok2) x = mk(x2, 0.0f);
        if (ok1) x = x1;
        if (ok1 || ok2 || ok3 || ok4) {
            const V2 d = x - aa;
            const V2 P1 = d.x * normal, P2 = d.y * normal;
            vB = vB + mB * (P1 + P2);
            wB = wB + iB * (cross(r.rb0, P1) + cross(r.rb1, P2));
            r.ni0 = x.x;
            r.ni1 = x.y;
        }
    }
}

// one point of b2ContactSolver::SolvePositionConstraints (pseudo-impulse with Baumgarte, body against the static ground)
__device__ __forceinline__ void position_point(int type, V2 local_normal, V2 local_point, V2 lp, V2 centroid, float mB, float iB, V2& cB, float& aB, float& min_separation) {
    const Rot qB = rot(aB);
    const V2 pB = cB - rmul(qB, centroid);
    V2 normal, point;
    float separation;
    if (type == MF_FACE_A) {
        normal = local_normal;
        const V2 clip = rmul(qB, lp) + pB;
        separation = dot(clip - local_point, normal) - POLYGON_RADIUS - POLYGON_RADIUS;
        point = clip;
    } else {
        normal = rmul(qB, local_normal);
        const V2 plane = rmul(qB, local_point) + pB;
        separation = dot(lp - plane, normal) - POLYGON_RADIUS - POLYGON_RADIUS;
        point = lp;
        normal = -normal;
    }
    const V2 rB = point - cB;
    min_separation = minf(min_separation, separation);
    const float C = clampf(BAUMGARTE * (separation + LINEAR_SLOP), -MAX_LINEAR_CORRECTION, 0.0f);
    const float rnB = cross(rB, normal);
    const float K = mB + iB * rnB * rnB;
    const float impulse = K > 0.0f ? -C / K : 0.0f;
    const V2 P = impulse * normal;
    cB = cB + mB * P;
    aB = aB + iB * cross(rB, P);
}

// HAS_PAIRS = false: the caller guarantees that no contact exists (the pair list is empty) -- the narrow phase and every
// contact row compile out (the free-flight kernel).

// ---- branch-free inner loops ------------------------------------------------------------------------------------------
// A velocity or position iteration below is ONE basic block: every data-dependent choice (limit active / released /
// inactive, one- or two-point manifold, the four cases of the block solver, face-A / face-B manifolds, a body without a
// manifold) is a select over values that are all computed, never a branch.  The chains of the three bodies' contact rows
// are independent inside an iteration, and with no branch between them the scheduler interleaves them: the launch is bound
// by the dependent float32 latency of its slowest lander, and instruction-level parallelism is the only thing that shortens
// it (ncu, round 2: the branchy version issued 0.25 instructions per cycle on the sub-partitions that hold such a lander).
// Every selected value is produced by exactly the operations of the plain per-contact loops: same bits.

// b2RevoluteJoint::SolveVelocityConstraints of joint ji (fuselage A = 0, leg B = 1 + ji).  limit_state is never
// LIMIT_EQUAL for these joints (|upper - lower| = 0.5 >= 2 angularSlop, set by InitVelocityConstraints every step).
template <int ji>
__device__ __forceinline__ void joint_velocity(Joint& J, const JointWork& W, V2& vA, float& wA, V2& vB, float& wB) {
    constexpr int A = 0, B = 1 + ji;
    const float mA = SHAPES[A].inv_mass, iA = SHAPES[A].inv_inertia, mB = SHAPES[B].inv_mass, iB = SHAPES[B].inv_inertia;
    {   // motor
        const float Cdot = wB - wA - JOINTS[ji].motor_speed;
        float impulse = -W.motor_mass * Cdot;
        const float old_impulse = J.motor;
        const float max_impulse = DT * MAX_MOTOR_TORQUE;
        J.motor = clampf(old_impulse + impulse, -max_impulse, max_impulse);
        impulse = J.motor - old_impulse;
        wA = wA - iA * impulse;
        wB = wB + iB * impulse;
    }
    const bool active = J.limit_state != LIMIT_INACTIVE;
    const V2 Cdot1 = vB + cross_sv(wB, W.rb) - vA - cross_sv(wA, W.ra);
    const float Cdot2 = wB - wA;
    // impulse = -mass.Solve33(Cdot), mass symmetric: ex=(exx,eyx,ezx) ey=(eyx,eyy,ezy) ez=(ezx,ezy,ezz); cross(ey, ez) and 1 / det from JointWork
    const float exx = W.m_exx, exy = W.m_eyx, exz = W.m_ezx;
    const float eyx = W.m_eyx, eyy = W.m_eyy, eyz = W.m_ezy;
    const float ezx = W.m_ezx, ezy = W.m_ezy, ezz = W.m_ezz;
    const float bx = Cdot1.x, by = Cdot1.y, bz = Cdot2;
    const float det = W.inv_det33;
    const float sx = det * (bx * W.c1x + by * W.c1y + bz * W.c1z);
    const float c2x = by * ezz - bz * ezy, c2y = bz * ezx - bx * ezz, c2z = bx * ezy - by * ezx;   // cross(b, ez)
    const float sy = det * (exx * c2x + exy * c2y + exz * c2z);
    const float c3x = eyy * bz - eyz * by, c3y = eyz * bx - eyx * bz, c3z = eyx * by - eyy * bx;   // cross(ey, b)
    const float sz = det * (exx * c3x + exy * c3y + exz * c3z);
    const float new_impulse = J.iz + (-sz);
    const bool release = active && (J.limit_state == LIMIT_AT_LOWER ? new_impulse < 0.0f : new_impulse > 0.0f);
    // the 2x2 solve of the released limit (rhs = -Cdot1 + impulse.z * ez.xy) and of the inactive limit (rhs = -Cdot) is one solve of the selected rhs
    const V2 rhs_released = -Cdot1 + J.iz * mk(W.m_ezx, W.m_ezy);
    const V2 rhs_inactive = -Cdot1;
    const V2 rhs = active ? rhs_released : rhs_inactive;
    const V2 red = solve22_pre(W.m_exx, W.m_eyx, W.m_eyx, W.m_eyy, W.inv_det22, rhs);
    const bool full = active && !release;
    const float imx = full ? -sx : red.x, imy = full ? -sy : red.y;
    const float imz = full ? -sz : -J.iz;        // (inactive: not used)
    J.ix += imx; J.iy += imy;
    J.iz = full ? J.iz + imz : (release ? 0.0f : J.iz);
    const V2 P = mk(imx, imy);
    const float ta = cross(W.ra, P), tb = cross(W.rb, P);
    vA = vA - mA * P;
    wA = wA - iA * (active ? ta + imz : ta);
    vB = vB + mB * P;
    wB = wB + iB * (active ? tb + imz : tb);
}

// A select the compiler cannot turn back into a branch: the condition is an all-ones / all-zeros word whose origin is hidden
// behind an empty asm, applied with integer and / or (one LOP3).  With plain `c ? a : b` nvcc sinks a row's whole
// computation under `if (row exists)` -- three divergent regions per iteration, executed one after the other (ncu, round 2:
// 12 of 32 lanes active in the row code) -- and the three independent chains can no longer be interleaved.
__device__ __forceinline__ int opaque_mask(bool c) {
    int m = c ? -1 : 0;
    asm volatile("" : "+r"(m));
    return m;
}
__device__ __forceinline__ float msel(int m, float a, float b) { return __int_as_float((__float_as_int(a) & m) | (__float_as_int(b) & ~m)); }
__device__ __forceinline__ V2 msel(int m, V2 a, V2 b) { return mk(msel(m, a.x, b.x), msel(m, a.y, b.y)); }

// b2ContactSolver::SolveVelocityConstraints for the manifold of one body against the static ground, committed only if the
// row exists (r.count > 0): friction of the points, then the normal row (one point) or the block LCP (two points).
__device__ __forceinline__ void velocity_row_select(VelRow& r, V2& v, float& w) {
    const bool two = r.count == 2;
    const int has = opaque_mask(r.count > 0), has2 = opaque_mask(r.count == 2);
    const float mB = r.inv_mass, iB = r.inv_inertia;
    const V2 normal = r.normal;
    const V2 tangent = cross_vs(normal, 1.0f);
    V2 vB = v; float wB = w;
    float ti0 = r.ti0, ti1 = r.ti1;
    friction_point(vB, wB, mB, iB, tangent, r.rb0, r.tm0, r.friction, r.ni0, ti0);
    {
        V2 v2 = vB; float w2 = wB; float t2 = ti1;
        friction_point(v2, w2, mB, iB, tangent, r.rb1, r.tm1, r.friction, r.ni1, t2);
        vB = two ? v2 : vB; wB = two ? w2 : wB; ti1 = two ? t2 : ti1;
    }
    // one point
    V2 v1; float w1, n1;
    {
        const V2 dv = vB + cross_sv(wB, r.rb0);
        const float vn = dot(dv, normal);
        float lambda = -r.nm0 * vn;   // velocity bias 0: restitution 0 (:242, :268)
        n1 = maxf(r.ni0 + lambda, 0.0f);
        lambda = n1 - r.ni0;
        const V2 P = lambda * normal;
        v1 = vB + mB * P;
        w1 = wB + iB * cross(r.rb0, P);
    }
    // two points: the LCP by enumeration, first valid candidate
    V2 v2p; float w2p, n20, n21;
    {
        const V2 aa = mk(r.ni0, r.ni1);
        const V2 dv1 = vB + cross_sv(wB, r.rb0);
        const V2 dv2 = vB + cross_sv(wB, r.rb1);
        const float vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
        V2 b = mk(vn1, vn2);
        b = b - mk(r.k11 * aa.x + r.k12 * aa.y, r.k12 * aa.x + r.k22 * aa.y);
        const V2 x1 = -mk(r.i11 * b.x + r.i12 * b.y, r.i21 * b.x + r.i22 * b.y);
        const bool ok1 = x1.x >= 0.0f && x1.y >= 0.0f;
        const float x2 = -r.nm0 * b.x;
        const bool ok2 = x2 >= 0.0f && (r.k12 * x2 + b.y) >= 0.0f;
        const float y3 = -r.nm1 * b.y;
        const bool ok3 = y3 >= 0.0f && (r.k12 * y3 + b.x) >= 0.0f;
        const bool ok4 = b.x >= 0.0f && b.y >= 0.0f;
        V2 x = mk(0.0f, 0.0f);
        x = ok3 ? mk(0.0f, y3) : x;
        x = ok2 ? mk(x2, 0.0f) : x;
        x = ok1 ? x1 : x;
        const bool solved = ok1 || ok2 || ok3 || ok4;
        const V2 d = x - aa;
        const V2 P1 = d.x * normal, P2 = d.y * normal;
        const V2 vs = vB + mB * (P1 + P2);
        const float ws = wB + iB * (cross(r.rb0, P1) + cross(r.rb1, P2));
        v2p = solved ? vs : vB; w2p = solved ? ws : wB;
        n20 = solved ? x.x : r.ni0; n21 = solved ? x.y : r.ni1;
    }
    const V2 vn_ = two ? v2p : v1;
    const float wn_ = two ? w2p : w1;
    v = msel(has, vn_, v);
    w = msel(has, wn_, w);
    r.ti0 = msel(has, ti0, r.ti0); r.ti1 = msel(has, ti1, r.ti1);
    r.ni0 = msel(has, two ? n20 : n1, r.ni0);
    r.ni1 = msel(has2, n21, r.ni1);
}

// rotation of a body angle known to be far below the 32768 rad where the argument reduction changes path (checked once per step)
__device__ __forceinline__ Rot rot_in_range(float a) { Rot q; const float2 sc = sincos_det<true>(a); q.s = sc.x; q.c = sc.y; return q; }

// one point of b2ContactSolver::SolvePositionConstraints, both manifold types computed, committed only if `on`
template <bool IN_RANGE>
__device__ __forceinline__ void position_point_select(bool on, int type, V2 local_normal, V2 local_point, V2 lp, V2 centroid, float mB, float iB,
                                                      V2& c, float& a, float& min_separation) {
    const Rot qB = IN_RANGE ? rot_in_range(a) : rot(a);
    const V2 pB = c - rmul(qB, centroid);
    // face A: the plane is the (static) edge's, the clip point rides on the body
    const V2 clipA = rmul(qB, lp) + pB;
    const float sepA = dot(clipA - local_point, local_normal) - POLYGON_RADIUS - POLYGON_RADIUS;
    // face B: the plane rides on the body, the clip point is the edge's
    const V2 nB = rmul(qB, local_normal);
    const V2 planeB = rmul(qB, local_point) + pB;
    const float sepB = dot(lp - planeB, nB) - POLYGON_RADIUS - POLYGON_RADIUS;
    const bool faceA = type == MF_FACE_A;
    const V2 normal = faceA ? local_normal : -nB;
    const V2 point = faceA ? clipA : lp;
    const float separation = faceA ? sepA : sepB;
    const V2 rB = point - c;
    const float ms = minf(min_separation, separation);
    const float C = clampf(BAUMGARTE * (separation + LINEAR_SLOP), -MAX_LINEAR_CORRECTION, 0.0f);
    const float rnB = cross(rB, normal);
    const float K = mB + iB * rnB * rnB;
    // K >= inv_mass > 0 for a real row; -C / K is 0 or far from the subnormal range (C is a multiple of ulp(linearSlop) ~ 5e-10 scaled by 0.2).
    // The `K > 0` of b2ContactSolver (false only for a NaN here) is kept as an opaque-mask select: as a branch it put a
    // convergence barrier around every one of the six points of an iteration and serialised the three bodies' chains.
    const int kpos = opaque_mask(K > 0.0f);
    const float impulse = msel(kpos, div_inrange(-C, msel(kpos, K, 1.0f)), 0.0f);
    const V2 P = impulse * normal;
    const V2 cn = c + mB * P;
    const float an = a + iB * cross(rB, P);
    c = on ? cn : c;
    a = on ? an : a;
    min_separation = on ? ms : min_separation;
}

// b2RevoluteJoint::SolvePositionConstraints of joint ji; returns position_error <= linearSlop && angular_error <= angularSlop
template <int ji, bool IN_RANGE>
__device__ __forceinline__ bool joint_position(const Joint& J, float motor_mass, V2& cA, float& aA, V2& cB, float& aB) {
    constexpr int A = 0, B = 1 + ji;
    const float mA = SHAPES[A].inv_mass, iA = SHAPES[A].inv_inertia, mB = SHAPES[B].inv_mass, iB = SHAPES[B].inv_inertia;
    const bool active = J.limit_state != LIMIT_INACTIVE, lower = J.limit_state == LIMIT_AT_LOWER;
    const float angle = aB - aA - JOINTS[ji].ref_angle;
    const float Cl = angle - JOINTS[ji].lower, Cu = angle - JOINTS[ji].upper;
    const float angular_error = active ? (lower ? -Cl : Cu) : 0.0f;
    const float Cl2 = clampf(Cl + ANGULAR_SLOP, -MAX_ANGULAR_CORRECTION, 0.0f);
    const float Cu2 = clampf(Cu - ANGULAR_SLOP, 0.0f, MAX_ANGULAR_CORRECTION);
    const float limit_impulse = -motor_mass * (lower ? Cl2 : Cu2);
    const float aAn = aA - iA * limit_impulse, aBn = aB + iB * limit_impulse;
    aA = active ? aAn : aA;
    aB = active ? aBn : aB;
    const Rot qA = IN_RANGE ? rot_in_range(aA) : rot(aA), qB = IN_RANGE ? rot_in_range(aB) : rot(aB);
    const V2 rA = rmul(qA, mk(0.0f, 0.0f) - SHAPES[A].centroid);
    const V2 rB = rmul(qB, JOINTS[ji].anchor_b - SHAPES[B].centroid);
    const V2 C = cB + rB - cA - rA;
    // position_error = sqrt(dot(C, C)) is only compared with linearSlop: sqrt is correctly rounded and monotonic, so
    // `sqrtf(x) <= LINEAR_SLOP` holds exactly for x <= LINEAR_SLOP_SQ_MAX, the largest float32 whose rounded root does not
    // exceed LINEAR_SLOP (0x1.a36e3p-16; tests/test_host_helpers.py checks the boundary) -- no MUFU.RSQ + fix-up call in the loop
    const float position_error_sq = dot(C, C);
    const float kxx = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
    const float kxy = -iA * rA.x * rA.y - iB * rB.x * rB.y;
    const float kyy = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
    // b2Mat22::Solve; the determinant of this K is of order (1 / m_leg)^2: normal, never 0 -- its reciprocal without the range check of `/`
    float det = kxx * kyy - kxy * kxy;
    det = det != 0.0f ? div_inrange(1.0f, det) : det;
    const V2 impulse = -solve22_pre(kxx, kxy, kxy, kyy, det, C);
    cA = cA - mA * impulse;
    aA = aA - iA * cross(rA, impulse);
    cB = cB + mB * impulse;
    aB = aB + iB * cross(rB, impulse);
    return position_error_sq <= LINEAR_SLOP_SQ_MAX && angular_error <= ANGULAR_SLOP;
}

// ---- three lanes per lander (TRIO) ------------------------------------------------------------------------------------
// The contact class of the partition is latency-bound: ONE warp per sub-partition runs the dependent chain of 180 velocity
// and up to 60 position iterations, and nothing but a shorter instruction stream per iteration shortens the launch.  In the
// TRIO variant three neighbouring lanes (sub = 0, 1, 2 = fuselage, leg 0, leg 1) step the SAME lander: everything outside
// the iteration loops is executed redundantly (identical inputs, identical results, no extra time in SIMT), and inside the
// loops every lane keeps only ITS body and ITS manifold rows -- a lane runs one contact row per iteration instead of three.
// A joint couples the fuselage with one leg: all three lanes fetch both bodies with shuffles and evaluate the joint (the
// third lane is a clone of the leg's lane, so every lane's accumulators stay right), and each commits only its own body.
// The order of operations per body is that of the plain loops (joint 1, joint 0, then the body's rows; rows of different
// bodies commute): same bits.
__device__ __forceinline__ float trio_get(unsigned tmask, float x, int src) { return __shfl_sync(tmask, x, src); }
__device__ __forceinline__ V2 trio_get(unsigned tmask, V2 x, int src) { return mk(__shfl_sync(tmask, x.x, src), __shfl_sync(tmask, x.y, src)); }

template <int ji>
__device__ __forceinline__ void joint_velocity_trio(Joint& J, const JointWork& W, unsigned tmask, int base, int sub, V2& vm, float& wm) {
    V2 vA = trio_get(tmask, vm, base), vB = trio_get(tmask, vm, base + 1 + ji);
    float wA = trio_get(tmask, wm, base), wB = trio_get(tmask, wm, base + 1 + ji);
    joint_velocity<ji>(J, W, vA, wA, vB, wB);
    const bool isA = sub == 0, isB = sub == 1 + ji;
    vm = isA ? vA : (isB ? vB : vm);
    wm = isA ? wA : (isB ? wB : wm);
}

template <int ji, bool IN_RANGE>
__device__ __forceinline__ bool joint_position_trio(const Joint& J, float motor_mass, unsigned tmask, int base, int sub, V2& cm, float& am) {
    V2 cA = trio_get(tmask, cm, base), cB = trio_get(tmask, cm, base + 1 + ji);
    float aA = trio_get(tmask, am, base), aB = trio_get(tmask, am, base + 1 + ji);
    const bool ok = joint_position<ji, IN_RANGE>(J, motor_mass, cA, aA, cB, aB);
    const bool isA = sub == 0, isB = sub == 1 + ji;
    cm = isA ? cA : (isB ? cB : cm);
    am = isA ? aA : (isB ? aB : am);
    return ok;
}

// the touching manifolds of a step (creation order) and their island order
struct Contacts { ActiveContact ac[MAXC]; int nc; uint32_t order; };

// ---- ContactManager.Collide: every existing contact, in creation order
template <bool HAS_PAIRS>
__device__ __forceinline__ void collide(Lander& L, Contacts& C) {
    ActiveContact* ac = C.ac;
    int nc = 0;
    if (HAS_PAIRS && (L.flags & F_AWAKE)) {
        Rot q[3]; V2 p[3];
#pragma unroll
        for (int body = 0; body < 3; ++body) { q[body] = rot(L.b[body].a); p[body] = L.b[body].c - rmul(q[body], SHAPES[body].centroid); }
        const int np = pair_count(L);
        int kept = 0;
        for (int k = 0; k < np; ++k) {
            const uint32_t pr = pair_at(L, k);
            const int body = (int)(pr >> 4), e = (int)(pr & 15u);
            const bool was = (L.touch[body] >> e) & 1u;
            if (!boxes_overlap(edge_fat_box(L, e), fat_of(L, body))) {   // the fat boxes parted: the contact is destroyed
                if (was) { end_contact(L, body); L.touch[body] &= ~(1u << e); }
                continue;
            }
            pair_put(L, kept++, pr);
            V2 v1, v2;
            edge_points(L, e, &v1, &v2);
            Manifold m;
            m.count = 0; m.type = 0;
            if (!edge_polygon_apart(v1, v2, body, GETB(p, body), GETB(q, body)))
                collide_edge_polygon(&m, v1, v2, SHAPES[body], GETB(p, body), GETB(q, body));
            const bool touching = m.count > 0 && nc < MAXC;
            if (touching) {
                L.touch[body] |= 1u << e;
                // b2Contact::Update: match old manifold points by id, copy their impulses (warm start).  Matched into scalars
                // first: stores through ac[nc] inside the slot / point / id loops were expanded by the compiler into one copy
                // per possible nc (10 000 instructions of selects)
                const int32_t pair = body * 16 + e;
                float wni[2] = {0.0f, 0.0f}, wti[2] = {0.0f, 0.0f};
#pragma unroll
                for (int sl = 0; sl < MAXC; ++sl) {
                    const bool slot = L.c[sl].pair == pair;
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int o = 1; o >= 0; --o) {   // the FIRST old point with this id wins (b2Contact::Update breaks at the first match; a face-B manifold can carry the same id twice)
                            const bool hit = slot && j < m.count && L.c[sl].key[o] != NO_KEY && L.c[sl].key[o] == m.key[j];
                            wni[j] = hit ? L.c[sl].ni[o] : wni[j];
                            wti[j] = hit ? L.c[sl].ti[o] : wti[j];
                        }
                }
                ActiveContact& c = ac[nc++];
                c.body = body; c.edge = e; c.count = m.count; c.m = m;
                for (int j = 0; j < 2; ++j) { c.p[j].normal_impulse = wni[j]; c.p[j].tangent_impulse = wti[j]; }
            } else {
                L.touch[body] &= ~(1u << e);
            }
            if (touching && !was) begin_contact(L, body);
            if (!touching && was) end_contact(L, body);
        }
        for (int k = kept; k < np; ++k) pair_put(L, k, 0xffu);
    }
    // island order of the touching contacts: bodies in DFS order (fuselage, leg 0, leg 1), each body's contacts newest first;
    // four bits per entry
    uint32_t order = 0u;
    if (HAS_PAIRS) {
        int no = 0;
#pragma unroll
        for (int body = 0; body < 3; ++body)
            for (int k = nc - 1; k >= 0; --k)
                if (ac[k].body == body) order |= (uint32_t)k << (4 * no++);
    }
    C.nc = nc; C.order = order;
}

// Timing probe only (tools/lunar_phase_probe.py builds a variant with -DLUNAR_PHASE_CLOCKS): clock64() of every thread at the
// phase boundaries of a step, read back through gymcuda_debug_lunar_phase.  Never defined in the product build.
#ifdef LUNAR_PHASE_CLOCKS
__device__ long long g_lunar_phase[2][10][65536];
#define LUNAR_PHASE(HP, k) do { const unsigned _t = blockIdx.x * blockDim.x + threadIdx.x; if (_t < 65536u) g_lunar_phase[(HP) ? 1 : 0][k][_t] = clock64(); } while (0)
// slots 8 / 9: %globaltimer (ns, one clock for the whole GPU) at the start / end of the step -- when the two kernels of the partition really ran
#define LUNAR_PHASE_NS(HP, k) do { const unsigned _t = blockIdx.x * blockDim.x + threadIdx.x; if (_t < 65536u) { unsigned long long _ns; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_ns)); g_lunar_phase[(HP) ? 1 : 0][k][_t] = (long long)_ns; } } while (0)
#else
#define LUNAR_PHASE_NS(HP, k) do {} while (0)
#define LUNAR_PHASE(HP, k) do {} while (0)
#endif

// ---- Solve (b2Island::Solve) -- the island is always {fuselage, leg0, leg1} -- then the broad-phase update and ClearForces
template <bool HAS_PAIRS, bool TRIO = false>
__device__ __forceinline__ void solve(Lander& L, Contacts& C) {
    static_assert(HAS_PAIRS || !TRIO, "the free-flight class is throughput-bound: one lane per lander");
    const float h = DT;
    const float dt_ratio = (L.flags & F_FIRST_STEP) ? 0.0f : 1.0f;   // inv_dt0 * dt: 0 on a new World, then 50 * 0.02f = 1
    ActiveContact* ac = C.ac;
    const int nc = HAS_PAIRS ? C.nc : 0;
    const uint32_t order = C.order;
    if (L.flags & F_AWAKE) {
        V2 c[3], v[3];
        float a[3], w[3];
        V2 c0[3]; float a0[3];   // the sweep's c0 / a0: the pose this step started from
        Joint Jl[2] = {L.j[0], L.j[1]};   // joint accumulators live in registers during the iterations
        const V2 gravity = mk(0.0f, L.gravity);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            c[i] = L.b[i].c; a[i] = L.b[i].a; v[i] = L.b[i].v; w[i] = L.b[i].w;
            c0[i] = c[i]; a0[i] = a[i];
            const V2 f = i == 0 ? L.force : mk(0.0f, 0.0f);
            const float tq = i == 0 ? L.torque : 0.0f;
            v[i] = v[i] + h * (gravity + SHAPES[i].inv_mass * f);
            w[i] = w[i] + h * SHAPES[i].inv_inertia * tq;
            v[i] = (1.0f / (1.0f + h * 0.0f)) * v[i];   // linear damping 0
            w[i] = w[i] * (1.0f / (1.0f + h * 0.0f));
        }

        // contact solver: InitializeVelocityConstraints (any order: rows do not interact here)
        for (int k = 0; k < nc; ++k) {
            ActiveContact& cc = ac[k];
            const int B = cc.body;
            V2 cB = GETB(c, B); float aB = GETB(a, B);
            const float mB = SHAPES[B].inv_mass, iB = SHAPES[B].inv_inertia;
            const Rot qB = rot(aB);
            const V2 pB = cB - rmul(qB, SHAPES[B].centroid);
            cc.friction = sqrtf(edge_friction(cc.edge) * SHAPES[B].friction);   // MixFriction
            V2 normal, pts[2];
            if (cc.m.type == MF_FACE_A) {
                normal = cc.m.local_normal;
                const V2 plane = cc.m.local_point;
                for (int j = 0; j < cc.count; ++j) {
                    const V2 clip = rmul(qB, cc.m.lp[j]) + pB;
                    const V2 cA = clip + (POLYGON_RADIUS - dot(clip - plane, normal)) * normal;
                    const V2 cB = clip - POLYGON_RADIUS * normal;
                    pts[j] = 0.5f * (cA + cB);
                }
            } else {
                normal = rmul(qB, cc.m.local_normal);
                const V2 plane = rmul(qB, cc.m.local_point) + pB;
                for (int j = 0; j < cc.count; ++j) {
                    const V2 clip = cc.m.lp[j];
                    const V2 cB = clip + (POLYGON_RADIUS - dot(clip - plane, normal)) * normal;
                    const V2 cA = clip - POLYGON_RADIUS * normal;
                    pts[j] = 0.5f * (cA + cB);
                }
                normal = -normal;
            }
            cc.normal = normal;
            const V2 tangent = cross_vs(normal, 1.0f);
            for (int j = 0; j < cc.count; ++j) {
                VelPoint& vp = cc.p[j];
                vp.normal_impulse = dt_ratio * vp.normal_impulse;
                vp.tangent_impulse = dt_ratio * vp.tangent_impulse;
                vp.rb = pts[j] - cB;
                const float rnB = cross(vp.rb, normal);
                const float kn = mB + iB * rnB * rnB;
                vp.normal_mass = kn > 0.0f ? 1.0f / kn : 0.0f;
                const float rtB = cross(vp.rb, tangent);
                const float kt = mB + iB * rtB * rtB;
                vp.tangent_mass = kt > 0.0f ? 1.0f / kt : 0.0f;
                vp.velocity_bias = 0.0f;   // restitution 0 (:242, :268)
            }
            if (cc.count == 2) {
                const float rn1 = cross(cc.p[0].rb, normal), rn2 = cross(cc.p[1].rb, normal);
                const float k11 = mB + iB * rn1 * rn1;
                const float k22 = mB + iB * rn2 * rn2;
                const float k12 = mB + iB * rn1 * rn2;
                const float k_max_condition = 1000.0f;
                if (k11 * k11 < k_max_condition * (k11 * k22 - k12 * k12)) {
                    cc.k11 = k11; cc.k12 = k12; cc.k22 = k22;
                    float det = k11 * k22 - k12 * k12;
                    if (det != 0.0f) det = 1.0f / det;
                    cc.nm11 = det * k22; cc.nm12 = -det * k12; cc.nm21 = -det * k12; cc.nm22 = det * k11;
                } else {
                    cc.count = 1;
                }
            }
        }
        // contact solver: WarmStart, island order
        for (int kk = 0; kk < nc; ++kk) {
            ActiveContact& cc = ac[(order >> (4 * kk)) & 15u];
            const int B = cc.body;
            V2 vB = GETB(v, B); float wB = GETB(w, B);
            const V2 tangent = cross_vs(cc.normal, 1.0f);
            for (int j = 0; j < cc.count; ++j) {
                const V2 P = cc.p[j].normal_impulse * cc.normal + cc.p[j].tangent_impulse * tangent;
                wB = wB + SHAPES[B].inv_inertia * cross(cc.p[j].rb, P);
                vB = vB + SHAPES[B].inv_mass * P;
            }
            SETB(v, B, vB); SETB(w, B, wB);
        }
        // joints: InitVelocityConstraints, island order leg1's joint, then leg0's
        JointWork jw[2];
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const int ji = 1 - jj;
            const int A = 0, B = 1 + ji;
            Joint& J = Jl[ji];
            JointWork& W = jw[ji];
            const float mA = SHAPES[A].inv_mass, iA = SHAPES[A].inv_inertia, mB = SHAPES[B].inv_mass, iB = SHAPES[B].inv_inertia;
            const Rot qA = rot(a[A]), qB = rot(a[B]);
            W.ra = rmul(qA, mk(0.0f, 0.0f) - SHAPES[A].centroid);
            W.rb = rmul(qB, JOINTS[ji].anchor_b - SHAPES[B].centroid);
            W.m_exx = mA + mB + W.ra.y * W.ra.y * iA + W.rb.y * W.rb.y * iB;
            W.m_eyx = -W.ra.y * W.ra.x * iA - W.rb.y * W.rb.x * iB;
            W.m_ezx = -W.ra.y * iA - W.rb.y * iB;
            W.m_eyy = mA + mB + W.ra.x * W.ra.x * iA + W.rb.x * W.rb.x * iB;
            W.m_ezy = W.ra.x * iA + W.rb.x * iB;
            W.m_ezz = iA + iB;
            W.motor_mass = iA + iB;
            if (W.motor_mass > 0.0f) W.motor_mass = 1.0f / W.motor_mass;
            {   // the matrix-only part of Solve33 (cross(ey, ez), det) and of Solve22
                const float exx = W.m_exx, exy = W.m_eyx, exz = W.m_ezx;
                const float eyx = W.m_eyx, eyy = W.m_eyy, eyz = W.m_ezy;
                const float ezx = W.m_ezx, ezy = W.m_ezy, ezz = W.m_ezz;
                W.c1x = eyy * ezz - eyz * ezy; W.c1y = eyz * ezx - eyx * ezz; W.c1z = eyx * ezy - eyy * ezx;
                float det = exx * W.c1x + exy * W.c1y + exz * W.c1z;
                if (det != 0.0f) det = 1.0f / det;
                W.inv_det33 = det;
                W.inv_det22 = inv_det22(W.m_exx, W.m_eyx, W.m_eyx, W.m_eyy);
            }
            const float angle = a[B] - a[A] - JOINTS[ji].ref_angle;
            if (fabsf(JOINTS[ji].upper - JOINTS[ji].lower) < 2.0f * ANGULAR_SLOP) {
                J.limit_state = LIMIT_EQUAL;
            } else if (angle <= JOINTS[ji].lower) {
                if (J.limit_state != LIMIT_AT_LOWER) J.iz = 0.0f;
                J.limit_state = LIMIT_AT_LOWER;
            } else if (angle >= JOINTS[ji].upper) {
                if (J.limit_state != LIMIT_AT_UPPER) J.iz = 0.0f;
                J.limit_state = LIMIT_AT_UPPER;
            } else {
                J.limit_state = LIMIT_INACTIVE;
                J.iz = 0.0f;
            }
            // warm start
            J.ix = J.ix * dt_ratio; J.iy = J.iy * dt_ratio; J.iz = J.iz * dt_ratio; J.motor = J.motor * dt_ratio;
            const V2 P = mk(J.ix, J.iy);
            v[A] = v[A] - mA * P;
            w[A] = w[A] - iA * (cross(W.ra, P) + J.motor + J.iz);
            v[B] = v[B] + mB * P;
            w[B] = w[B] + iB * (cross(W.rb, P) + J.motor + J.iz);
        }

        // ---- velocity iterations: one branch-free loop body for every lane.  Register rows: the newest manifold of each body
        // (rows of different bodies commute: they share no variable); a body's older manifolds (a polygon astride a terrain
        // vertex) follow from local memory, in island order, skipped by the whole warp when no lane has one.
        // TRIO: lane `sub` of the trio keeps body `sub` (vm, wm / cm, am), its newest manifold (rm / pm) and its older ones (extra_m)
        const unsigned tlane = TRIO ? (threadIdx.x & 31u) : 0u;
        const int sub = (int)(tlane % 3u), tbase = (int)tlane - sub;
        const unsigned tmask = 7u << tbase;
        int mine = -1;            // TRIO: island entry of my body's newest manifold
        uint32_t extra_m = 0u;    // TRIO: my body's older manifolds, island order
        int nextra_m = 0;
        VelRow row[(HAS_PAIRS && !TRIO) ? 3 : 1];
        uint32_t extra = 0u;   // island positions (4 bits each) of the manifolds that are not a body's newest
        int nextra = 0;
        bool warp_rows = false, warp_extra = false;
        bool warp_row0 = true;   // some lane of the warp holds a fuselage row (an extra manifold of the fuselage implies a newest one)
        if constexpr (TRIO) {
            int first_of[3] = {-1, -1, -1};
            for (int kk = 0; kk < nc; ++kk) {
                const int k = (int)((order >> (4 * kk)) & 15u);
                const int B = ac[k].body;
                if (GETB(first_of, B) < 0) { SETB(first_of, B, k); }
                else { extra |= (uint32_t)k << (4 * nextra++); if (B == sub) extra_m |= (uint32_t)k << (4 * nextra_m++); }
            }
            mine = GETB(first_of, sub);
            VelRow& q = row[0];
            q.count = 0; q.body = sub;
            q.inv_mass = SHAPES[sub].inv_mass; q.inv_inertia = SHAPES[sub].inv_inertia;
            q.normal = mk(0.0f, 1.0f); q.rb0 = q.rb1 = mk(0.0f, 0.0f);
            q.ni0 = q.ni1 = q.ti0 = q.ti1 = q.nm0 = q.nm1 = q.tm0 = q.tm1 = q.friction = 0.0f;
            q.k11 = q.k12 = q.k22 = q.i11 = q.i12 = q.i21 = q.i22 = 0.0f;
            if (mine >= 0) {
                const ActiveContact& cc = ac[mine];
                q.normal = cc.normal; q.rb0 = cc.p[0].rb; q.rb1 = cc.p[1].rb;
                q.ni0 = cc.p[0].normal_impulse; q.ni1 = cc.p[1].normal_impulse; q.ti0 = cc.p[0].tangent_impulse; q.ti1 = cc.p[1].tangent_impulse;
                q.nm0 = cc.p[0].normal_mass; q.nm1 = cc.p[1].normal_mass; q.tm0 = cc.p[0].tangent_mass; q.tm1 = cc.p[1].tangent_mass;
                q.friction = cc.friction;
                q.k11 = cc.k11; q.k12 = cc.k12; q.k22 = cc.k22; q.i11 = cc.nm11; q.i12 = cc.nm12; q.i21 = cc.nm21; q.i22 = cc.nm22;
                q.count = cc.count;
            }
            const unsigned am = __activemask();
            warp_rows = __ballot_sync(am, nc > 0) != 0u;
            warp_extra = __ballot_sync(am, nextra > 0) != 0u;
        } else if (HAS_PAIRS) {
            int first_of[3] = {-1, -1, -1};
            for (int kk = 0; kk < nc; ++kk) {
                const int k = (int)((order >> (4 * kk)) & 15u);
                const int B = ac[k].body;
                if (GETB(first_of, B) < 0) { SETB(first_of, B, k); }
                else { extra |= (uint32_t)k << (4 * nextra++); }
            }
#pragma unroll
            for (int B = 0; B < 3; ++B) {
                VelRow& q = row[B];
                q.count = 0; q.body = B;
                q.inv_mass = SHAPES[B].inv_mass; q.inv_inertia = SHAPES[B].inv_inertia;
                q.normal = mk(0.0f, 1.0f); q.rb0 = q.rb1 = mk(0.0f, 0.0f);
                q.ni0 = q.ni1 = q.ti0 = q.ti1 = q.nm0 = q.nm1 = q.tm0 = q.tm1 = q.friction = 0.0f;
                q.k11 = q.k12 = q.k22 = q.i11 = q.i12 = q.i21 = q.i22 = 0.0f;
                if (first_of[B] >= 0) {
                    const ActiveContact& cc = ac[first_of[B]];
                    q.normal = cc.normal; q.rb0 = cc.p[0].rb; q.rb1 = cc.p[1].rb;
                    q.ni0 = cc.p[0].normal_impulse; q.ni1 = cc.p[1].normal_impulse; q.ti0 = cc.p[0].tangent_impulse; q.ti1 = cc.p[1].tangent_impulse;
                    q.nm0 = cc.p[0].normal_mass; q.nm1 = cc.p[1].normal_mass; q.tm0 = cc.p[0].tangent_mass; q.tm1 = cc.p[1].tangent_mass;
                    q.friction = cc.friction;
                    q.k11 = cc.k11; q.k12 = cc.k12; q.k22 = cc.k22; q.i11 = cc.nm11; q.i12 = cc.nm12; q.i21 = cc.nm21; q.i22 = cc.nm22;
                    q.count = cc.count;
                }
            }
            const unsigned am = __activemask();
            warp_rows = __ballot_sync(am, nc > 0) != 0u;
            warp_extra = __ballot_sync(am, nextra > 0) != 0u;
            warp_row0 = __ballot_sync(am, row[0].count > 0) != 0u;
        }
        LUNAR_PHASE(HAS_PAIRS, 3);
        if constexpr (TRIO) {
            V2 vm = GETB(v, sub); float wm = GETB(w, sub);
            VelRow& rm = row[0];
#pragma unroll 1
            for (int it = 0; it < VELOCITY_ITERATIONS; ++it) {
                joint_velocity_trio<1>(Jl[1], jw[1], tmask, tbase, sub, vm, wm);   // island order: leg 1's joint, then leg 0's
                joint_velocity_trio<0>(Jl[0], jw[0], tmask, tbase, sub, vm, wm);
                if (warp_rows) {
                    velocity_row_select(rm, vm, wm);
                    if (warp_extra) {
                        for (int kk = 0; kk < nextra_m; ++kk) {
                            ActiveContact& cc = ac[(extra_m >> (4 * kk)) & 15u];
                            VelRow q;
                            q.normal = cc.normal; q.rb0 = cc.p[0].rb; q.rb1 = cc.p[1].rb;
                            q.ni0 = cc.p[0].normal_impulse; q.ni1 = cc.p[1].normal_impulse; q.ti0 = cc.p[0].tangent_impulse; q.ti1 = cc.p[1].tangent_impulse;
                            q.nm0 = cc.p[0].normal_mass; q.nm1 = cc.p[1].normal_mass; q.tm0 = cc.p[0].tangent_mass; q.tm1 = cc.p[1].tangent_mass;
                            q.friction = cc.friction;
                            q.k11 = cc.k11; q.k12 = cc.k12; q.k22 = cc.k22; q.i11 = cc.nm11; q.i12 = cc.nm12; q.i21 = cc.nm21; q.i22 = cc.nm22;
                            q.inv_mass = rm.inv_mass; q.inv_inertia = rm.inv_inertia;
                            q.count = cc.count; q.body = sub;
                            velocity_row_select(q, vm, wm);
                            cc.p[0].normal_impulse = q.ni0; cc.p[0].tangent_impulse = q.ti0;
                            cc.p[1].normal_impulse = q.ni1; cc.p[1].tangent_impulse = q.ti1;
                        }
                    }
                }
            }
            // every lane gets every body's velocities and impulses back: the rest of the step is redundant again
#pragma unroll
            for (int B = 0; B < 3; ++B) { v[B] = trio_get(tmask, vm, tbase + B); w[B] = trio_get(tmask, wm, tbase + B); }
            int first_of[3] = {-1, -1, -1};
            for (int kk = nc - 1; kk >= 0; --kk) { const int k = (int)((order >> (4 * kk)) & 15u); SETB(first_of, ac[k].body, k); }
#pragma unroll
            for (int B = 0; B < 3; ++B) {
                const float ni0 = trio_get(tmask, rm.ni0, tbase + B), ti0 = trio_get(tmask, rm.ti0, tbase + B);
                const float ni1 = trio_get(tmask, rm.ni1, tbase + B), ti1 = trio_get(tmask, rm.ti1, tbase + B);
                if (first_of[B] >= 0) {
                    ActiveContact& cc = ac[first_of[B]];
                    cc.p[0].normal_impulse = ni0; cc.p[0].tangent_impulse = ti0;
                    cc.p[1].normal_impulse = ni1; cc.p[1].tangent_impulse = ti1;
                }
            }
            for (int kk = 0; kk < nextra; ++kk) {   // (trio-uniform trip count)
                ActiveContact& cc = ac[(extra >> (4 * kk)) & 15u];
                const int src = tbase + cc.body;
                cc.p[0].normal_impulse = trio_get(tmask, cc.p[0].normal_impulse, src); cc.p[0].tangent_impulse = trio_get(tmask, cc.p[0].tangent_impulse, src);
                cc.p[1].normal_impulse = trio_get(tmask, cc.p[1].normal_impulse, src); cc.p[1].tangent_impulse = trio_get(tmask, cc.p[1].tangent_impulse, src);
            }
        } else {
        // A fuselage manifold means a crash, and under auto-reset a crashed lander is re-created before the solve (step_autoreset):
        // its island has no rows at all.  Warps in which no lane holds a fuselage row -- nearly all -- run the loop without it
        // (one instance of the loop per case: each stays one basic block), a fifth fewer instructions per iteration.
        auto velocity_iterations = [&](auto row0_tag) {
        constexpr int B0 = decltype(row0_tag)::value ? 0 : 1;
#pragma unroll 1
        for (int it = 0; it < VELOCITY_ITERATIONS; ++it) {
            joint_velocity<1>(Jl[1], jw[1], v[0], w[0], v[2], w[2]);   // island order: leg 1's joint, then leg 0's
            joint_velocity<0>(Jl[0], jw[0], v[0], w[0], v[1], w[1]);
            if (HAS_PAIRS && warp_rows) {
#pragma unroll
                for (int B = B0; B < 3; ++B) velocity_row_select(row[B], v[B], w[B]);
                if (warp_extra) {
                    for (int kk = 0; kk < nextra; ++kk) {
                        ActiveContact& cc = ac[(extra >> (4 * kk)) & 15u];
                        const int B = cc.body;
                        V2 vB = GETB(v, B); float wB = GETB(w, B);   // only velocities change in a velocity iteration
                        VelRow q;
                        q.normal = cc.normal; q.rb0 = cc.p[0].rb; q.rb1 = cc.p[1].rb;
                        q.ni0 = cc.p[0].normal_impulse; q.ni1 = cc.p[1].normal_impulse; q.ti0 = cc.p[0].tangent_impulse; q.ti1 = cc.p[1].tangent_impulse;
                        q.nm0 = cc.p[0].normal_mass; q.nm1 = cc.p[1].normal_mass; q.tm0 = cc.p[0].tangent_mass; q.tm1 = cc.p[1].tangent_mass;
                        q.friction = cc.friction;
                        q.k11 = cc.k11; q.k12 = cc.k12; q.k22 = cc.k22; q.i11 = cc.nm11; q.i12 = cc.nm12; q.i21 = cc.nm21; q.i22 = cc.nm22;
                        q.inv_mass = SHAPES[B].inv_mass; q.inv_inertia = SHAPES[B].inv_inertia;
                        q.count = cc.count; q.body = B;
                        velocity_row_select(q, vB, wB);
                        cc.p[0].normal_impulse = q.ni0; cc.p[0].tangent_impulse = q.ti0;
                        cc.p[1].normal_impulse = q.ni1; cc.p[1].tangent_impulse = q.ti1;
                        SETB(v, B, vB); SETB(w, B, wB);
                    }
                }
            }
        }
        };
        if (HAS_PAIRS && warp_row0) velocity_iterations(std::true_type{}); else velocity_iterations(std::false_type{});
        if (HAS_PAIRS) {
            int first_of[3] = {-1, -1, -1};
            for (int kk = nc - 1; kk >= 0; --kk) { const int k = (int)((order >> (4 * kk)) & 15u); SETB(first_of, ac[k].body, k); }
#pragma unroll
            for (int B = 0; B < 3; ++B)
                if (first_of[B] >= 0) {
                    ActiveContact& cc = ac[first_of[B]];
                    cc.p[0].normal_impulse = row[B].ni0; cc.p[0].tangent_impulse = row[B].ti0;
                    cc.p[1].normal_impulse = row[B].ni1; cc.p[1].tangent_impulse = row[B].ti1;
                }
        }
        }   // !TRIO

        LUNAR_PHASE(HAS_PAIRS, 4);
        // integrate positions
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const V2 translation = h * v[i];
            if (dot(translation, translation) > MAX_TRANSLATION * MAX_TRANSLATION) {
                const float ratio = MAX_TRANSLATION / sqrtf(dot(translation, translation));
                v[i] = ratio * v[i];
            }
            const float rotation = h * w[i];
            if (rotation * rotation > MAX_ROTATION * MAX_ROTATION) {
                const float ratio = MAX_ROTATION / fabsf(rotation);
                w[i] = w[i] * ratio;
            }
            c[i] = c[i] + h * v[i];
            a[i] = a[i] + h * w[i];
        }

        // ---- position iterations
        // A position iteration is a pure function of the nine coordinates (c, a) of the three bodies.  Box2D's exit
        // test never fires while a leg rests on its joint limit (the limit is corrected down to angular_error ~
        // ANGULAR_SLOP + 1 ulp, just above what joints_okay accepts), but after two or three iterations the
        // corrections round to nothing: once an iteration returns bit-identical coordinates, the remaining ones
        // would too, so the loop stops there -- same result as all 60, position_solved stays false.
        PosRow prow[(HAS_PAIRS && !TRIO) ? 3 : 1];
        if constexpr (TRIO) {
            PosRow& q = prow[0];
            q.count = 0; q.body = sub; q.type = MF_FACE_A;
            q.centroid = SHAPES[sub].centroid; q.inv_mass = SHAPES[sub].inv_mass; q.inv_inertia = SHAPES[sub].inv_inertia;
            q.local_normal = mk(0.0f, 1.0f); q.local_point = q.lp0 = q.lp1 = mk(0.0f, 0.0f);
            if (mine >= 0) {
                const Manifold& m = ac[mine].m;
                q.local_normal = m.local_normal; q.local_point = m.local_point; q.lp0 = m.lp[0]; q.lp1 = m.lp[1];
                q.type = m.type; q.count = m.count;
            }
        } else if (HAS_PAIRS) {
            int first_of[3] = {-1, -1, -1};
            for (int kk = nc - 1; kk >= 0; --kk) { const int k = (int)((order >> (4 * kk)) & 15u); SETB(first_of, ac[k].body, k); }
#pragma unroll
            for (int B = 0; B < 3; ++B) {
                PosRow& q = prow[B];
                q.count = 0; q.body = B; q.type = MF_FACE_A;
                q.centroid = SHAPES[B].centroid; q.inv_mass = SHAPES[B].inv_mass; q.inv_inertia = SHAPES[B].inv_inertia;
                q.local_normal = mk(0.0f, 1.0f); q.local_point = q.lp0 = q.lp1 = mk(0.0f, 0.0f);
                if (first_of[B] >= 0) {
                    const Manifold& m = ac[first_of[B]].m;
                    q.local_normal = m.local_normal; q.local_point = m.local_point; q.lp0 = m.lp[0]; q.lp1 = m.lp[1];
                    q.type = m.type; q.count = m.count;
                }
            }
        }
        // every body angle far inside the single-path range of the argument reduction (a step turns a body by at most pi / 2)?
        const bool small_angles = fabsf(a[0]) < 30000.0f && fabsf(a[1]) < 30000.0f && fabsf(a[2]) < 30000.0f;
        bool position_solved = false;
        auto position_iterations_trio = [&](auto in_range_tag) {
            constexpr bool IN_RANGE = decltype(in_range_tag)::value;
            V2 cm = GETB(c, sub); float am = GETB(a, sub);
            const PosRow& m = prow[0];
#pragma unroll 1
            for (int it = 0; it < POSITION_ITERATIONS; ++it) {
                const V2 c_in = cm; const float a_in = am;
                float min_separation = 0.0f;
                if (warp_rows) {
                    position_point_select<IN_RANGE>(m.count > 0, m.type, m.local_normal, m.local_point, m.lp0, m.centroid, m.inv_mass, m.inv_inertia, cm, am, min_separation);
                    position_point_select<IN_RANGE>(m.count > 1, m.type, m.local_normal, m.local_point, m.lp1, m.centroid, m.inv_mass, m.inv_inertia, cm, am, min_separation);
                    if (warp_extra) {
                        for (int kk = 0; kk < nextra_m; ++kk) {
                            const ActiveContact& cc = ac[(extra_m >> (4 * kk)) & 15u];
                            for (int j = 0; j < cc.m.count; ++j)
                                position_point(cc.m.type, cc.m.local_normal, cc.m.local_point, cc.m.lp[j], m.centroid, m.inv_mass, m.inv_inertia, cm, am, min_separation);
                        }
                    }
                }
                // min over the island's points >= -3 linearSlop  <=>  every lane's own minimum is
                const bool contacts_okay = __all_sync(tmask, min_separation >= -3.0f * LINEAR_SLOP) != 0;
                const bool j1 = joint_position_trio<1, IN_RANGE>(Jl[1], jw[1].motor_mass, tmask, tbase, sub, cm, am);
                const bool j0 = joint_position_trio<0, IN_RANGE>(Jl[0], jw[0].motor_mass, tmask, tbase, sub, cm, am);
                if (contacts_okay && j1 && j0) { position_solved = true; break; }
                const int moved = (__float_as_int(cm.x) ^ __float_as_int(c_in.x)) | (__float_as_int(cm.y) ^ __float_as_int(c_in.y)) | (__float_as_int(am) ^ __float_as_int(a_in));
                if (__ballot_sync(tmask, moved != 0) == 0u) break;   // fixed point of the whole island
            }
#pragma unroll
            for (int B = 0; B < 3; ++B) { c[B] = trio_get(tmask, cm, tbase + B); a[B] = trio_get(tmask, am, tbase + B); }
        };
        auto position_iterations = [&](auto in_range_tag, auto row0_tag) {
            constexpr bool IN_RANGE = decltype(in_range_tag)::value;
            constexpr int B0 = decltype(row0_tag)::value ? 0 : 1;
#pragma unroll 1
            for (int it = 0; it < POSITION_ITERATIONS; ++it) {
                const V2 c_in[3] = {c[0], c[1], c[2]};
                const float a_in[3] = {a[0], a[1], a[2]};
                float min_separation = 0.0f;
                if (HAS_PAIRS && warp_rows) {
#pragma unroll
                    for (int B = B0; B < 3; ++B) {
                        const PosRow& m = prow[B];
                        position_point_select<IN_RANGE>(m.count > 0, m.type, m.local_normal, m.local_point, m.lp0, m.centroid, m.inv_mass, m.inv_inertia, c[B], a[B], min_separation);
                        position_point_select<IN_RANGE>(m.count > 1, m.type, m.local_normal, m.local_point, m.lp1, m.centroid, m.inv_mass, m.inv_inertia, c[B], a[B], min_separation);
                    }
                    if (warp_extra) {
                        for (int kk = 0; kk < nextra; ++kk) {
                            const ActiveContact& cc = ac[(extra >> (4 * kk)) & 15u];
                            const int B = cc.body;
                            V2 cB = GETB(c, B); float aB = GETB(a, B);   // only positions change in a position iteration
                            for (int j = 0; j < cc.m.count; ++j)
                                position_point(cc.m.type, cc.m.local_normal, cc.m.local_point, cc.m.lp[j], SHAPES[B].centroid, SHAPES[B].inv_mass, SHAPES[B].inv_inertia, cB, aB, min_separation);
                            SETB(c, B, cB); SETB(a, B, aB);
                        }
                    }
                }
                const bool contacts_okay = min_separation >= -3.0f * LINEAR_SLOP;
                const bool j1 = joint_position<1, IN_RANGE>(Jl[1], jw[1].motor_mass, c[0], a[0], c[2], a[2]);
                const bool j0 = joint_position<0, IN_RANGE>(Jl[0], jw[0].motor_mass, c[0], a[0], c[1], a[1]);
                if (contacts_okay && j1 && j0) { position_solved = true; break; }
                int moved = 0;
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    moved |= (__float_as_int(c[i].x) ^ __float_as_int(c_in[i].x)) | (__float_as_int(c[i].y) ^ __float_as_int(c_in[i].y)) |
                             (__float_as_int(a[i]) ^ __float_as_int(a_in[i]));
                if (moved == 0) break;   // fixed point
            }
        };
        if constexpr (TRIO) { if (small_angles) position_iterations_trio(std::true_type{}); else position_iterations_trio(std::false_type{}); }
        else if (HAS_PAIRS && warp_row0) { if (small_angles) position_iterations(std::true_type{}, std::true_type{}); else position_iterations(std::false_type{}, std::true_type{}); }
        else { if (small_angles) position_iterations(std::true_type{}, std::false_type{}); else position_iterations(std::false_type{}, std::false_type{}); }
        LUNAR_PHASE(HAS_PAIRS, 5);

        // copy back, store impulses (b2ContactSolver::StoreImpulses): slots in creation order
#pragma unroll
        for (int i = 0; i < 3; ++i) { L.b[i].c = c[i]; L.b[i].a = a[i]; L.b[i].v = v[i]; L.b[i].w = w[i]; }
        L.j[0] = Jl[0]; L.j[1] = Jl[1];
        if (HAS_PAIRS) {
            for (int s = 0; s < MAXC; ++s) {
                ContactSlot& cs = L.c[s];
                if (s < nc) {
                    cs.pair = ac[s].body * 16 + ac[s].edge;
                    for (int k = 0; k < 2; ++k) {
                        const bool live = k < ac[s].m.count;
                        cs.key[k] = live ? ac[s].m.key[k] : NO_KEY;
                        cs.ni[k] = live ? ac[s].p[k].normal_impulse : 0.0f;
                        cs.ti[k] = live ? ac[s].p[k].tangent_impulse : 0.0f;
                    }
                } else {
                    cs.pair = -1; cs.key[0] = cs.key[1] = NO_KEY; cs.ni[0] = cs.ni[1] = cs.ti[0] = cs.ti[1] = 0.0f;
                }
            }
        }

        // sleeping (b2Island::Solve tail)
        float min_sleep = 3.4028234663852886e38f;
        const float lin_tol_sqr = LINEAR_SLEEP_TOL * LINEAR_SLEEP_TOL, ang_tol_sqr = ANGULAR_SLEEP_TOL * ANGULAR_SLEEP_TOL;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (L.b[i].w * L.b[i].w > ang_tol_sqr || dot(L.b[i].v, L.b[i].v) > lin_tol_sqr) {
                L.b[i].sleep_time = 0.0f;
                min_sleep = 0.0f;
            } else {
                L.b[i].sleep_time = L.b[i].sleep_time + h;
                min_sleep = minf(min_sleep, L.b[i].sleep_time);
            }
        }
        if (min_sleep >= TIME_TO_SLEEP && position_solved) set_awake(L, false);

        // b2Body::SynchronizeFixtures: the proxy box must hold the tight boxes of the start and end pose of the step; when it
        // does not, it is replaced by their union grown by aabbExtension and stretched along twice the displacement
        uint32_t moved = 0u;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const Rot q1 = rot(a0[i]), q2 = rot(a[i]);
            const V2 p1 = c0[i] - rmul(q1, SHAPES[i].centroid), p2 = c[i] - rmul(q2, SHAPES[i].centroid);
            const Box b1 = body_box(SHAPES[i], p1, q1), b2 = body_box(SHAPES[i], p2, q2);
            const Box u = Box{minf(b1.lx, b2.lx), minf(b1.ly, b2.ly), maxf(b1.hx, b2.hx), maxf(b1.hy, b2.hy)};
            const Box f = fat_of(L, i);
            if (f.lx <= u.lx && f.ly <= u.ly && u.hx <= f.hx && u.hy <= f.hy) continue;
            Box n = Box{u.lx - AABB_EXTENSION, u.ly - AABB_EXTENSION, u.hx + AABB_EXTENSION, u.hy + AABB_EXTENSION};
            const V2 d = AABB_MULTIPLIER * (p2 - p1);
            if (d.x < 0.0f) n.lx = n.lx + d.x; else n.hx = n.hx + d.x;
            if (d.y < 0.0f) n.ly = n.ly + d.y; else n.hy = n.hy + d.y;
            L.fat[i][0] = n.lx; L.fat[i][1] = n.ly; L.fat[i][2] = n.hx; L.fat[i][3] = n.hy;
            moved |= 1u << i;
        }
        // ContactManager.FindNewContacts: pairs of a moved proxy, sorted by proxy id (body, then edge in creation order: the
        // base edge was created first, :541, then the ten terrain edges), appended unless they exist already
        if (moved) {
            int np = pair_count(L);
            for (int i = 0; i < 3; ++i) {
                if (!((moved >> i) & 1u)) continue;
                const Box f = fat_of(L, i);
                for (int ee = 0; ee < NUM_EDGES; ++ee) {
                    const int e = ee == 0 ? BASE_EDGE : ee - 1;
                    if (!boxes_overlap(f, edge_fat_box(L, e))) continue;
                    const uint32_t pr = (uint32_t)(i * 16 + e);
                    bool exists = false;
                    for (int k = 0; k < np; ++k) if (pair_at(L, k) == pr) exists = true;
                    if (!exists && np < MAXP) pair_put(L, np++, pr);
                }
            }
        }
    }
    // ClearForces
    L.force = mk(0.0f, 0.0f);
    L.torque = 0.0f;
    L.flags &= ~F_FIRST_STEP;
}

// ---------------------------------------------------------------- Body.ApplyLinearImpulse(impulse, point)
__device__ __forceinline__ void apply_linear_impulse(Lander& L, V2 impulse, V2 point) {
    set_awake(L, true);
    Body& f = L.b[0];
    f.v = f.v + SHAPES[0].inv_mass * impulse;
    f.w = f.w + SHAPES[0].inv_inertia * cross(point - f.c, impulse);
}

__device__ __forceinline__ void observe(const Lander& L, float* o) {   // LunarLanderEnv.cs:733-747
    const Rot q1 = rot(L.b[0].a);
    const V2 p1 = L.b[0].c - rmul(q1, SHAPES[0].centroid);
    const float helipad_y = VIEW_H / 4.0f;                                                    // :517
    o[0] = (p1.x - VIEW_W / 2.0f) / (VIEW_W / 2.0f);
    o[1] = (p1.y - (helipad_y + LEG_DOWN / SCALE)) / (VIEW_H / 2.0f);
    o[2] = L.b[0].v.x * ((VIEW_W / 2.0f) / 50.0f);
    o[3] = L.b[0].v.y * ((VIEW_H / 2.0f) / 50.0f);
    o[4] = L.b[0].a;
    o[5] = 20.0f * L.b[0].w / 50.0f;
    o[6] = (L.flags & F_LEG0) ? 1.0f : 0.0f;
    o[7] = (L.flags & F_LEG1) ? 1.0f : 0.0f;
}

struct StepResult { float reward; uint8_t done; uint8_t did_reset; };
struct Powers { float m_power, s_power; };

// LunarLanderEnv.Step up to the physics (:574-716): wind, engine impulses.  `t` indexes the DYNAMICS stream (the two dispersion draws, :611-612).
__device__ __forceinline__ Powers pre_physics(Lander& L, uint64_t seed, uint32_t gid, uint64_t t, int i_action, const float* c_action) {
    const bool continuous = (L.flags & F_CONTINUOUS) != 0;
    float a0 = 0.0f, a1 = 0.0f;
    if (continuous) { a0 = clampf(c_action[0], -1.0f, 1.0f); a1 = clampf(c_action[1], -1.0f, 1.0f); }   // :600
    // wind (:588-596): only while no leg touches the ground
    if (L.use_wind && !(L.flags & (F_LEG0 | F_LEG1))) {
        const float wind_mag = (float)(tanh(sin(0.02 * L.wind_idx) + sin(3.14159265358979323846 * 0.01 * L.wind_idx))) * L.wind_power;
        L.wind_idx += 1;
        set_awake(L, true);
        {   // Body.ApplyForce(Vector2) applies the force at the body ORIGIN (Farseer lineage: ApplyForce(ref force, ref _xf.p)), so it also makes a torque about the centre
            const Rot qw = rot(L.b[0].a);
            const V2 origin = L.b[0].c - rmul(qw, SHAPES[0].centroid);
            L.force = L.force + mk(wind_mag, 0.0f);
            L.torque = L.torque + ((origin.x - L.b[0].c.x) * 0.0f - (origin.y - L.b[0].c.y) * wind_mag);
        }
        const float torque_mag = (float)(tanh(sin(0.02 * L.torque_idx) + sin(3.14159265358979323846 * 0.01 * L.torque_idx))) * L.turbulence_power;
        L.torque_idx += 1;
        L.torque = L.torque + torque_mag;
    }
    // fuselage pose: Body.Position is the body origin, Rotation the angle (:609, :657)
    const Rot qf = rot(L.b[0].a);
    const V2 pos0 = L.b[0].c - rmul(qf, SHAPES[0].centroid);
    const V2 tip = mk(qf.s, qf.c);                   // :609
    const V2 side = mk(-tip.y, tip.x);               // :610
    const Block d = draw(seed, gid, t, STREAM_DYNAMICS);
    const float disp_x = uniformf(-1.0f, 1.0f, d.w0) / SCALE;   // :611
    const float disp_y = uniformf(-1.0f, 1.0f, d.w1) / SCALE;   // :612
    bool fire_main, fire_thruster;
    if (continuous) { fire_main = a0 > 0.0f; fire_thruster = fabsf(a1) > 0.5f; }          // :618-621
    else { fire_main = i_action == 2; fire_thruster = i_action == 1 || i_action == 3; }      // :629-630
    float m_power = 0.0f;
    if (fire_main) {
        m_power = continuous ? (clampf(a0, 0.0f, 1.0f) + 1.0f) * 0.5f : 1.0f;                 // :640, :652
        const float ox = tip.x * (4.0f / SCALE + 2.0f * disp_x) + side.x * disp_y;            // :655
        const float oy = -tip.y * (4.0f / SCALE + 2.0f * disp_x) - side.y * disp_y;           // :656
        const V2 impulse_pos = pos0 + mk(ox, oy);                                             // :657
        const V2 impulse = mk(-ox * MAIN_ENGINE_POWER * m_power, -oy * MAIN_ENGINE_POWER * m_power);   // :665
        apply_linear_impulse(L, impulse, impulse_pos);                                        // :670
    }
    float s_power = 0.0f;
    if (fire_thruster) {
        float direction;
        if (continuous) { direction = a1 < 0.0f ? -1.0f : 1.0f; s_power = clampf(fabsf(a1), 0.5f, 1.0f); }   // :683-684
        else { direction = (float)i_action - 2.0f; s_power = 1.0f; }                          // :697-698
        const float ox = tip.x * disp_x + side.x * (3.0f * disp_y + direction * SIDE_ENGINE_AWAY / SCALE);    // :700
        const float oy = -tip.y * disp_x - side.y * (3.0f * disp_y + direction * SIDE_ENGINE_AWAY / SCALE);   // :701
        const V2 impulse_pos = pos0 + mk(ox - tip.x * 17.0f / SCALE, oy + tip.y * SIDE_ENGINE_HEIGHT / SCALE);   // :702
        const V2 impulse = mk(-ox * SIDE_ENGINE_POWER * s_power, -oy * SIDE_ENGINE_POWER * s_power);   // :709
        apply_linear_impulse(L, impulse, impulse_pos);                                        // :714
    }
    return Powers{m_power, s_power};
}

// LunarLanderEnv.Step after the physics (:726-772): GameOver, observation, shaping reward, termination
__device__ __forceinline__ StepResult post_physics(Lander& L, Powers pw) {
    if (L.flags & F_FUSELAGE) L.flags |= F_GAME_OVER;                                         // :726-729
    observe(L, L.obs);                                                                        // :733-747
    const float px = L.obs[0], py = L.obs[1], vx = L.obs[2], vy = L.obs[3], angle = L.obs[4];
    const float l0 = L.obs[6], l1 = L.obs[7];
    // reward (:748-760)
    float reward = 0.0f;
    float shaping = -100.0f * sqrtf(px * px + py * py);
    shaping += -100.0f * sqrtf(vx * vx + vy * vy);
    shaping += -100.0f * fabsf(angle);
    shaping += 10.0f * l0;
    shaping += 10.0f * l1;
    if (L.prev_shaping != -3.4028234663852886e38f) reward = shaping - L.prev_shaping;          // float.MinValue sentinel
    L.prev_shaping = shaping;
    reward -= pw.m_power * 0.3f;
    reward -= pw.s_power * 0.03f;
    uint8_t done = 0;
    if ((L.flags & F_GAME_OVER) || px > 1.0f) { done = 1; reward = -100.0f; }                 // :762 one-sided: no `< -1` check
    if (!(L.flags & F_AWAKE)) { done = 1; reward = 100.0f; }                                  // :767
    return StepResult{reward, done, 0};
}

// LunarLanderEnv.Step (:574-774)
template <bool HAS_PAIRS = true>
__device__ __noinline__ StepResult step(Lander& L, uint64_t seed, uint32_t gid, uint64_t t, int i_action, const float* c_action) {
    const Powers pw = pre_physics(L, seed, gid, t, i_action, c_action);
    Contacts C;
    collide<HAS_PAIRS>(L, C);
    solve<HAS_PAIRS>(L, C);                                                                   // :721-725
    return post_physics(L, pw);
}

// LunarLanderEnv.Reset (:489-572) up to its zero step.  `index` = episode ordinal; draws: sub-block 0 = (fx, fy, h0, h1),
// 1 = (h2..h5), 2 = (h6..h9), 3 = (h10, h11).
__device__ __forceinline__ void reset_prepare(Lander& L, uint64_t seed, uint32_t gid, uint64_t index, bool continuous,
                  float gravity, int use_wind, float wind_power, float turbulence_power) {
    const int32_t wi = L.wind_idx, ti = L.torque_idx;   // drawn in the constructor, persist across episodes (:409-410)
    zero_lander(L);
    L.wind_idx = wi; L.torque_idx = ti;
    L.gravity = gravity; L.use_wind = use_wind; L.wind_power = wind_power; L.turbulence_power = turbulence_power;
    L.flags = F_AWAKE | F_FIRST_STEP | (continuous ? F_CONTINUOUS : 0);
    for (int s = 0; s < MAXC; ++s) { L.c[s].pair = -1; L.c[s].key[0] = L.c[s].key[1] = NO_KEY; }
    const Block b0 = draw(seed, gid, index, STREAM_RESET, 0), b1 = draw(seed, gid, index, STREAM_RESET, 1),
                b2 = draw(seed, gid, index, STREAM_RESET, 2), b3 = draw(seed, gid, index, STREAM_RESET, 3);
    L.force = mk(uniformf(-INITIAL_RANDOM, INITIAL_RANDOM, b0.w0), uniformf(-INITIAL_RANDOM, INITIAL_RANDOM, b0.w1));   // :496
    {   // ... applied at the body origin while the fuselage still sits at (0, 0), rotation 0 (:496 precedes :561): torque about the centre of mass
        const V2 c_at_creation = rmul(rot(0.0f), SHAPES[0].centroid) + mk(0.0f, 0.0f);
        L.torque = (0.0f - c_at_creation.x) * L.force.y - (0.0f - c_at_creation.y) * L.force.x;
    }
    L.prev_shaping = -3.4028234663852886e38f;                                                  // :498
    float height[CHUNKS + 1];
    const uint32_t hw[12] = {b0.w2, b0.w3, b1.w0, b1.w1, b1.w2, b1.w3, b2.w0, b2.w1, b2.w2, b2.w3, b3.w0, b3.w1};
    for (int i = 0; i < CHUNKS + 1; ++i) height[i] = uniformf(0.0f, VIEW_H / 2.0f, hw[i]);   // :507
    const int mid = CHUNKS / 2;
    const float helipad_y = VIEW_H / 4.0f;
    for (int i = mid - 2; i <= mid + 2; ++i) height[i] = helipad_y;                             // :518-522
    for (int i = 0; i < CHUNKS; ++i) {
        const float h1 = i > 0 ? height[i - 1] : 0.0f;                                         // :526-530
        float y = 0.33f * (h1 + height[i] + height[i + 1]);                                    // :531
        if (y > VIEW_H) y = VIEW_H / 4.0f;                                                     // :532-535
        L.terrain[i] = y;
    }
    // bodies: created at the origin / (+-LEG_AWAY/S, 0) with rotation -+0.05 (:259-261), then moved by (W/2, H) (:561-565)
    const V2 dv = mk(VIEW_W / 2.0f, VIEW_H);
    for (int i = 0; i < 3; ++i) {
        const float ang = i == 0 ? 0.0f : (i == 1 ? -0.05f : 0.05f);
        const V2 origin = i == 0 ? mk(0.0f, 0.0f) : mk((i == 1 ? -1.0f : 1.0f) * LEG_AWAY / SCALE, 0.0f);
        const V2 pos = origin + dv;
        L.b[i].a = ang;
        L.b[i].c = rmul(rot(ang), SHAPES[i].centroid) + pos;
        L.b[i].v = mk(0.0f, 0.0f); L.b[i].w = 0.0f; L.b[i].sleep_time = 0.0f;
        const Box bx = body_box(SHAPES[i], pos, rot(ang));   // Body.Position setter -> MoveProxy with no displacement
        L.fat[i][0] = bx.lx - AABB_EXTENSION; L.fat[i][1] = bx.ly - AABB_EXTENSION; L.fat[i][2] = bx.hx + AABB_EXTENSION; L.fat[i][3] = bx.hy + AABB_EXTENSION;
    }
}

// LunarLanderEnv.Reset (:489-572): ends with the zero step (:567-571) at step index `t`
__device__ __forceinline__ void reset(Lander& L, uint64_t seed, uint32_t gid, uint64_t index, bool continuous, uint64_t t,
                  float gravity, int use_wind, float wind_power, float turbulence_power) {
    reset_prepare(L, seed, gid, index, continuous, gravity, use_wind, wind_power, turbulence_power);
    const float zero[2] = {0.0f, 0.0f};
    step<false>(L, seed, gid, t, 0, zero);                                                     // :567-571 (a new world: no contact exists yet)
}

// Step with the in-kernel auto-reset folded in.  A lander whose fuselage touched the ground in this step's Collide is over
// (GameOver, :726-729: reward -100, done) whatever its solve would give, and under auto-reset its state is replaced by a new
// episode whose first act is a whole World.Step of its own (the zero step of Reset, :567-571).  Both solves would run one after
// the other in the slowest lane of the warp; instead the lander is re-created right after Collide and THIS step's solve is the
// new episode's zero step: same outputs (reward -100, done, the observation after the zero step) and same state, one solve
// instead of two.  Not taken when the old solve could still matter: the island could fall asleep in this very step (done with
// +100 overrides -100, :767), or the caller wants the terminal observation (`allow` false).
template <bool HAS_PAIRS = true, bool TRIO = false>
__device__ __noinline__ StepResult step_autoreset(Lander& L, uint64_t seed, uint32_t gid, uint64_t t, int i_action, const float* c_action,
                                                  bool allow, uint64_t next_ordinal) {
    LUNAR_PHASE_NS(HAS_PAIRS, 8);
    LUNAR_PHASE(HAS_PAIRS, 0);
    Powers pw = pre_physics(L, seed, gid, t, i_action, c_action);
    LUNAR_PHASE(HAS_PAIRS, 1);
    Contacts C;
    collide<HAS_PAIRS>(L, C);
    LUNAR_PHASE(HAS_PAIRS, 2);
    bool early = false;
    if (HAS_PAIRS && allow && (L.flags & (F_FUSELAGE | F_GAME_OVER)) && (L.flags & F_AWAKE)) {
        bool may_sleep = true;   // the island sleeps only if EVERY body's timer reaches timeToSleep: one that cannot, after this step, rules it out
#pragma unroll
        for (int i = 0; i < 3; ++i) if (!(L.b[i].sleep_time + DT >= TIME_TO_SLEEP)) may_sleep = false;
        if (!may_sleep) {
            const bool continuous = (L.flags & F_CONTINUOUS) != 0;
            reset_prepare(L, seed, gid, next_ordinal, continuous, L.gravity, L.use_wind, L.wind_power, L.turbulence_power);
            const float zero[2] = {0.0f, 0.0f};
            pw = pre_physics(L, seed, gid, t + 1, 0, zero);   // the zero step of the new episode (:567-571)
            C.nc = 0; C.order = 0u;
            early = true;
        }
    }
    solve<HAS_PAIRS, TRIO>(L, C);                                                             // :721-725
    LUNAR_PHASE(HAS_PAIRS, 6);
    StepResult r = post_physics(L, pw);
    if (early) r = StepResult{-100.0f, 1, 1};
    LUNAR_PHASE(HAS_PAIRS, 7);
    LUNAR_PHASE_NS(HAS_PAIRS, 9);
    return r;
}

}}  // namespace gymcuda::lunar

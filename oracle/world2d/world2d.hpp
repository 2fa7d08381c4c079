// TEST INFRASTRUCTURE ONLY.  Nothing under gym.net_b200/ may include, link or load this.
//
// world2d -- a small GENERIC 2-D rigid-body engine in the Box2D-2.3 / Farseer-3.5 lineage that
// Aether.Physics2D 1.6.1 descends from (the un-vendored NuGet dependency behind LunarLanderEnv's
// World.Step, src/Gym.Environments/Gym.Environments.csproj:33; call site
// src/Gym.Environments/Envs/Aether/LunarLanderEnv.cs:721-725).
//
// Purpose: an oracle for LunarLander that is NOT the kernel's text.  gym.net_b200/csrc/lunar_core.cuh
// (and its twin oracle/lunar.hpp) are a fixed-topology specialisation: three hard-coded bodies, two
// hard-coded joints, polygon-vs-static-edge only, one-body contact formulas, precomputed mass data.
// This file knows nothing about landers.  It has
//   * bodies / fixtures / shapes created at run time, convex hulls and mass data computed from vertices,
//   * a broad phase over fat AABBs with a move buffer and sorted pair creation,
//   * contacts as objects in per-body linked lists (head insertion) with begin/end callbacks,
//   * the general two-body sequential-impulse contact solver (block solver, warm starting),
//   * a general revolute joint (motor + limit), DFS island construction, island sleeping,
//   * optionally the continuous (time-of-impact) pass of b2World::SolveTOI.
// oracle/world2d/lunar_sim.cpp builds LunarLanderEnv's world on top of it exactly as the C# does
// (CreateBody / CreateFixture / RevoluteJoint / ContactManager.BeginContact ...).
//
// Arithmetic: float32, one IEEE operation per source operation (-ffp-contract=off), in the operation
// order of the published Box2D 2.3.0 sources.  sin/cos of a body angle go through `World::sincos`
// (default: (float)sin((double)a), what Aether's Complex/Rot constructors compute with Math.Sin/Cos).
//
// Where the lineage leaves a choice open (Aether 1.6.1 is not on this machine) the choice is an
// explicit, documented option in `WorldOptions` -- see oracle/world2d/README.md.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <vector>

namespace w2d {

// ------------------------------------------------------------------------------------------------
// settings (Farseer Settings.cs / b2Settings.h)
// ------------------------------------------------------------------------------------------------
constexpr float kPi = 3.14159265359f;
constexpr float kEpsilon = FLT_EPSILON;
constexpr float kMaxFloat = FLT_MAX;
constexpr int kMaxManifoldPoints = 2;
constexpr int kMaxPolygonVertices = 8;
constexpr float kAabbExtension = 0.1f;
constexpr float kAabbMultiplier = 2.0f;
constexpr float kLinearSlop = 0.005f;
constexpr float kAngularSlop = 2.0f / 180.0f * kPi;
constexpr float kPolygonRadius = 2.0f * kLinearSlop;
constexpr int kMaxSubSteps = 8;
constexpr int kMaxTOIContacts = 32;
constexpr float kVelocityThreshold = 1.0f;
constexpr float kMaxLinearCorrection = 0.2f;
constexpr float kMaxAngularCorrection = 8.0f / 180.0f * kPi;
constexpr float kMaxTranslation = 2.0f;
constexpr float kMaxTranslationSquared = kMaxTranslation * kMaxTranslation;
constexpr float kMaxRotation = 0.5f * kPi;
constexpr float kMaxRotationSquared = kMaxRotation * kMaxRotation;
constexpr float kBaumgarte = 0.2f;
constexpr float kToiBaumgarte = 0.75f;
constexpr float kTimeToSleep = 0.5f;
constexpr float kLinearSleepTolerance = 0.01f;
constexpr float kAngularSleepTolerance = 2.0f / 180.0f * kPi;
constexpr float kDefaultFriction = 0.2f;      // Farseer: new Fixture() -> Friction = 0.2f
constexpr float kDefaultRestitution = 0.0f;

// ------------------------------------------------------------------------------------------------
// math
// ------------------------------------------------------------------------------------------------
struct Vec2 {
    float x = 0.0f, y = 0.0f;
    Vec2() = default;
    Vec2(float x_, float y_) : x(x_), y(y_) {}
    Vec2 operator-() const { return Vec2(-x, -y); }
    void operator+=(const Vec2& o) { x += o.x; y += o.y; }
    void operator-=(const Vec2& o) { x -= o.x; y -= o.y; }
    void operator*=(float s) { x *= s; y *= s; }
    float lengthSquared() const { return x * x + y * y; }
    float length() const { return std::sqrt(x * x + y * y); }
    float normalize() {
        const float len = length();
        if (len < kEpsilon) return 0.0f;
        const float inv = 1.0f / len;
        x *= inv; y *= inv;
        return len;
    }
};
inline Vec2 operator+(const Vec2& a, const Vec2& b) { return Vec2(a.x + b.x, a.y + b.y); }
inline Vec2 operator-(const Vec2& a, const Vec2& b) { return Vec2(a.x - b.x, a.y - b.y); }
inline Vec2 operator*(float s, const Vec2& a) { return Vec2(s * a.x, s * a.y); }
inline float dot(const Vec2& a, const Vec2& b) { return a.x * b.x + a.y * b.y; }
inline float cross(const Vec2& a, const Vec2& b) { return a.x * b.y - a.y * b.x; }
inline Vec2 cross(const Vec2& a, float s) { return Vec2(s * a.y, -s * a.x); }
inline Vec2 cross(float s, const Vec2& a) { return Vec2(-s * a.y, s * a.x); }
inline float clampf(float a, float lo, float hi) { return std::max(lo, std::min(a, hi)); }
inline Vec2 vmin(const Vec2& a, const Vec2& b) { return Vec2(std::min(a.x, b.x), std::min(a.y, b.y)); }
inline Vec2 vmax(const Vec2& a, const Vec2& b) { return Vec2(std::max(a.x, b.x), std::max(a.y, b.y)); }

struct Vec3 {
    float x = 0.0f, y = 0.0f, z = 0.0f;
    Vec3() = default;
    Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    Vec3 operator-() const { return Vec3(-x, -y, -z); }
    void operator+=(const Vec3& o) { x += o.x; y += o.y; z += o.z; }
    void operator*=(float s) { x *= s; y *= s; z *= s; }
};
inline float dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(const Vec3& a, const Vec3& b) { return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

struct Mat22 {
    Vec2 ex, ey;
    Mat22 inverse() const {
        const float a = ex.x, b = ey.x, c = ex.y, d = ey.y;
        float det = a * d - b * c;
        if (det != 0.0f) det = 1.0f / det;
        Mat22 B;
        B.ex.x = det * d; B.ey.x = -det * b;
        B.ex.y = -det * c; B.ey.y = det * a;
        return B;
    }
    Vec2 solve(const Vec2& b) const {
        const float a11 = ex.x, a12 = ey.x, a21 = ex.y, a22 = ey.y;
        float det = a11 * a22 - a12 * a21;
        if (det != 0.0f) det = 1.0f / det;
        return Vec2(det * (a22 * b.x - a12 * b.y), det * (a11 * b.y - a21 * b.x));
    }
};
inline Vec2 mul(const Mat22& A, const Vec2& v) { return Vec2(A.ex.x * v.x + A.ey.x * v.y, A.ex.y * v.x + A.ey.y * v.y); }

struct Mat33 {
    Vec3 ex, ey, ez;
    Vec3 solve33(const Vec3& b) const {
        float det = dot(ex, cross(ey, ez));
        if (det != 0.0f) det = 1.0f / det;
        Vec3 x;
        x.x = det * dot(b, cross(ey, ez));
        x.y = det * dot(ex, cross(b, ez));
        x.z = det * dot(ex, cross(ey, b));
        return x;
    }
    Vec2 solve22(const Vec2& b) const {
        const float a11 = ex.x, a12 = ey.x, a21 = ex.y, a22 = ey.y;
        float det = a11 * a22 - a12 * a21;
        if (det != 0.0f) det = 1.0f / det;
        return Vec2(det * (a22 * b.x - a12 * b.y), det * (a11 * b.y - a21 * b.x));
    }
};

// sin/cos provider of the running world (set for the duration of World::step and of the calls that build bodies)
using SinCosFn = void (*)(float, float*, float*);
inline void sincos_libm_double(float a, float* s, float* c) { *s = (float)std::sin((double)a); *c = (float)std::cos((double)a); }
inline SinCosFn& current_sincos() { static thread_local SinCosFn f = sincos_libm_double; return f; }

struct Rot {
    float s = 0.0f, c = 1.0f;
    Rot() = default;
    explicit Rot(float angle) { set(angle); }
    void set(float angle) { current_sincos()(angle, &s, &c); }
};
inline Vec2 mul(const Rot& q, const Vec2& v) { return Vec2(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
inline Vec2 mulT(const Rot& q, const Vec2& v) { return Vec2(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
inline Rot mulT(const Rot& q, const Rot& r) { Rot o; o.s = q.c * r.s - q.s * r.c; o.c = q.c * r.c + q.s * r.s; return o; }

struct Transform { Vec2 p; Rot q; };
inline Vec2 mul(const Transform& T, const Vec2& v) { return Vec2((T.q.c * v.x - T.q.s * v.y) + T.p.x, (T.q.s * v.x + T.q.c * v.y) + T.p.y); }
inline Vec2 mulT(const Transform& T, const Vec2& v) {
    const float px = v.x - T.p.x, py = v.y - T.p.y;
    return Vec2(T.q.c * px + T.q.s * py, -T.q.s * px + T.q.c * py);
}
inline Transform mulT(const Transform& A, const Transform& B) { Transform C; C.q = mulT(A.q, B.q); C.p = mulT(A.q, B.p - A.p); return C; }

struct Sweep {
    Vec2 localCenter, c0, c;
    float a0 = 0.0f, a = 0.0f, alpha0 = 0.0f;
    void getTransform(Transform* xf, float beta) const {
        xf->p = (1.0f - beta) * c0 + beta * c;
        const float angle = (1.0f - beta) * a0 + beta * a;
        xf->q.set(angle);
        xf->p -= mul(xf->q, localCenter);
    }
    void advance(float alpha) {
        const float beta = (alpha - alpha0) / (1.0f - alpha0);
        c0 += beta * (c - c0);
        a0 += beta * (a - a0);
        alpha0 = alpha;
    }
};

struct AABB {
    Vec2 lo, hi;
    bool contains(const AABB& o) const { return lo.x <= o.lo.x && lo.y <= o.lo.y && o.hi.x <= hi.x && o.hi.y <= hi.y; }
    void combine(const AABB& a, const AABB& b) { lo = vmin(a.lo, b.lo); hi = vmax(a.hi, b.hi); }
};
inline bool overlap(const AABB& a, const AABB& b) {
    const Vec2 d1 = b.lo - a.hi, d2 = a.lo - b.hi;
    if (d1.x > 0.0f || d1.y > 0.0f) return false;
    if (d2.x > 0.0f || d2.y > 0.0f) return false;
    return true;
}

// ------------------------------------------------------------------------------------------------
// shapes
// ------------------------------------------------------------------------------------------------
struct MassData { float mass = 0.0f; Vec2 center; float inertia = 0.0f; };   // inertia about the shape's origin

struct Shape {
    enum Type { EDGE = 1, POLYGON = 2 };
    Type type;
    float radius;
    float density = 0.0f;
    explicit Shape(Type t, float r) : type(t), radius(r) {}
    virtual ~Shape() = default;
    virtual AABB computeAABB(const Transform& xf) const = 0;
    virtual MassData massData() const = 0;
};

struct EdgeShape : Shape {
    Vec2 v0, v1, v2, v3;
    bool hasVertex0 = false, hasVertex3 = false;
    EdgeShape(Vec2 a, Vec2 b) : Shape(EDGE, kPolygonRadius), v1(a), v2(b) {}
    AABB computeAABB(const Transform& xf) const override {
        const Vec2 a = mul(xf, v1), b = mul(xf, v2);
        const Vec2 lo = vmin(a, b), hi = vmax(a, b), r(radius, radius);
        return AABB{lo - r, hi + r};
    }
    MassData massData() const override { MassData m; m.center = 0.5f * (v1 + v2); return m; }
};

struct PolygonShape : Shape {
    std::vector<Vec2> vertices, normals;
    MassData md;
    // PolygonShape(Vertices, density): convex hull by gift wrapping from the right-most (then lowest) vertex,
    // counter-clockwise; edge normals; mass data with the vertex average as the triangle-fan apex.
    PolygonShape(const std::vector<Vec2>& pts, float density_) : Shape(POLYGON, kPolygonRadius) {
        density = density_;
        const int n = (int)pts.size();
        int i0 = 0;
        float x0 = pts[0].x;
        for (int i = 1; i < n; ++i) {
            const float x = pts[i].x;
            if (x > x0 || (x == x0 && pts[i].y < pts[i0].y)) { i0 = i; x0 = x; }
        }
        std::vector<int> hull;
        int ih = i0;
        for (;;) {
            hull.push_back(ih);
            int ie = 0;
            for (int j = 1; j < n; ++j) {
                if (ie == ih) { ie = j; continue; }
                const Vec2 r = pts[ie] - pts[hull.back()];
                const Vec2 v = pts[j] - pts[hull.back()];
                const float c = cross(r, v);
                if (c < 0.0f) ie = j;
                if (c == 0.0f && v.lengthSquared() > r.lengthSquared()) ie = j;   // collinear: keep the farther one
            }
            ih = ie;
            if (ie == i0) break;
        }
        for (int h : hull) vertices.push_back(pts[h]);
        const int m = (int)vertices.size();
        for (int i = 0; i < m; ++i) {
            const Vec2 edge = vertices[i + 1 < m ? i + 1 : 0] - vertices[i];
            Vec2 nrm(edge.y, -edge.x);
            nrm.normalize();
            normals.push_back(nrm);
        }
        computeProperties();
    }
    void computeProperties() {
        const int m = (int)vertices.size();
        Vec2 center(0.0f, 0.0f);
        float area = 0.0f, I = 0.0f;
        Vec2 s(0.0f, 0.0f);
        for (int i = 0; i < m; ++i) s += vertices[i];
        s *= 1.0f / (float)m;
        const float inv3 = 1.0f / 3.0f;
        for (int i = 0; i < m; ++i) {
            const Vec2 e1 = vertices[i] - s;
            const Vec2 e2 = i + 1 < m ? vertices[i + 1] - s : vertices[0] - s;
            const float D = cross(e1, e2);
            const float triangleArea = 0.5f * D;
            area += triangleArea;
            center += triangleArea * inv3 * (e1 + e2);
            const float ex1 = e1.x, ey1 = e1.y, ex2 = e2.x, ey2 = e2.y;
            const float intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2;
            const float inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
            I += (0.25f * inv3 * D) * (intx2 + inty2);
        }
        md.mass = density * area;
        center *= 1.0f / area;
        md.center = center + s;
        md.inertia = density * I;
        md.inertia += md.mass * (dot(md.center, md.center) - dot(center, center));
    }
    AABB computeAABB(const Transform& xf) const override {
        Vec2 lo = mul(xf, vertices[0]), hi = lo;
        for (size_t i = 1; i < vertices.size(); ++i) { const Vec2 v = mul(xf, vertices[i]); lo = vmin(lo, v); hi = vmax(hi, v); }
        const Vec2 r(radius, radius);
        return AABB{lo - r, hi + r};
    }
    MassData massData() const override { return md; }
};

// ------------------------------------------------------------------------------------------------
// manifolds and the narrow phase
// ------------------------------------------------------------------------------------------------
enum FeatureType : uint8_t { FEATURE_VERTEX = 0, FEATURE_FACE = 1 };
struct ContactFeature {
    uint8_t indexA = 0, indexB = 0, typeA = 0, typeB = 0;
    uint32_t key() const { return (uint32_t)indexA | ((uint32_t)indexB << 8) | ((uint32_t)typeA << 16) | ((uint32_t)typeB << 24); }
};
struct ManifoldPoint { Vec2 localPoint; float normalImpulse = 0.0f, tangentImpulse = 0.0f; ContactFeature id; };
struct Manifold {
    enum Type { CIRCLES = 0, FACE_A = 1, FACE_B = 2 };
    ManifoldPoint points[kMaxManifoldPoints];
    Vec2 localNormal, localPoint;
    Type type = CIRCLES;
    int pointCount = 0;
};
struct WorldManifold {
    Vec2 normal, points[kMaxManifoldPoints];
    void initialize(const Manifold& m, const Transform& xfA, float radiusA, const Transform& xfB, float radiusB) {
        if (m.pointCount == 0) return;
        if (m.type == Manifold::FACE_A) {
            normal = mul(xfA.q, m.localNormal);
            const Vec2 planePoint = mul(xfA, m.localPoint);
            for (int i = 0; i < m.pointCount; ++i) {
                const Vec2 clipPoint = mul(xfB, m.points[i].localPoint);
                const Vec2 cA = clipPoint + (radiusA - dot(clipPoint - planePoint, normal)) * normal;
                const Vec2 cB = clipPoint - radiusB * normal;
                points[i] = 0.5f * (cA + cB);
            }
        } else if (m.type == Manifold::FACE_B) {
            normal = mul(xfB.q, m.localNormal);
            const Vec2 planePoint = mul(xfB, m.localPoint);
            for (int i = 0; i < m.pointCount; ++i) {
                const Vec2 clipPoint = mul(xfA, m.points[i].localPoint);
                const Vec2 cB = clipPoint + (radiusB - dot(clipPoint - planePoint, normal)) * normal;
                const Vec2 cA = clipPoint - radiusA * normal;
                points[i] = 0.5f * (cA + cB);
            }
            normal = -normal;   // ensure the normal points from A to B
        }
    }
};

struct ClipVertex { Vec2 v; ContactFeature id; };

inline int clipSegmentToLine(ClipVertex vOut[2], const ClipVertex vIn[2], const Vec2& normal, float offset, int vertexIndexA) {
    int numOut = 0;
    const float distance0 = dot(normal, vIn[0].v) - offset;
    const float distance1 = dot(normal, vIn[1].v) - offset;
    if (distance0 <= 0.0f) vOut[numOut++] = vIn[0];
    if (distance1 <= 0.0f) vOut[numOut++] = vIn[1];
    if (distance0 * distance1 < 0.0f) {
        const float interp = distance0 / (distance0 - distance1);
        vOut[numOut].v = vIn[0].v + interp * (vIn[1].v - vIn[0].v);
        vOut[numOut].id.indexA = (uint8_t)vertexIndexA;   // vertex A is hitting edge B
        vOut[numOut].id.indexB = vIn[0].id.indexB;
        vOut[numOut].id.typeA = FEATURE_VERTEX;
        vOut[numOut].id.typeB = FEATURE_FACE;
        ++numOut;
    }
    return numOut;
}

// Edge (shape A) against convex polygon (shape B): the EPCollider of b2CollideEdge.cpp, including the
// adjacency handling of chained edges (hasVertex0 / hasVertex3), which isolated EdgeShapes leave unset.
class EdgePolygonCollider {
    struct Axis { enum Type { UNKNOWN, EDGE_A, EDGE_B } type; int index; float separation; };
    struct TempPolygon { Vec2 vertices[kMaxPolygonVertices], normals[kMaxPolygonVertices]; int count; };
    struct ReferenceFace { int i1, i2; Vec2 v1, v2, normal, sideNormal1, sideNormal2; float sideOffset1, sideOffset2; };
    TempPolygon polyB_;
    Transform xf_;
    Vec2 centroidB_, v0_, v1_, v2_, v3_, normal0_, normal1_, normal2_, normal_, lowerLimit_, upperLimit_;
    float radius_ = 0.0f;
    bool front_ = false;

    Axis edgeSeparation() const {
        Axis axis{Axis::EDGE_A, front_ ? 0 : 1, kMaxFloat};
        for (int i = 0; i < polyB_.count; ++i) {
            const float s = dot(normal_, polyB_.vertices[i] - v1_);
            if (s < axis.separation) axis.separation = s;
        }
        return axis;
    }
    Axis polygonSeparation() const {
        Axis axis{Axis::UNKNOWN, -1, -kMaxFloat};
        const Vec2 perp(-normal_.y, normal_.x);
        for (int i = 0; i < polyB_.count; ++i) {
            const Vec2 n = -polyB_.normals[i];
            const float s1 = dot(n, polyB_.vertices[i] - v1_);
            const float s2 = dot(n, polyB_.vertices[i] - v2_);
            const float s = std::min(s1, s2);
            if (s > radius_) { axis.type = Axis::EDGE_B; axis.index = i; axis.separation = s; return axis; }   // no collision
            if (dot(n, perp) >= 0.0f) { if (dot(n - upperLimit_, normal_) < -kAngularSlop) continue; }
            else { if (dot(n - lowerLimit_, normal_) < -kAngularSlop) continue; }
            if (s > axis.separation) { axis.type = Axis::EDGE_B; axis.index = i; axis.separation = s; }
        }
        return axis;
    }

public:
    void collide(Manifold* manifold, const EdgeShape& edgeA, const Transform& xfA, const PolygonShape& polygonB, const Transform& xfB) {
        xf_ = mulT(xfA, xfB);
        centroidB_ = mul(xf_, polygonB.md.center);
        v0_ = edgeA.v0; v1_ = edgeA.v1; v2_ = edgeA.v2; v3_ = edgeA.v3;
        const bool hasVertex0 = edgeA.hasVertex0, hasVertex3 = edgeA.hasVertex3;
        Vec2 edge1 = v2_ - v1_;
        edge1.normalize();
        normal1_ = Vec2(edge1.y, -edge1.x);
        const float offset1 = dot(normal1_, centroidB_ - v1_);
        float offset0 = 0.0f, offset2 = 0.0f;
        bool convex1 = false, convex2 = false;
        if (hasVertex0) {
            Vec2 edge0 = v1_ - v0_;
            edge0.normalize();
            normal0_ = Vec2(edge0.y, -edge0.x);
            convex1 = cross(edge0, edge1) >= 0.0f;
            offset0 = dot(normal0_, centroidB_ - v0_);
        }
        if (hasVertex3) {
            Vec2 edge2 = v3_ - v2_;
            edge2.normalize();
            normal2_ = Vec2(edge2.y, -edge2.x);
            convex2 = cross(edge1, edge2) > 0.0f;
            offset2 = dot(normal2_, centroidB_ - v2_);
        }
        // front / back side and the admissible range of collision normals
        if (hasVertex0 && hasVertex3) {
            if (convex1 && convex2) {
                front_ = offset0 >= 0.0f || offset1 >= 0.0f || offset2 >= 0.0f;
                if (front_) { normal_ = normal1_; lowerLimit_ = normal0_; upperLimit_ = normal2_; }
                else { normal_ = -normal1_; lowerLimit_ = -normal1_; upperLimit_ = -normal1_; }
            } else if (convex1) {
                front_ = offset0 >= 0.0f || (offset1 >= 0.0f && offset2 >= 0.0f);
                if (front_) { normal_ = normal1_; lowerLimit_ = normal0_; upperLimit_ = normal1_; }
                else { normal_ = -normal1_; lowerLimit_ = -normal2_; upperLimit_ = -normal1_; }
            } else if (convex2) {
                front_ = offset2 >= 0.0f || (offset0 >= 0.0f && offset1 >= 0.0f);
                if (front_) { normal_ = normal1_; lowerLimit_ = normal1_; upperLimit_ = normal2_; }
                else { normal_ = -normal1_; lowerLimit_ = -normal1_; upperLimit_ = -normal0_; }
            } else {
                front_ = offset0 >= 0.0f && offset1 >= 0.0f && offset2 >= 0.0f;
                if (front_) { normal_ = normal1_; lowerLimit_ = normal1_; upperLimit_ = normal1_; }
                else { normal_ = -normal1_; lowerLimit_ = -normal2_; upperLimit_ = -normal0_; }
            }
        } else if (hasVertex0) {
            if (convex1) {
                front_ = offset0 >= 0.0f || offset1 >= 0.0f;
                if (front_) { normal_ = normal1_; lowerLimit_ = normal0_; upperLimit_ = -normal1_; }
                else { normal_ = -normal1_; lowerLimit_ = normal1_; upperLimit_ = -normal1_; }
            } else {
                front_ = offset0 >= 0.0f && offset1 >= 0.0f;
                if (front_) { normal_ = normal1_; lowerLimit_ = normal1_; upperLimit_ = -normal1_; }
                else { normal_ = -normal1_; lowerLimit_ = normal1_; upperLimit_ = -normal0_; }
            }
        } else if (hasVertex3) {
            if (convex2) {
                front_ = offset1 >= 0.0f || offset2 >= 0.0f;
                if (front_) { normal_ = normal1_; lowerLimit_ = -normal1_; upperLimit_ = normal2_; }
                else { normal_ = -normal1_; lowerLimit_ = -normal1_; upperLimit_ = normal1_; }
            } else {
                front_ = offset1 >= 0.0f && offset2 >= 0.0f;
                if (front_) { normal_ = normal1_; lowerLimit_ = -normal1_; upperLimit_ = normal1_; }
                else { normal_ = -normal1_; lowerLimit_ = -normal2_; upperLimit_ = normal1_; }
            }
        } else {
            front_ = offset1 >= 0.0f;
            if (front_) { normal_ = normal1_; lowerLimit_ = -normal1_; upperLimit_ = -normal1_; }
            else { normal_ = -normal1_; lowerLimit_ = normal1_; upperLimit_ = normal1_; }
        }

        polyB_.count = (int)polygonB.vertices.size();
        for (int i = 0; i < polyB_.count; ++i) {
            polyB_.vertices[i] = mul(xf_, polygonB.vertices[i]);
            polyB_.normals[i] = mul(xf_.q, polygonB.normals[i]);
        }
        radius_ = 2.0f * kPolygonRadius;
        manifold->pointCount = 0;

        const Axis edgeAxis = edgeSeparation();
        if (edgeAxis.type == Axis::UNKNOWN) return;
        if (edgeAxis.separation > radius_) return;
        const Axis polygonAxis = polygonSeparation();
        if (polygonAxis.type != Axis::UNKNOWN && polygonAxis.separation > radius_) return;

        const float relativeTol = 0.98f, absoluteTol = 0.001f;   // hysteresis for jitter reduction
        Axis primary;
        if (polygonAxis.type == Axis::UNKNOWN) primary = edgeAxis;
        else if (polygonAxis.separation > relativeTol * edgeAxis.separation + absoluteTol) primary = polygonAxis;
        else primary = edgeAxis;

        ClipVertex ie[2];
        ReferenceFace rf;
        if (primary.type == Axis::EDGE_A) {
            manifold->type = Manifold::FACE_A;
            int bestIndex = 0;   // the polygon normal most anti-parallel to the edge normal
            float bestValue = dot(normal_, polyB_.normals[0]);
            for (int i = 1; i < polyB_.count; ++i) {
                const float value = dot(normal_, polyB_.normals[i]);
                if (value < bestValue) { bestValue = value; bestIndex = i; }
            }
            const int i1 = bestIndex, i2 = i1 + 1 < polyB_.count ? i1 + 1 : 0;
            ie[0].v = polyB_.vertices[i1]; ie[0].id = ContactFeature{0, (uint8_t)i1, FEATURE_FACE, FEATURE_VERTEX};
            ie[1].v = polyB_.vertices[i2]; ie[1].id = ContactFeature{0, (uint8_t)i2, FEATURE_FACE, FEATURE_VERTEX};
            if (front_) { rf.i1 = 0; rf.i2 = 1; rf.v1 = v1_; rf.v2 = v2_; rf.normal = normal1_; }
            else { rf.i1 = 1; rf.i2 = 0; rf.v1 = v2_; rf.v2 = v1_; rf.normal = -normal1_; }
        } else {
            manifold->type = Manifold::FACE_B;
            ie[0].v = v1_; ie[0].id = ContactFeature{0, (uint8_t)primary.index, FEATURE_VERTEX, FEATURE_FACE};
            ie[1].v = v2_; ie[1].id = ContactFeature{0, (uint8_t)primary.index, FEATURE_VERTEX, FEATURE_FACE};
            rf.i1 = primary.index;
            rf.i2 = rf.i1 + 1 < polyB_.count ? rf.i1 + 1 : 0;
            rf.v1 = polyB_.vertices[rf.i1]; rf.v2 = polyB_.vertices[rf.i2]; rf.normal = polyB_.normals[rf.i1];
        }
        rf.sideNormal1 = Vec2(rf.normal.y, -rf.normal.x);
        rf.sideNormal2 = -rf.sideNormal1;
        rf.sideOffset1 = dot(rf.sideNormal1, rf.v1);
        rf.sideOffset2 = dot(rf.sideNormal2, rf.v2);

        ClipVertex clip1[2], clip2[2];
        if (clipSegmentToLine(clip1, ie, rf.sideNormal1, rf.sideOffset1, rf.i1) < kMaxManifoldPoints) return;
        if (clipSegmentToLine(clip2, clip1, rf.sideNormal2, rf.sideOffset2, rf.i2) < kMaxManifoldPoints) return;

        if (primary.type == Axis::EDGE_A) { manifold->localNormal = rf.normal; manifold->localPoint = rf.v1; }
        else { manifold->localNormal = polygonB.normals[rf.i1]; manifold->localPoint = polygonB.vertices[rf.i1]; }
        int pointCount = 0;
        for (int i = 0; i < kMaxManifoldPoints; ++i) {
            const float separation = dot(rf.normal, clip2[i].v - rf.v1);
            if (separation <= radius_) {
                ManifoldPoint& cp = manifold->points[pointCount];
                if (primary.type == Axis::EDGE_A) {
                    cp.localPoint = mulT(xf_, clip2[i].v);
                    cp.id = clip2[i].id;
                } else {
                    cp.localPoint = clip2[i].v;
                    cp.id.typeA = clip2[i].id.typeB; cp.id.typeB = clip2[i].id.typeA;
                    cp.id.indexA = clip2[i].id.indexB; cp.id.indexB = clip2[i].id.indexA;
                }
                ++pointCount;
            }
        }
        manifold->pointCount = pointCount;
    }
};

// ------------------------------------------------------------------------------------------------
// GJK distance and conservative advancement (b2Distance.cpp, b2TimeOfImpact.cpp) -- used by the TOI pass
// ------------------------------------------------------------------------------------------------
struct DistanceProxy {
    Vec2 buffer[2];
    const Vec2* vertices = nullptr;
    int count = 0;
    float radius = 0.0f;
    void set(const Shape* shape) {
        if (shape->type == Shape::POLYGON) {
            const PolygonShape* p = static_cast<const PolygonShape*>(shape);
            vertices = p->vertices.data(); count = (int)p->vertices.size(); radius = p->radius;
        } else {
            const EdgeShape* e = static_cast<const EdgeShape*>(shape);
            buffer[0] = e->v1; buffer[1] = e->v2; vertices = buffer; count = 2; radius = e->radius;
        }
    }
    int support(const Vec2& d) const {
        int best = 0;
        float bestValue = dot(vertices[0], d);
        for (int i = 1; i < count; ++i) { const float v = dot(vertices[i], d); if (v > bestValue) { bestValue = v; best = i; } }
        return best;
    }
    const Vec2& vertex(int i) const { return vertices[i]; }
};
struct SimplexCache { float metric = 0.0f; uint16_t count = 0; uint8_t indexA[3] = {0, 0, 0}, indexB[3] = {0, 0, 0}; };
struct DistanceInput { DistanceProxy proxyA, proxyB; Transform transformA, transformB; bool useRadii = false; };
struct DistanceOutput { Vec2 pointA, pointB; float distance = 0.0f; int iterations = 0; };

struct SimplexVertex { Vec2 wA, wB, w; float a = 0.0f; int indexA = 0, indexB = 0; };
struct Simplex {
    SimplexVertex v[3];
    int count = 0;
    void readCache(const SimplexCache* cache, const DistanceProxy* proxyA, const Transform& xfA, const DistanceProxy* proxyB, const Transform& xfB) {
        count = cache->count;
        for (int i = 0; i < count; ++i) {
            SimplexVertex* s = v + i;
            s->indexA = cache->indexA[i]; s->indexB = cache->indexB[i];
            s->wA = mul(xfA, proxyA->vertex(s->indexA));
            s->wB = mul(xfB, proxyB->vertex(s->indexB));
            s->w = s->wB - s->wA;
            s->a = 0.0f;
        }
        if (count > 1) {   // flush the simplex if the metric changed a lot
            const float metric1 = cache->metric, metric2 = metric();
            if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < kEpsilon) count = 0;
        }
        if (count == 0) {
            SimplexVertex* s = v;
            s->indexA = 0; s->indexB = 0;
            s->wA = mul(xfA, proxyA->vertex(0));
            s->wB = mul(xfB, proxyB->vertex(0));
            s->w = s->wB - s->wA;
            s->a = 1.0f;
            count = 1;
        }
    }
    void writeCache(SimplexCache* cache) const {
        cache->metric = metric();
        cache->count = (uint16_t)count;
        for (int i = 0; i < count; ++i) { cache->indexA[i] = (uint8_t)v[i].indexA; cache->indexB[i] = (uint8_t)v[i].indexB; }
    }
    Vec2 searchDirection() const {
        if (count == 1) return -v[0].w;
        const Vec2 e12 = v[1].w - v[0].w;
        const float sgn = cross(e12, -v[0].w);
        return sgn > 0.0f ? cross(1.0f, e12) : cross(e12, 1.0f);
    }
    Vec2 closestPoint() const {
        if (count == 1) return v[0].w;
        if (count == 2) return v[0].a * v[0].w + v[1].a * v[1].w;
        return Vec2(0.0f, 0.0f);
    }
    void witnessPoints(Vec2* pA, Vec2* pB) const {
        if (count == 1) { *pA = v[0].wA; *pB = v[0].wB; }
        else if (count == 2) { *pA = v[0].a * v[0].wA + v[1].a * v[1].wA; *pB = v[0].a * v[0].wB + v[1].a * v[1].wB; }
        else { *pA = v[0].a * v[0].wA + v[1].a * v[1].wA + v[2].a * v[2].wA; *pB = *pA; }
    }
    float metric() const {
        if (count == 2) return (v[0].w - v[1].w).length();
        if (count == 3) return cross(v[1].w - v[0].w, v[2].w - v[0].w);
        return 0.0f;
    }
    void solve2() {
        const Vec2 w1 = v[0].w, w2 = v[1].w, e12 = w2 - w1;
        const float d12_2 = -dot(w1, e12);
        if (d12_2 <= 0.0f) { v[0].a = 1.0f; count = 1; return; }
        const float d12_1 = dot(w2, e12);
        if (d12_1 <= 0.0f) { v[1].a = 1.0f; count = 1; v[0] = v[1]; return; }
        const float inv = 1.0f / (d12_1 + d12_2);
        v[0].a = d12_1 * inv; v[1].a = d12_2 * inv; count = 2;
    }
    void solve3() {
        const Vec2 w1 = v[0].w, w2 = v[1].w, w3 = v[2].w;
        const Vec2 e12 = w2 - w1;
        const float d12_1 = dot(w2, e12), d12_2 = -dot(w1, e12);
        const Vec2 e13 = w3 - w1;
        const float d13_1 = dot(w3, e13), d13_2 = -dot(w1, e13);
        const Vec2 e23 = w3 - w2;
        const float d23_1 = dot(w3, e23), d23_2 = -dot(w2, e23);
        const float n123 = cross(e12, e13);
        const float d123_1 = n123 * cross(w2, w3), d123_2 = n123 * cross(w3, w1), d123_3 = n123 * cross(w1, w2);
        if (d12_2 <= 0.0f && d13_2 <= 0.0f) { v[0].a = 1.0f; count = 1; return; }
        if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) { const float inv = 1.0f / (d12_1 + d12_2); v[0].a = d12_1 * inv; v[1].a = d12_2 * inv; count = 2; return; }
        if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) { const float inv = 1.0f / (d13_1 + d13_2); v[0].a = d13_1 * inv; v[2].a = d13_2 * inv; count = 2; v[1] = v[2]; return; }
        if (d12_1 <= 0.0f && d23_2 <= 0.0f) { v[1].a = 1.0f; count = 1; v[0] = v[1]; return; }
        if (d13_1 <= 0.0f && d23_1 <= 0.0f) { v[2].a = 1.0f; count = 1; v[0] = v[2]; return; }
        if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) { const float inv = 1.0f / (d23_1 + d23_2); v[1].a = d23_1 * inv; v[2].a = d23_2 * inv; count = 2; v[0] = v[2]; return; }
        const float inv = 1.0f / (d123_1 + d123_2 + d123_3);
        v[0].a = d123_1 * inv; v[1].a = d123_2 * inv; v[2].a = d123_3 * inv; count = 3;
    }
};

inline void distance(DistanceOutput* output, SimplexCache* cache, const DistanceInput* input) {
    const DistanceProxy* proxyA = &input->proxyA;
    const DistanceProxy* proxyB = &input->proxyB;
    const Transform xfA = input->transformA, xfB = input->transformB;
    Simplex simplex;
    simplex.readCache(cache, proxyA, xfA, proxyB, xfB);
    SimplexVertex* vertices = simplex.v;
    const int maxIters = 20;
    int saveA[3], saveB[3], saveCount = 0;
    float distanceSqr1 = kMaxFloat, distanceSqr2 = distanceSqr1;
    int iter = 0;
    while (iter < maxIters) {
        saveCount = simplex.count;
        for (int i = 0; i < saveCount; ++i) { saveA[i] = vertices[i].indexA; saveB[i] = vertices[i].indexB; }
        if (simplex.count == 2) simplex.solve2();
        else if (simplex.count == 3) simplex.solve3();
        if (simplex.count == 3) break;   // the origin is inside the triangle
        const Vec2 p = simplex.closestPoint();
        distanceSqr2 = p.lengthSquared();
        if (distanceSqr2 >= distanceSqr1) { /* no progress -- b2Distance keeps going (the test is commented out upstream) */ }
        distanceSqr1 = distanceSqr2;
        const Vec2 d = simplex.searchDirection();
        if (d.lengthSquared() < kEpsilon * kEpsilon) break;   // the origin is on a segment or a vertex
        SimplexVertex* vertex = vertices + simplex.count;
        vertex->indexA = proxyA->support(mulT(xfA.q, -d));
        vertex->wA = mul(xfA, proxyA->vertex(vertex->indexA));
        vertex->indexB = proxyB->support(mulT(xfB.q, d));
        vertex->wB = mul(xfB, proxyB->vertex(vertex->indexB));
        vertex->w = vertex->wB - vertex->wA;
        ++iter;
        bool duplicate = false;
        for (int i = 0; i < saveCount; ++i) if (vertex->indexA == saveA[i] && vertex->indexB == saveB[i]) { duplicate = true; break; }
        if (duplicate) break;
        ++simplex.count;
    }
    simplex.witnessPoints(&output->pointA, &output->pointB);
    output->distance = (output->pointA - output->pointB).length();
    output->iterations = iter;
    simplex.writeCache(cache);
    if (input->useRadii) {
        const float rA = proxyA->radius, rB = proxyB->radius;
        if (output->distance > rA + rB && output->distance > kEpsilon) {
            output->distance -= rA + rB;
            Vec2 normal = output->pointB - output->pointA;
            normal.normalize();
            output->pointA += rA * normal;
            output->pointB -= rB * normal;
        } else {
            const Vec2 p = 0.5f * (output->pointA + output->pointB);
            output->pointA = p; output->pointB = p; output->distance = 0.0f;
        }
    }
}

struct TOIInput { DistanceProxy proxyA, proxyB; Sweep sweepA, sweepB; float tMax = 1.0f; };
struct TOIOutput { enum State { UNKNOWN, FAILED, OVERLAPPED, TOUCHING, SEPARATED } state = UNKNOWN; float t = 0.0f; };

struct SeparationFunction {
    enum Type { POINTS, FACE_A, FACE_B };
    const DistanceProxy* proxyA = nullptr;
    const DistanceProxy* proxyB = nullptr;
    Sweep sweepA, sweepB;
    Type type = POINTS;
    Vec2 localPoint, axis;

    float initialize(const SimplexCache* cache, const DistanceProxy* pA, const Sweep& sA, const DistanceProxy* pB, const Sweep& sB, float t1) {
        proxyA = pA; proxyB = pB;
        const int count = cache->count;
        sweepA = sA; sweepB = sB;
        Transform xfA, xfB;
        sweepA.getTransform(&xfA, t1);
        sweepB.getTransform(&xfB, t1);
        if (count == 1) {
            type = POINTS;
            const Vec2 localPointA = proxyA->vertex(cache->indexA[0]), localPointB = proxyB->vertex(cache->indexB[0]);
            const Vec2 pointA = mul(xfA, localPointA), pointB = mul(xfB, localPointB);
            axis = pointB - pointA;
            return axis.normalize();
        } else if (cache->indexA[0] == cache->indexA[1]) {
            type = FACE_B;   // two points on B, one on A
            const Vec2 localPointB1 = proxyB->vertex(cache->indexB[0]), localPointB2 = proxyB->vertex(cache->indexB[1]);
            axis = cross(localPointB2 - localPointB1, 1.0f);
            axis.normalize();
            const Vec2 normal = mul(xfB.q, axis);
            localPoint = 0.5f * (localPointB1 + localPointB2);
            const Vec2 pointB = mul(xfB, localPoint);
            const Vec2 pointA = mul(xfA, proxyA->vertex(cache->indexA[0]));
            float s = dot(pointA - pointB, normal);
            if (s < 0.0f) { axis = -axis; s = -s; }
            return s;
        } else {
            type = FACE_A;   // two points on A, one or two on B
            const Vec2 localPointA1 = proxyA->vertex(cache->indexA[0]), localPointA2 = proxyA->vertex(cache->indexA[1]);
            axis = cross(localPointA2 - localPointA1, 1.0f);
            axis.normalize();
            const Vec2 normal = mul(xfA.q, axis);
            localPoint = 0.5f * (localPointA1 + localPointA2);
            const Vec2 pointA = mul(xfA, localPoint);
            const Vec2 pointB = mul(xfB, proxyB->vertex(cache->indexB[0]));
            float s = dot(pointB - pointA, normal);
            if (s < 0.0f) { axis = -axis; s = -s; }
            return s;
        }
    }
    float findMinSeparation(int* indexA, int* indexB, float t) const {
        Transform xfA, xfB;
        sweepA.getTransform(&xfA, t);
        sweepB.getTransform(&xfB, t);
        switch (type) {
        case POINTS: {
            const Vec2 axisA = mulT(xfA.q, axis), axisB = mulT(xfB.q, -axis);
            *indexA = proxyA->support(axisA); *indexB = proxyB->support(axisB);
            const Vec2 pointA = mul(xfA, proxyA->vertex(*indexA)), pointB = mul(xfB, proxyB->vertex(*indexB));
            return dot(pointB - pointA, axis);
        }
        case FACE_A: {
            const Vec2 normal = mul(xfA.q, axis), pointA = mul(xfA, localPoint);
            const Vec2 axisB = mulT(xfB.q, -normal);
            *indexA = -1; *indexB = proxyB->support(axisB);
            const Vec2 pointB = mul(xfB, proxyB->vertex(*indexB));
            return dot(pointB - pointA, normal);
        }
        default: {
            const Vec2 normal = mul(xfB.q, axis), pointB = mul(xfB, localPoint);
            const Vec2 axisA = mulT(xfA.q, -normal);
            *indexB = -1; *indexA = proxyA->support(axisA);
            const Vec2 pointA = mul(xfA, proxyA->vertex(*indexA));
            return dot(pointA - pointB, normal);
        }
        }
    }
    float evaluate(int indexA, int indexB, float t) const {
        Transform xfA, xfB;
        sweepA.getTransform(&xfA, t);
        sweepB.getTransform(&xfB, t);
        switch (type) {
        case POINTS: {
            const Vec2 pointA = mul(xfA, proxyA->vertex(indexA)), pointB = mul(xfB, proxyB->vertex(indexB));
            return dot(pointB - pointA, axis);
        }
        case FACE_A: {
            const Vec2 normal = mul(xfA.q, axis), pointA = mul(xfA, localPoint), pointB = mul(xfB, proxyB->vertex(indexB));
            return dot(pointB - pointA, normal);
        }
        default: {
            const Vec2 normal = mul(xfB.q, axis), pointB = mul(xfB, localPoint), pointA = mul(xfA, proxyA->vertex(indexA));
            return dot(pointA - pointB, normal);
        }
        }
    }
};

// b2TimeOfImpact: conservative advancement with a local separating axis and a bracketed root finder
inline void timeOfImpact(TOIOutput* output, const TOIInput* input) {
    output->state = TOIOutput::UNKNOWN;
    output->t = input->tMax;
    const DistanceProxy* proxyA = &input->proxyA;
    const DistanceProxy* proxyB = &input->proxyB;
    Sweep sweepA = input->sweepA, sweepB = input->sweepB;
    // (b2Sweep::Normalize: angles of the lander never leave [-2pi, 2pi] by more than a step; kept for fidelity)
    auto normalize = [](Sweep& s) { const float twoPi = 2.0f * kPi; const float d = twoPi * std::floor(s.a0 / twoPi); s.a0 -= d; s.a -= d; };
    normalize(sweepA); normalize(sweepB);
    const float tMax = input->tMax;
    const float totalRadius = proxyA->radius + proxyB->radius;
    const float target = std::max(kLinearSlop, totalRadius - 3.0f * kLinearSlop);
    const float tolerance = 0.25f * kLinearSlop;
    float t1 = 0.0f;
    const int maxIterations = 20;
    int iter = 0;
    SimplexCache cache;
    DistanceInput distanceInput;
    distanceInput.proxyA = input->proxyA; distanceInput.proxyB = input->proxyB; distanceInput.useRadii = false;
    for (;;) {
        Transform xfA, xfB;
        sweepA.getTransform(&xfA, t1);
        sweepB.getTransform(&xfB, t1);
        distanceInput.transformA = xfA; distanceInput.transformB = xfB;
        DistanceOutput distanceOutput;
        distance(&distanceOutput, &cache, &distanceInput);
        if (distanceOutput.distance <= 0.0f) { output->state = TOIOutput::OVERLAPPED; output->t = 0.0f; break; }
        if (distanceOutput.distance < target + tolerance) { output->state = TOIOutput::TOUCHING; output->t = t1; break; }
        SeparationFunction fcn;
        fcn.initialize(&cache, proxyA, sweepA, proxyB, sweepB, t1);
        bool done = false;
        float t2 = tMax;
        int pushBackIter = 0;
        for (;;) {
            int indexA, indexB;
            float s2 = fcn.findMinSeparation(&indexA, &indexB, t2);
            if (s2 > target + tolerance) { output->state = TOIOutput::SEPARATED; output->t = tMax; done = true; break; }
            if (s2 > target - tolerance) { t1 = t2; break; }   // advance the sweeps
            float s1 = fcn.evaluate(indexA, indexB, t1);
            if (s1 < target - tolerance) { output->state = TOIOutput::FAILED; output->t = t1; done = true; break; }
            if (s1 <= target + tolerance) { output->state = TOIOutput::TOUCHING; output->t = t1; done = true; break; }
            int rootIterCount = 0;
            float a1 = t1, a2 = t2;
            for (;;) {
                float t;
                if (rootIterCount & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);   // secant
                else t = 0.5f * (a1 + a2);                                               // bisection
                ++rootIterCount;
                const float s = fcn.evaluate(indexA, indexB, t);
                if (std::fabs(s - target) < tolerance) { t2 = t; break; }
                if (s > target) { a1 = t; s1 = s; } else { a2 = t; s2 = s; }
                if (rootIterCount == 50) break;
            }
            ++pushBackIter;
            if (pushBackIter == kMaxPolygonVertices) break;
        }
        ++iter;
        if (done) break;
        if (iter == maxIterations) { output->state = TOIOutput::FAILED; output->t = t1; break; }
    }
}

// ------------------------------------------------------------------------------------------------
// dynamics: bodies, fixtures, contacts, joints
// ------------------------------------------------------------------------------------------------
class World;
struct Body;
struct Contact;
struct Joint;

struct Fixture {
    Body* body = nullptr;
    std::unique_ptr<Shape> shape;
    float friction = kDefaultFriction, restitution = kDefaultRestitution;
    uint32_t category = 0x0001u;      // Category.Cat1
    uint32_t collidesWith = 0xffffffffu;   // Category.All
    int16_t group = 0;
    int proxyId = -1;
    int userIndex = -1;               // free for the caller (lunar_sim: terrain edge number)
};

struct ContactEdge { Body* other = nullptr; Contact* contact = nullptr; ContactEdge* prev = nullptr; ContactEdge* next = nullptr; };
struct JointEdge { Body* other = nullptr; Joint* joint = nullptr; JointEdge* prev = nullptr; JointEdge* next = nullptr; };

enum BodyType { STATIC_BODY = 0, KINEMATIC_BODY = 1, DYNAMIC_BODY = 2 };

struct Body {
    World* world = nullptr;
    BodyType type = STATIC_BODY;
    Transform xf;
    Sweep sweep;
    Vec2 linearVelocity;
    float angularVelocity = 0.0f;
    Vec2 force;
    float torque = 0.0f;
    float mass = 0.0f, invMass = 0.0f, inertia = 0.0f, invI = 0.0f;
    float linearDamping = 0.0f, angularDamping = 0.0f, gravityScale = 1.0f;
    float sleepTime = 0.0f;
    bool awake = true, autoSleep = true, enabled = true, bullet = false, fixedRotation = false;
    bool islandFlag = false, toiFlag = false;
    int islandIndex = 0;
    std::vector<std::unique_ptr<Fixture>> fixtures;
    ContactEdge* contactList = nullptr;
    JointEdge* jointList = nullptr;
    int userIndex = -1;

    Vec2 position() const { return xf.p; }
    float rotation() const { return sweep.a; }
    void setAwake(bool flag) {
        if (flag) { if (!awake) { awake = true; sleepTime = 0.0f; } }
        else { awake = false; sleepTime = 0.0f; linearVelocity = Vec2(0.0f, 0.0f); angularVelocity = 0.0f; force = Vec2(0.0f, 0.0f); torque = 0.0f; }
    }
    void setTransform(const Vec2& position, float angle);   // Body.Position / Body.Rotation setters
    void setPosition(const Vec2& p) { setTransform(p, sweep.a); }
    void setRotation(float a) { setTransform(xf.p, a); }
    void setType(BodyType t);
    Fixture* createFixture(std::unique_ptr<Shape> shape);
    void resetMassData();
    void synchronizeTransform() { xf.q.set(sweep.a); xf.p = sweep.c - mul(xf.q, sweep.localCenter); }
    void synchronizeFixtures();
    void advance(float alpha) {   // advance to the new safe time; does not sync the broad phase
        sweep.advance(alpha);
        sweep.c = sweep.c0; sweep.a = sweep.a0;
        xf.q.set(sweep.a);
        xf.p = sweep.c - mul(xf.q, sweep.localCenter);
    }
    bool shouldCollide(const Body* other) const;
    void applyForce(const Vec2& f, const Vec2& point) {
        if (type != DYNAMIC_BODY) return;
        if (!awake) setAwake(true);
        force += f;
        torque += (point.x - sweep.c.x) * f.y - (point.y - sweep.c.y) * f.x;
    }
    void applyTorque(float t) { if (type != DYNAMIC_BODY) return; if (!awake) setAwake(true); torque += t; }
    void applyLinearImpulse(const Vec2& impulse, const Vec2& point) {
        if (type != DYNAMIC_BODY) return;
        if (!awake) setAwake(true);
        linearVelocity += invMass * impulse;
        angularVelocity += invI * ((point.x - sweep.c.x) * impulse.y - (point.y - sweep.c.y) * impulse.x);
    }
};

struct Contact {
    Fixture* fixtureA = nullptr;
    Fixture* fixtureB = nullptr;
    Manifold manifold;
    ContactEdge nodeA, nodeB;
    float friction = 0.0f, restitution = 0.0f, tangentSpeed = 0.0f;
    bool touching = false, enabled = true, islandFlag = false, toiFlag = false;
    int toiCount = 0;
    float toi = 1.0f;
    uint64_t serial = 0;   // creation ordinal
    void evaluate(Manifold* m, const Transform& xfA, const Transform& xfB) const {
        // dispatch on the shape pair; Contact.Create has already put the edge first (Edge+Polygon is the registered order)
        EdgePolygonCollider collider;
        collider.collide(m, *static_cast<const EdgeShape*>(fixtureA->shape.get()), xfA, *static_cast<const PolygonShape*>(fixtureB->shape.get()), xfB);
    }
};

struct TimeStep { float dt, inv_dt, dtRatio; int velocityIterations, positionIterations; bool warmStarting; };
struct Position { Vec2 c; float a; };
struct Velocity { Vec2 v; float w; };
struct SolverData { TimeStep step; Position* positions; Velocity* velocities; };

struct Joint {
    Body* bodyA = nullptr;
    Body* bodyB = nullptr;
    JointEdge edgeA, edgeB;
    bool islandFlag = false, collideConnected = false;
    virtual ~Joint() = default;
    virtual void initVelocityConstraints(const SolverData& data) = 0;
    virtual void solveVelocityConstraints(const SolverData& data) = 0;
    virtual bool solvePositionConstraints(const SolverData& data) = 0;
};

enum LimitState { INACTIVE_LIMIT = 0, AT_LOWER_LIMIT = 1, AT_UPPER_LIMIT = 2, EQUAL_LIMITS = 3 };

// b2RevoluteJoint (2.3.0: rigid limits, three-row point + angle block)
struct RevoluteJoint : Joint {
    Vec2 localAnchorA, localAnchorB;
    Vec3 impulse;
    float motorImpulse = 0.0f;
    bool enableMotor = false, enableLimit = false;
    float maxMotorTorque = 0.0f, motorSpeed = 0.0f, referenceAngle = 0.0f, lowerAngle = 0.0f, upperAngle = 0.0f;
    LimitState limitState = INACTIVE_LIMIT;
    // solver temp
    int indexA = 0, indexB = 0;
    Vec2 rA, rB, localCenterA, localCenterB;
    float invMassA = 0.0f, invMassB = 0.0f, invIA = 0.0f, invIB = 0.0f, motorMass = 0.0f;
    Mat33 mass;

    RevoluteJoint(Body* a, Body* b, const Vec2& anchorA, const Vec2& anchorB) {
        bodyA = a; bodyB = b; localAnchorA = anchorA; localAnchorB = anchorB;
        referenceAngle = b->rotation() - a->rotation();
    }
    void initVelocityConstraints(const SolverData& data) override {
        indexA = bodyA->islandIndex; indexB = bodyB->islandIndex;
        localCenterA = bodyA->sweep.localCenter; localCenterB = bodyB->sweep.localCenter;
        invMassA = bodyA->invMass; invMassB = bodyB->invMass; invIA = bodyA->invI; invIB = bodyB->invI;
        const float aA = data.positions[indexA].a;
        Vec2 vA = data.velocities[indexA].v; float wA = data.velocities[indexA].w;
        const float aB = data.positions[indexB].a;
        Vec2 vB = data.velocities[indexB].v; float wB = data.velocities[indexB].w;
        const Rot qA(aA), qB(aB);
        rA = mul(qA, localAnchorA - localCenterA);
        rB = mul(qB, localAnchorB - localCenterB);
        const float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
        const bool fixedRotation = (iA + iB == 0.0f);
        mass.ex.x = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
        mass.ey.x = -rA.y * rA.x * iA - rB.y * rB.x * iB;
        mass.ez.x = -rA.y * iA - rB.y * iB;
        mass.ex.y = mass.ey.x;
        mass.ey.y = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
        mass.ez.y = rA.x * iA + rB.x * iB;
        mass.ex.z = mass.ez.x;
        mass.ey.z = mass.ez.y;
        mass.ez.z = iA + iB;
        motorMass = iA + iB;
        if (motorMass > 0.0f) motorMass = 1.0f / motorMass;
        if (!enableMotor || fixedRotation) motorImpulse = 0.0f;
        if (enableLimit && !fixedRotation) {
            const float jointAngle = aB - aA - referenceAngle;
            if (std::fabs(upperAngle - lowerAngle) < 2.0f * kAngularSlop) {
                limitState = EQUAL_LIMITS;
            } else if (jointAngle <= lowerAngle) {
                if (limitState != AT_LOWER_LIMIT) impulse.z = 0.0f;
                limitState = AT_LOWER_LIMIT;
            } else if (jointAngle >= upperAngle) {
                if (limitState != AT_UPPER_LIMIT) impulse.z = 0.0f;
                limitState = AT_UPPER_LIMIT;
            } else {
                limitState = INACTIVE_LIMIT;
                impulse.z = 0.0f;
            }
        } else {
            limitState = INACTIVE_LIMIT;
        }
        if (data.step.warmStarting) {
            impulse *= data.step.dtRatio;
            motorImpulse *= data.step.dtRatio;
            const Vec2 P(impulse.x, impulse.y);
            vA -= mA * P;
            wA -= iA * (cross(rA, P) + motorImpulse + impulse.z);
            vB += mB * P;
            wB += iB * (cross(rB, P) + motorImpulse + impulse.z);
        } else {
            impulse = Vec3(0.0f, 0.0f, 0.0f);
            motorImpulse = 0.0f;
        }
        data.velocities[indexA].v = vA; data.velocities[indexA].w = wA;
        data.velocities[indexB].v = vB; data.velocities[indexB].w = wB;
    }
    void solveVelocityConstraints(const SolverData& data) override {
        Vec2 vA = data.velocities[indexA].v; float wA = data.velocities[indexA].w;
        Vec2 vB = data.velocities[indexB].v; float wB = data.velocities[indexB].w;
        const float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
        const bool fixedRotation = (iA + iB == 0.0f);
        if (enableMotor && limitState != EQUAL_LIMITS && !fixedRotation) {
            const float Cdot = wB - wA - motorSpeed;
            float imp = -motorMass * Cdot;
            const float oldImpulse = motorImpulse;
            const float maxImpulse = data.step.dt * maxMotorTorque;
            motorImpulse = clampf(motorImpulse + imp, -maxImpulse, maxImpulse);
            imp = motorImpulse - oldImpulse;
            wA -= iA * imp;
            wB += iB * imp;
        }
        if (enableLimit && limitState != INACTIVE_LIMIT && !fixedRotation) {
            const Vec2 Cdot1 = vB + cross(wB, rB) - vA - cross(wA, rA);
            const float Cdot2 = wB - wA;
            const Vec3 Cdot(Cdot1.x, Cdot1.y, Cdot2);
            Vec3 imp = -mass.solve33(Cdot);
            if (limitState == EQUAL_LIMITS) {
                impulse += imp;
            } else if (limitState == AT_LOWER_LIMIT) {
                const float newImpulse = impulse.z + imp.z;
                if (newImpulse < 0.0f) {
                    const Vec2 rhs = -Cdot1 + impulse.z * Vec2(mass.ez.x, mass.ez.y);
                    const Vec2 reduced = mass.solve22(rhs);
                    imp.x = reduced.x; imp.y = reduced.y; imp.z = -impulse.z;
                    impulse.x += reduced.x; impulse.y += reduced.y; impulse.z = 0.0f;
                } else {
                    impulse += imp;
                }
            } else if (limitState == AT_UPPER_LIMIT) {
                const float newImpulse = impulse.z + imp.z;
                if (newImpulse > 0.0f) {
                    const Vec2 rhs = -Cdot1 + impulse.z * Vec2(mass.ez.x, mass.ez.y);
                    const Vec2 reduced = mass.solve22(rhs);
                    imp.x = reduced.x; imp.y = reduced.y; imp.z = -impulse.z;
                    impulse.x += reduced.x; impulse.y += reduced.y; impulse.z = 0.0f;
                } else {
                    impulse += imp;
                }
            }
            const Vec2 P(imp.x, imp.y);
            vA -= mA * P;
            wA -= iA * (cross(rA, P) + imp.z);
            vB += mB * P;
            wB += iB * (cross(rB, P) + imp.z);
        } else {
            const Vec2 Cdot = vB + cross(wB, rB) - vA - cross(wA, rA);
            const Vec2 imp = mass.solve22(-Cdot);
            impulse.x += imp.x; impulse.y += imp.y;
            vA -= mA * imp;
            wA -= iA * cross(rA, imp);
            vB += mB * imp;
            wB += iB * cross(rB, imp);
        }
        data.velocities[indexA].v = vA; data.velocities[indexA].w = wA;
        data.velocities[indexB].v = vB; data.velocities[indexB].w = wB;
    }
    bool solvePositionConstraints(const SolverData& data) override {
        Vec2 cA = data.positions[indexA].c; float aA = data.positions[indexA].a;
        Vec2 cB = data.positions[indexB].c; float aB = data.positions[indexB].a;
        float angularError = 0.0f, positionError = 0.0f;
        const bool fixedRotation = (invIA + invIB == 0.0f);
        if (enableLimit && limitState != INACTIVE_LIMIT && !fixedRotation) {
            const float angle = aB - aA - referenceAngle;
            float limitImpulse = 0.0f;
            if (limitState == EQUAL_LIMITS) {
                const float C = clampf(angle - lowerAngle, -kMaxAngularCorrection, kMaxAngularCorrection);
                limitImpulse = -motorMass * C;
                angularError = std::fabs(C);
            } else if (limitState == AT_LOWER_LIMIT) {
                float C = angle - lowerAngle;
                angularError = -C;
                C = clampf(C + kAngularSlop, -kMaxAngularCorrection, 0.0f);
                limitImpulse = -motorMass * C;
            } else if (limitState == AT_UPPER_LIMIT) {
                float C = angle - upperAngle;
                angularError = C;
                C = clampf(C - kAngularSlop, 0.0f, kMaxAngularCorrection);
                limitImpulse = -motorMass * C;
            }
            aA -= invIA * limitImpulse;
            aB += invIB * limitImpulse;
        }
        {
            const Rot qA(aA), qB(aB);
            const Vec2 ra = mul(qA, localAnchorA - localCenterA);
            const Vec2 rb = mul(qB, localAnchorB - localCenterB);
            const Vec2 C = cB + rb - cA - ra;
            positionError = C.length();
            const float mA = invMassA, mB = invMassB, iA = invIA, iB = invIB;
            Mat22 K;
            K.ex.x = mA + mB + iA * ra.y * ra.y + iB * rb.y * rb.y;
            K.ex.y = -iA * ra.x * ra.y - iB * rb.x * rb.y;
            K.ey.x = K.ex.y;
            K.ey.y = mA + mB + iA * ra.x * ra.x + iB * rb.x * rb.x;
            const Vec2 imp = -K.solve(C);
            cA -= mA * imp;
            aA -= iA * cross(ra, imp);
            cB += mB * imp;
            aB += iB * cross(rb, imp);
        }
        data.positions[indexA].c = cA; data.positions[indexA].a = aA;
        data.positions[indexB].c = cB; data.positions[indexB].a = aB;
        return positionError <= kLinearSlop && angularError <= kAngularSlop;
    }
};

// ------------------------------------------------------------------------------------------------
// contact solver (b2ContactSolver.cpp), two-body form
// ------------------------------------------------------------------------------------------------
struct VelocityConstraintPoint { Vec2 rA, rB; float normalImpulse, tangentImpulse, normalMass, tangentMass, velocityBias; };
struct ContactVelocityConstraint {
    VelocityConstraintPoint points[kMaxManifoldPoints];
    Vec2 normal;
    Mat22 normalMass, K;
    int indexA, indexB;
    float invMassA, invMassB, invIA, invIB, friction, restitution, tangentSpeed;
    int pointCount, contactIndex;
};
struct ContactPositionConstraint {
    Vec2 localPoints[kMaxManifoldPoints], localNormal, localPoint;
    int indexA, indexB;
    float invMassA, invMassB;
    Vec2 localCenterA, localCenterB;
    float invIA, invIB;
    Manifold::Type type;
    float radiusA, radiusB;
    int pointCount;
};

class ContactSolver {
    TimeStep step_;
    Position* positions_;
    Velocity* velocities_;
    std::vector<Contact*>& contacts_;
    std::vector<ContactVelocityConstraint> vcs_;
    std::vector<ContactPositionConstraint> pcs_;

    struct PositionSolverManifold {
        Vec2 normal, point;
        float separation;
        void initialize(const ContactPositionConstraint& pc, const Transform& xfA, const Transform& xfB, int index) {
            if (pc.type == Manifold::FACE_A) {
                normal = mul(xfA.q, pc.localNormal);
                const Vec2 planePoint = mul(xfA, pc.localPoint);
                const Vec2 clipPoint = mul(xfB, pc.localPoints[index]);
                separation = dot(clipPoint - planePoint, normal) - pc.radiusA - pc.radiusB;
                point = clipPoint;
            } else {
                normal = mul(xfB.q, pc.localNormal);
                const Vec2 planePoint = mul(xfB, pc.localPoint);
                const Vec2 clipPoint = mul(xfA, pc.localPoints[index]);
                separation = dot(clipPoint - planePoint, normal) - pc.radiusA - pc.radiusB;
                point = clipPoint;
                normal = -normal;
            }
        }
    };

public:
    ContactSolver(const TimeStep& step, std::vector<Contact*>& contacts, Position* positions, Velocity* velocities)
        : step_(step), positions_(positions), velocities_(velocities), contacts_(contacts) {
        const int count = (int)contacts.size();
        vcs_.resize(count); pcs_.resize(count);
        for (int i = 0; i < count; ++i) {
            Contact* contact = contacts[i];
            Fixture* fA = contact->fixtureA; Fixture* fB = contact->fixtureB;
            Body* bA = fA->body; Body* bB = fB->body;
            const Manifold& m = contact->manifold;
            ContactVelocityConstraint& vc = vcs_[i];
            vc.friction = contact->friction; vc.restitution = contact->restitution; vc.tangentSpeed = contact->tangentSpeed;
            vc.indexA = bA->islandIndex; vc.indexB = bB->islandIndex;
            vc.invMassA = bA->invMass; vc.invMassB = bB->invMass; vc.invIA = bA->invI; vc.invIB = bB->invI;
            vc.contactIndex = i; vc.pointCount = m.pointCount;
            vc.K = Mat22(); vc.normalMass = Mat22();
            ContactPositionConstraint& pc = pcs_[i];
            pc.indexA = vc.indexA; pc.indexB = vc.indexB;
            pc.invMassA = bA->invMass; pc.invMassB = bB->invMass;
            pc.localCenterA = bA->sweep.localCenter; pc.localCenterB = bB->sweep.localCenter;
            pc.invIA = bA->invI; pc.invIB = bB->invI;
            pc.localNormal = m.localNormal; pc.localPoint = m.localPoint;
            pc.pointCount = m.pointCount;
            pc.radiusA = fA->shape->radius; pc.radiusB = fB->shape->radius;
            pc.type = m.type;
            for (int j = 0; j < m.pointCount; ++j) {
                const ManifoldPoint& cp = m.points[j];
                VelocityConstraintPoint& vcp = vc.points[j];
                if (step_.warmStarting) {
                    vcp.normalImpulse = step_.dtRatio * cp.normalImpulse;
                    vcp.tangentImpulse = step_.dtRatio * cp.tangentImpulse;
                } else {
                    vcp.normalImpulse = 0.0f; vcp.tangentImpulse = 0.0f;
                }
                vcp.rA = Vec2(0.0f, 0.0f); vcp.rB = Vec2(0.0f, 0.0f);
                vcp.normalMass = 0.0f; vcp.tangentMass = 0.0f; vcp.velocityBias = 0.0f;
                pc.localPoints[j] = cp.localPoint;
            }
        }
    }
    void initializeVelocityConstraints() {
        for (size_t i = 0; i < vcs_.size(); ++i) {
            ContactVelocityConstraint& vc = vcs_[i];
            const ContactPositionConstraint& pc = pcs_[i];
            const float radiusA = pc.radiusA, radiusB = pc.radiusB;
            const Manifold& manifold = contacts_[vc.contactIndex]->manifold;
            const int indexA = vc.indexA, indexB = vc.indexB;
            const float mA = vc.invMassA, mB = vc.invMassB, iA = vc.invIA, iB = vc.invIB;
            const Vec2 localCenterA = pc.localCenterA, localCenterB = pc.localCenterB;
            const Vec2 cA = positions_[indexA].c; const float aA = positions_[indexA].a;
            const Vec2 vA = velocities_[indexA].v; const float wA = velocities_[indexA].w;
            const Vec2 cB = positions_[indexB].c; const float aB = positions_[indexB].a;
            const Vec2 vB = velocities_[indexB].v; const float wB = velocities_[indexB].w;
            Transform xfA, xfB;
            xfA.q.set(aA); xfB.q.set(aB);
            xfA.p = cA - mul(xfA.q, localCenterA);
            xfB.p = cB - mul(xfB.q, localCenterB);
            WorldManifold worldManifold;
            worldManifold.initialize(manifold, xfA, radiusA, xfB, radiusB);
            vc.normal = worldManifold.normal;
            const int pointCount = vc.pointCount;
            for (int j = 0; j < pointCount; ++j) {
                VelocityConstraintPoint& vcp = vc.points[j];
                vcp.rA = worldManifold.points[j] - cA;
                vcp.rB = worldManifold.points[j] - cB;
                const float rnA = cross(vcp.rA, vc.normal), rnB = cross(vcp.rB, vc.normal);
                const float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
                vcp.normalMass = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
                const Vec2 tangent = cross(vc.normal, 1.0f);
                const float rtA = cross(vcp.rA, tangent), rtB = cross(vcp.rB, tangent);
                const float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
                vcp.tangentMass = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
                vcp.velocityBias = 0.0f;
                const float vRel = dot(vc.normal, vB + cross(wB, vcp.rB) - vA - cross(wA, vcp.rA));
                if (vRel < -kVelocityThreshold) vcp.velocityBias = -vc.restitution * vRel;
            }
            if (vc.pointCount == 2) {
                const VelocityConstraintPoint& vcp1 = vc.points[0];
                const VelocityConstraintPoint& vcp2 = vc.points[1];
                const float rn1A = cross(vcp1.rA, vc.normal), rn1B = cross(vcp1.rB, vc.normal);
                const float rn2A = cross(vcp2.rA, vc.normal), rn2B = cross(vcp2.rB, vc.normal);
                const float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
                const float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
                const float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
                const float maxConditionNumber = 1000.0f;
                if (k11 * k11 < maxConditionNumber * (k11 * k22 - k12 * k12)) {
                    vc.K.ex = Vec2(k11, k12);
                    vc.K.ey = Vec2(k12, k22);
                    vc.normalMass = vc.K.inverse();
                } else {
                    vc.pointCount = 1;   // the constraints are redundant: use one
                }
            }
        }
    }
    void warmStart() {
        for (size_t i = 0; i < vcs_.size(); ++i) {
            ContactVelocityConstraint& vc = vcs_[i];
            const int indexA = vc.indexA, indexB = vc.indexB;
            const float mA = vc.invMassA, iA = vc.invIA, mB = vc.invMassB, iB = vc.invIB;
            Vec2 vA = velocities_[indexA].v; float wA = velocities_[indexA].w;
            Vec2 vB = velocities_[indexB].v; float wB = velocities_[indexB].w;
            const Vec2 normal = vc.normal;
            const Vec2 tangent = cross(normal, 1.0f);
            for (int j = 0; j < vc.pointCount; ++j) {
                const VelocityConstraintPoint& vcp = vc.points[j];
                const Vec2 P = vcp.normalImpulse * normal + vcp.tangentImpulse * tangent;
                wA -= iA * cross(vcp.rA, P);
                vA -= mA * P;
                wB += iB * cross(vcp.rB, P);
                vB += mB * P;
            }
            velocities_[indexA].v = vA; velocities_[indexA].w = wA;
            velocities_[indexB].v = vB; velocities_[indexB].w = wB;
        }
    }
    void solveVelocityConstraints() {
        for (size_t i = 0; i < vcs_.size(); ++i) {
            ContactVelocityConstraint& vc = vcs_[i];
            const int indexA = vc.indexA, indexB = vc.indexB;
            const float mA = vc.invMassA, iA = vc.invIA, mB = vc.invMassB, iB = vc.invIB;
            const int pointCount = vc.pointCount;
            Vec2 vA = velocities_[indexA].v; float wA = velocities_[indexA].w;
            Vec2 vB = velocities_[indexB].v; float wB = velocities_[indexB].w;
            const Vec2 normal = vc.normal;
            const Vec2 tangent = cross(normal, 1.0f);
            const float friction = vc.friction;
            // tangent constraints first: non-penetration matters more than friction
            for (int j = 0; j < pointCount; ++j) {
                VelocityConstraintPoint& vcp = vc.points[j];
                const Vec2 dv = vB + cross(wB, vcp.rB) - vA - cross(wA, vcp.rA);
                const float vt = dot(dv, tangent) - vc.tangentSpeed;
                float lambda = vcp.tangentMass * (-vt);
                const float maxFriction = friction * vcp.normalImpulse;
                const float newImpulse = clampf(vcp.tangentImpulse + lambda, -maxFriction, maxFriction);
                lambda = newImpulse - vcp.tangentImpulse;
                vcp.tangentImpulse = newImpulse;
                const Vec2 P = lambda * tangent;
                vA -= mA * P;
                wA -= iA * cross(vcp.rA, P);
                vB += mB * P;
                wB += iB * cross(vcp.rB, P);
            }
            if (vc.pointCount == 1) {
                VelocityConstraintPoint& vcp = vc.points[0];
                const Vec2 dv = vB + cross(wB, vcp.rB) - vA - cross(wA, vcp.rA);
                const float vn = dot(dv, normal);
                float lambda = -vcp.normalMass * (vn - vcp.velocityBias);
                const float newImpulse = std::max(vcp.normalImpulse + lambda, 0.0f);
                lambda = newImpulse - vcp.normalImpulse;
                vcp.normalImpulse = newImpulse;
                const Vec2 P = lambda * normal;
                vA -= mA * P;
                wA -= iA * cross(vcp.rA, P);
                vB += mB * P;
                wB += iB * cross(vcp.rB, P);
            } else {
                // block solver: the 2-point LCP by enumeration of its four cases
                VelocityConstraintPoint& cp1 = vc.points[0];
                VelocityConstraintPoint& cp2 = vc.points[1];
                const Vec2 a(cp1.normalImpulse, cp2.normalImpulse);
                const Vec2 dv1 = vB + cross(wB, cp1.rB) - vA - cross(wA, cp1.rA);
                const Vec2 dv2 = vB + cross(wB, cp2.rB) - vA - cross(wA, cp2.rA);
                float vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
                Vec2 b(vn1 - cp1.velocityBias, vn2 - cp2.velocityBias);
                b -= mul(vc.K, a);
                for (;;) {
                    Vec2 x = -mul(vc.normalMass, b);
                    if (x.x >= 0.0f && x.y >= 0.0f) { applyBlock(vc, x, a, normal, vA, wA, vB, wB); break; }
                    x.x = -cp1.normalMass * b.x; x.y = 0.0f;
                    vn1 = 0.0f; vn2 = vc.K.ex.y * x.x + b.y;
                    if (x.x >= 0.0f && vn2 >= 0.0f) { applyBlock(vc, x, a, normal, vA, wA, vB, wB); break; }
                    x.x = 0.0f; x.y = -cp2.normalMass * b.y;
                    vn1 = vc.K.ey.x * x.y + b.x; vn2 = 0.0f;
                    if (x.y >= 0.0f && vn1 >= 0.0f) { applyBlock(vc, x, a, normal, vA, wA, vB, wB); break; }
                    x.x = 0.0f; x.y = 0.0f;
                    vn1 = b.x; vn2 = b.y;
                    if (vn1 >= 0.0f && vn2 >= 0.0f) { applyBlock(vc, x, a, normal, vA, wA, vB, wB); break; }
                    break;   // no solution: give up (hit only through round-off)
                }
            }
            velocities_[indexA].v = vA; velocities_[indexA].w = wA;
            velocities_[indexB].v = vB; velocities_[indexB].w = wB;
        }
    }
    static void applyBlock(ContactVelocityConstraint& vc, const Vec2& x, const Vec2& a, const Vec2& normal, Vec2& vA, float& wA, Vec2& vB, float& wB) {
        VelocityConstraintPoint& cp1 = vc.points[0];
        VelocityConstraintPoint& cp2 = vc.points[1];
        const Vec2 d = x - a;
        const Vec2 P1 = d.x * normal, P2 = d.y * normal;
        vA -= vc.invMassA * (P1 + P2);
        wA -= vc.invIA * (cross(cp1.rA, P1) + cross(cp2.rA, P2));
        vB += vc.invMassB * (P1 + P2);
        wB += vc.invIB * (cross(cp1.rB, P1) + cross(cp2.rB, P2));
        cp1.normalImpulse = x.x;
        cp2.normalImpulse = x.y;
    }
    template <class F> void hashImpulses(F&& mix) const {
        for (const ContactVelocityConstraint& vc : vcs_) for (int j = 0; j < vc.pointCount; ++j) { mix(vc.points[j].normalImpulse); mix(vc.points[j].tangentImpulse); }
    }
    void storeImpulses() {
        for (size_t i = 0; i < vcs_.size(); ++i) {
            const ContactVelocityConstraint& vc = vcs_[i];
            Manifold& manifold = contacts_[vc.contactIndex]->manifold;
            for (int j = 0; j < vc.pointCount; ++j) {
                manifold.points[j].normalImpulse = vc.points[j].normalImpulse;
                manifold.points[j].tangentImpulse = vc.points[j].tangentImpulse;
            }
        }
    }
    bool solvePositionConstraints() {
        float minSeparation = 0.0f;
        for (size_t i = 0; i < pcs_.size(); ++i) {
            const ContactPositionConstraint& pc = pcs_[i];
            const int indexA = pc.indexA, indexB = pc.indexB;
            const Vec2 localCenterA = pc.localCenterA, localCenterB = pc.localCenterB;
            const float mA = pc.invMassA, iA = pc.invIA, mB = pc.invMassB, iB = pc.invIB;
            Vec2 cA = positions_[indexA].c; float aA = positions_[indexA].a;
            Vec2 cB = positions_[indexB].c; float aB = positions_[indexB].a;
            for (int j = 0; j < pc.pointCount; ++j) {
                Transform xfA, xfB;
                xfA.q.set(aA); xfB.q.set(aB);
                xfA.p = cA - mul(xfA.q, localCenterA);
                xfB.p = cB - mul(xfB.q, localCenterB);
                PositionSolverManifold psm;
                psm.initialize(pc, xfA, xfB, j);
                const Vec2 normal = psm.normal, point = psm.point;
                const float separation = psm.separation;
                const Vec2 rA = point - cA, rB = point - cB;
                minSeparation = std::min(minSeparation, separation);
                const float C = clampf(kBaumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
                const float rnA = cross(rA, normal), rnB = cross(rB, normal);
                const float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
                const float imp = K > 0.0f ? -C / K : 0.0f;
                const Vec2 P = imp * normal;
                cA -= mA * P;
                aA -= iA * cross(rA, P);
                cB += mB * P;
                aB += iB * cross(rB, P);
            }
            positions_[indexA].c = cA; positions_[indexA].a = aA;
            positions_[indexB].c = cB; positions_[indexB].a = aB;
        }
        return minSeparation >= -3.0f * kLinearSlop;
    }
    // sequential position solver for the TOI sub-step: only the two TOI bodies move
    bool solveTOIPositionConstraints(int toiIndexA, int toiIndexB) {
        float minSeparation = 0.0f;
        for (size_t i = 0; i < pcs_.size(); ++i) {
            const ContactPositionConstraint& pc = pcs_[i];
            const int indexA = pc.indexA, indexB = pc.indexB;
            const Vec2 localCenterA = pc.localCenterA, localCenterB = pc.localCenterB;
            float mA = 0.0f, iA = 0.0f;
            if (indexA == toiIndexA || indexA == toiIndexB) { mA = pc.invMassA; iA = pc.invIA; }
            float mB = 0.0f, iB = 0.0f;
            if (indexB == toiIndexA || indexB == toiIndexB) { mB = pc.invMassB; iB = pc.invIB; }
            Vec2 cA = positions_[indexA].c; float aA = positions_[indexA].a;
            Vec2 cB = positions_[indexB].c; float aB = positions_[indexB].a;
            for (int j = 0; j < pc.pointCount; ++j) {
                Transform xfA, xfB;
                xfA.q.set(aA); xfB.q.set(aB);
                xfA.p = cA - mul(xfA.q, localCenterA);
                xfB.p = cB - mul(xfB.q, localCenterB);
                PositionSolverManifold psm;
                psm.initialize(pc, xfA, xfB, j);
                const Vec2 normal = psm.normal, point = psm.point;
                const float separation = psm.separation;
                const Vec2 rA = point - cA, rB = point - cB;
                minSeparation = std::min(minSeparation, separation);
                const float C = clampf(kToiBaumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
                const float rnA = cross(rA, normal), rnB = cross(rB, normal);
                const float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
                const float imp = K > 0.0f ? -C / K : 0.0f;
                const Vec2 P = imp * normal;
                cA -= mA * P;
                aA -= iA * cross(rA, P);
                cB += mB * P;
                aB += iB * cross(rB, P);
            }
            positions_[indexA].c = cA; positions_[indexA].a = aA;
            positions_[indexB].c = cB; positions_[indexB].a = aB;
        }
        return minSeparation >= -1.5f * kLinearSlop;
    }
};

// ------------------------------------------------------------------------------------------------
// world
// ------------------------------------------------------------------------------------------------
struct WorldOptions {
    // What a `false` returned by ContactManager.BeginContact does (LunarLanderEnv.cs:330 always returns false).
    //   0  ignored: the contact is solved from the step it begins to touch
    //   1  `Enabled = BeginContact(c)` for the step of the begin-touch transition only; the contact stays
    //      "touching", Update re-enables it at the top of the next step (Box2D's SetEnabled semantics)
    //   2  additionally `if (!Enabled) IsTouching = false` (Farseer 3.5's Contact.Update): BeginContact fires
    //      again on every later step and the contact is never solved
    int beginContactFalse = 0;
    bool continuousPhysics = false;     // the TOI pass (Settings.ContinuousPhysics)
    bool reverseSeedOrder = false;      // island seeds taken from the END of the body list (Farseer's World.Solve loop)
    bool contactListHeadInsertion = false;  // world contact list order for Collide(): false = creation order
    bool warmStarting = true, allowSleep = true;
    int toiVelocityIterations = 8, toiPositionIterations = 20;   // Settings.TOIVelocityIterations / TOIPositionIterations
    bool findContactsOnSetTransform = true;
    // Diagnostic only: ignore creation order and take contacts by (dynamic body's userIndex, other fixture's userIndex) ascending
    // in Collide() and inside an island -- the fixed order gym.net_b200 used before it tracked contact creation order.
    bool canonicalContactOrder = false;   // Farseer's Body.SetTransform ends with ContactManager.FindNewContacts()
    SinCosFn sincos = sincos_libm_double;
};

class World {
public:
    Vec2 gravity;
    WorldOptions opt;
    std::vector<std::unique_ptr<Body>> bodies;          // creation order (World.BodyList)
    std::vector<std::unique_ptr<Joint>> joints;         // creation order
    std::vector<std::unique_ptr<Contact>> contactStore;
    std::vector<Contact*> contactList;                  // world contact list, creation order
    std::function<bool(Contact*)> beginContact;         // ContactManager.BeginContact
    std::function<void(Contact*)> endContact;           // ContactManager.EndContact
    float inv_dt0 = 0.0f;
    bool newFixture = false, stepComplete = true;
    uint64_t contactSerial = 0;
    // statistics of the last step (for tests / reports)
    int lastToiEvents = 0;
    // analysis hook (tools/lunar_cycle_study.py): when set, the island solver records after which velocity / position iteration
    // the iterated state first repeats with period 1..4 (an exact fixed point or limit cycle of the float32 Gauss-Seidel sweep)
    bool recordCycles = false;
    int velCycleAt = -1, velCyclePeriod = 0, posCycleAt = -1, posCyclePeriod = 0, posIterationsRun = 0;

    explicit World(const Vec2& g, const WorldOptions& o = WorldOptions()) : gravity(g), opt(o) {}

    struct SinCosScope {   // bodies and joints built outside step() must see the same sin/cos provider
        SinCosFn saved;
        explicit SinCosScope(SinCosFn f) : saved(current_sincos()) { current_sincos() = f; }
        ~SinCosScope() { current_sincos() = saved; }
    };

    Body* createBody() {
        bodies.emplace_back(new Body());
        Body* b = bodies.back().get();
        b->world = this;
        return b;
    }
    template <class J> J* addJoint(std::unique_ptr<J> j) {
        J* raw = j.get();
        joints.emplace_back(std::move(j));
        // connect to the bodies' doubly linked lists: head insertion
        raw->edgeA.joint = raw; raw->edgeA.other = raw->bodyB; raw->edgeA.prev = nullptr; raw->edgeA.next = raw->bodyA->jointList;
        if (raw->bodyA->jointList) raw->bodyA->jointList->prev = &raw->edgeA;
        raw->bodyA->jointList = &raw->edgeA;
        raw->edgeB.joint = raw; raw->edgeB.other = raw->bodyA; raw->edgeB.prev = nullptr; raw->edgeB.next = raw->bodyB->jointList;
        if (raw->bodyB->jointList) raw->bodyB->jointList->prev = &raw->edgeB;
        raw->bodyB->jointList = &raw->edgeB;
        // (contacts between the two bodies would be flagged for filtering here; none can exist yet)
        return raw;
    }

    // ---- broad phase: fat AABB per fixture proxy, move buffer, sorted pair creation ---------------
    struct Proxy { AABB fat; Fixture* fixture; };
    std::vector<Proxy> proxies;
    std::vector<int> moveBuffer;
    int createProxy(Fixture* f, const AABB& aabb) {
        const Vec2 r(kAabbExtension, kAabbExtension);
        proxies.push_back(Proxy{AABB{aabb.lo - r, aabb.hi + r}, f});
        const int id = (int)proxies.size() - 1;
        moveBuffer.push_back(id);
        return id;
    }
    void moveProxy(int id, const AABB& aabb, const Vec2& displacement) {
        if (proxies[id].fat.contains(aabb)) return;
        const Vec2 r(kAabbExtension, kAabbExtension);
        AABB b{aabb.lo - r, aabb.hi + r};
        const Vec2 d = kAabbMultiplier * displacement;   // predict motion
        if (d.x < 0.0f) b.lo.x += d.x; else b.hi.x += d.x;
        if (d.y < 0.0f) b.lo.y += d.y; else b.hi.y += d.y;
        proxies[id].fat = b;
        moveBuffer.push_back(id);
    }
    void touchProxy(int id) { moveBuffer.push_back(id); }

    void findNewContacts() {
        std::vector<std::pair<int, int>> pairs;
        for (int moved : moveBuffer) {
            for (int other = 0; other < (int)proxies.size(); ++other) {
                if (other == moved) continue;
                if (!overlap(proxies[moved].fat, proxies[other].fat)) continue;
                pairs.emplace_back(std::min(moved, other), std::max(moved, other));
            }
        }
        moveBuffer.clear();
        std::sort(pairs.begin(), pairs.end());
        pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
        for (const auto& pr : pairs) addPair(proxies[pr.first].fixture, proxies[pr.second].fixture);
    }

    static bool shouldCollideFixtures(const Fixture* a, const Fixture* b) {
        if (a->group == b->group && a->group != 0) return a->group > 0;
        return (a->collidesWith & b->category) != 0 && (a->category & b->collidesWith) != 0;
    }

    void addPair(Fixture* fixtureA, Fixture* fixtureB) {
        Body* bodyA = fixtureA->body; Body* bodyB = fixtureB->body;
        if (bodyA == bodyB) return;
        for (ContactEdge* edge = bodyB->contactList; edge; edge = edge->next) {
            if (edge->other == bodyA) {
                Fixture* fA = edge->contact->fixtureA; Fixture* fB = edge->contact->fixtureB;
                if ((fA == fixtureA && fB == fixtureB) || (fA == fixtureB && fB == fixtureA)) return;   // already exists
            }
        }
        if (!bodyB->shouldCollide(bodyA)) return;
        if (!shouldCollideFixtures(fixtureA, fixtureB)) return;
        // Contact.Create: the registered order of a mixed pair puts the edge first
        const Shape::Type t1 = fixtureA->shape->type, t2 = fixtureB->shape->type;
        if (t1 == Shape::POLYGON && t2 == Shape::EDGE) std::swap(fixtureA, fixtureB);
        else if (!(t1 == Shape::EDGE && t2 == Shape::POLYGON)) return;   // pair types this oracle has no narrow phase for (none occur)
        contactStore.emplace_back(new Contact());
        Contact* c = contactStore.back().get();
        c->fixtureA = fixtureA; c->fixtureB = fixtureB;
        c->friction = std::sqrt(fixtureA->friction * fixtureB->friction);                 // Settings.MixFriction
        c->restitution = fixtureA->restitution > fixtureB->restitution ? fixtureA->restitution : fixtureB->restitution;
        c->serial = contactSerial++;
        bodyA = fixtureA->body; bodyB = fixtureB->body;
        contactList.push_back(c);
        c->nodeA.contact = c; c->nodeA.other = bodyB; c->nodeA.prev = nullptr; c->nodeA.next = bodyA->contactList;
        if (bodyA->contactList) bodyA->contactList->prev = &c->nodeA;
        bodyA->contactList = &c->nodeA;
        c->nodeB.contact = c; c->nodeB.other = bodyA; c->nodeB.prev = nullptr; c->nodeB.next = bodyB->contactList;
        if (bodyB->contactList) bodyB->contactList->prev = &c->nodeB;
        bodyB->contactList = &c->nodeB;
        // (Box2D wakes both bodies here unless one fixture is a sensor)
        bodyA->setAwake(true);
        bodyB->setAwake(true);
    }

    void destroyContact(Contact* c) {
        Body* bodyA = c->fixtureA->body; Body* bodyB = c->fixtureB->body;
        if (endContact && c->touching) endContact(c);
        contactList.erase(std::find(contactList.begin(), contactList.end(), c));
        if (c->nodeA.prev) c->nodeA.prev->next = c->nodeA.next;
        if (c->nodeA.next) c->nodeA.next->prev = c->nodeA.prev;
        if (&c->nodeA == bodyA->contactList) bodyA->contactList = c->nodeA.next;
        if (c->nodeB.prev) c->nodeB.prev->next = c->nodeB.next;
        if (c->nodeB.next) c->nodeB.next->prev = c->nodeB.prev;
        if (&c->nodeB == bodyB->contactList) bodyB->contactList = c->nodeB.next;
        for (auto it = contactStore.begin(); it != contactStore.end(); ++it) if (it->get() == c) { contactStore.erase(it); break; }
    }

    static bool canonicalLess(const Contact* a, const Contact* b) {
        const int ka = a->fixtureB->body->userIndex * 16 + a->fixtureA->userIndex, kb = b->fixtureB->body->userIndex * 16 + b->fixtureA->userIndex;
        return ka < kb;
    }

    // Contact.Update
    void updateContact(Contact* c) {
        const Manifold oldManifold = c->manifold;
        c->enabled = true;   // re-enable this contact
        bool touching = false;
        const bool wasTouching = c->touching;
        Body* bodyA = c->fixtureA->body; Body* bodyB = c->fixtureB->body;
        c->evaluate(&c->manifold, bodyA->xf, bodyB->xf);
        touching = c->manifold.pointCount > 0;
        // match old contact ids to new contact ids and copy the stored impulses to warm start the solver
        for (int i = 0; i < c->manifold.pointCount; ++i) {
            ManifoldPoint& mp2 = c->manifold.points[i];
            mp2.normalImpulse = 0.0f; mp2.tangentImpulse = 0.0f;
            const uint32_t key2 = mp2.id.key();
            for (int j = 0; j < oldManifold.pointCount; ++j) {
                const ManifoldPoint& mp1 = oldManifold.points[j];
                if (mp1.id.key() == key2) { mp2.normalImpulse = mp1.normalImpulse; mp2.tangentImpulse = mp1.tangentImpulse; break; }
            }
        }
        if (touching != wasTouching) { bodyA->setAwake(true); bodyB->setAwake(true); }
        c->touching = touching;
        if (!wasTouching && touching) {
            if (beginContact) {
                const bool ret = beginContact(c);
                if (opt.beginContactFalse >= 1) c->enabled = ret;
                if (opt.beginContactFalse >= 2 && !c->enabled) c->touching = false;
            }
        }
        if (wasTouching && !touching) { if (endContact) endContact(c); }
    }

    void collide() {
        // (iterating over a snapshot: contacts may be destroyed inside the loop)
        std::vector<Contact*> order(contactList);
        if (opt.contactListHeadInsertion) std::reverse(order.begin(), order.end());
        if (opt.canonicalContactOrder) std::sort(order.begin(), order.end(), canonicalLess);
        for (Contact* c : order) {
            Fixture* fA = c->fixtureA; Fixture* fB = c->fixtureB;
            Body* bA = fA->body; Body* bB = fB->body;
            const bool activeA = bA->awake && bA->type != STATIC_BODY;
            const bool activeB = bB->awake && bB->type != STATIC_BODY;
            if (!activeA && !activeB) continue;   // at least one body must be awake and dynamic or kinematic
            if (!overlap(proxies[fA->proxyId].fat, proxies[fB->proxyId].fat)) { destroyContact(c); continue; }
            updateContact(c);
        }
    }

    // ---- b2Island::Solve ----------------------------------------------------------------------------
    struct Island {
        std::vector<Body*> bodies;
        std::vector<Contact*> contacts;
        std::vector<Joint*> joints;
        std::vector<Position> positions;
        std::vector<Velocity> velocities;
        void clear() { bodies.clear(); contacts.clear(); joints.clear(); }
        void add(Body* b) { b->islandIndex = (int)bodies.size(); bodies.push_back(b); }
    };

    void solveIsland(Island& island, const TimeStep& step) {
        const float h = step.dt;
        const int n = (int)island.bodies.size();
        island.positions.resize(n); island.velocities.resize(n);
        for (int i = 0; i < n; ++i) {
            Body* b = island.bodies[i];
            const Vec2 c = b->sweep.c; const float a = b->sweep.a;
            Vec2 v = b->linearVelocity; float w = b->angularVelocity;
            b->sweep.c0 = b->sweep.c; b->sweep.a0 = b->sweep.a;   // store positions for continuous collision
            if (b->type == DYNAMIC_BODY) {
                v += h * (b->gravityScale * gravity + b->invMass * b->force);
                w += h * b->invI * b->torque;
                v *= 1.0f / (1.0f + h * b->linearDamping);
                w *= 1.0f / (1.0f + h * b->angularDamping);
            }
            island.positions[i].c = c; island.positions[i].a = a;
            island.velocities[i].v = v; island.velocities[i].w = w;
        }
        SolverData solverData{step, island.positions.data(), island.velocities.data()};
        ContactSolver contactSolver(step, island.contacts, island.positions.data(), island.velocities.data());
        contactSolver.initializeVelocityConstraints();
        if (step.warmStarting) contactSolver.warmStart();
        for (Joint* j : island.joints) j->initVelocityConstraints(solverData);
        std::vector<uint64_t> hist;
        velCycleAt = -1; velCyclePeriod = 0;
        for (int i = 0; i < step.velocityIterations; ++i) {
            for (Joint* j : island.joints) j->solveVelocityConstraints(solverData);
            contactSolver.solveVelocityConstraints();
            if (recordCycles && velCycleAt < 0) {
                uint64_t hsh = 1469598103934665603ull;
                auto mix = [&hsh](float f) { uint32_t u; std::memcpy(&u, &f, 4); hsh = (hsh ^ u) * 1099511628211ull; };
                for (int b = 0; b < n; ++b) { mix(island.velocities[b].v.x); mix(island.velocities[b].v.y); mix(island.velocities[b].w); }
                for (Joint* j : island.joints) { RevoluteJoint* r = static_cast<RevoluteJoint*>(j); mix(r->impulse.x); mix(r->impulse.y); mix(r->impulse.z); mix(r->motorImpulse); }
                contactSolver.hashImpulses(mix);
                for (int p = 1; p <= 4 && p <= (int)hist.size(); ++p) if (hist[hist.size() - p] == hsh) { velCycleAt = i; velCyclePeriod = p; break; }
                hist.push_back(hsh);
            }
        }
        contactSolver.storeImpulses();
        for (int i = 0; i < n; ++i) {
            Vec2 c = island.positions[i].c; float a = island.positions[i].a;
            Vec2 v = island.velocities[i].v; float w = island.velocities[i].w;
            const Vec2 translation = h * v;
            if (dot(translation, translation) > kMaxTranslationSquared) { const float ratio = kMaxTranslation / translation.length(); v *= ratio; }
            const float rotation = h * w;
            if (rotation * rotation > kMaxRotationSquared) { const float ratio = kMaxRotation / std::fabs(rotation); w *= ratio; }
            c += h * v;
            a += h * w;
            island.positions[i].c = c; island.positions[i].a = a;
            island.velocities[i].v = v; island.velocities[i].w = w;
        }
        bool positionSolved = false;
        std::vector<uint64_t> phist;
        posCycleAt = -1; posCyclePeriod = 0; posIterationsRun = 0;
        for (int i = 0; i < step.positionIterations; ++i) {
            const bool contactsOkay = contactSolver.solvePositionConstraints();
            bool jointsOkay = true;
            for (Joint* j : island.joints) { const bool jointOkay = j->solvePositionConstraints(solverData); jointsOkay = jointsOkay && jointOkay; }
            posIterationsRun = i + 1;
            if (contactsOkay && jointsOkay) { positionSolved = true; break; }
            if (recordCycles && posCycleAt < 0) {
                uint64_t hsh = 1469598103934665603ull;
                auto mix = [&hsh](float f) { uint32_t u; std::memcpy(&u, &f, 4); hsh = (hsh ^ u) * 1099511628211ull; };
                for (int b = 0; b < n; ++b) { mix(island.positions[b].c.x); mix(island.positions[b].c.y); mix(island.positions[b].a); }
                for (int p = 1; p <= 4 && p <= (int)phist.size(); ++p) if (phist[phist.size() - p] == hsh) { posCycleAt = i; posCyclePeriod = p; break; }
                phist.push_back(hsh);
            }
        }
        for (int i = 0; i < n; ++i) {
            Body* b = island.bodies[i];
            b->sweep.c = island.positions[i].c; b->sweep.a = island.positions[i].a;
            b->linearVelocity = island.velocities[i].v; b->angularVelocity = island.velocities[i].w;
            b->synchronizeTransform();
        }
        if (opt.allowSleep) {
            float minSleepTime = kMaxFloat;
            const float linTolSqr = kLinearSleepTolerance * kLinearSleepTolerance, angTolSqr = kAngularSleepTolerance * kAngularSleepTolerance;
            for (Body* b : island.bodies) {
                if (b->type == STATIC_BODY) continue;
                if (!b->autoSleep || b->angularVelocity * b->angularVelocity > angTolSqr || dot(b->linearVelocity, b->linearVelocity) > linTolSqr) {
                    b->sleepTime = 0.0f; minSleepTime = 0.0f;
                } else {
                    b->sleepTime += h; minSleepTime = std::min(minSleepTime, b->sleepTime);
                }
            }
            if (minSleepTime >= kTimeToSleep && positionSolved) for (Body* b : island.bodies) b->setAwake(false);
        }
    }

    void solve(const TimeStep& step) {
        for (auto& b : bodies) b->islandFlag = false;
        for (Contact* c : contactList) c->islandFlag = false;
        for (auto& j : joints) j->islandFlag = false;
        Island island;
        std::vector<Body*> stack;
        const int nb = (int)bodies.size();
        for (int k = 0; k < nb; ++k) {
            Body* seed = bodies[opt.reverseSeedOrder ? nb - 1 - k : k].get();
            if (seed->islandFlag) continue;
            if (!seed->awake || !seed->enabled) continue;
            if (seed->type == STATIC_BODY) continue;
            island.clear();
            stack.clear();
            stack.push_back(seed);
            seed->islandFlag = true;
            while (!stack.empty()) {   // depth-first search on the constraint graph
                Body* b = stack.back(); stack.pop_back();
                island.add(b);
                b->setAwake(true);
                if (b->type == STATIC_BODY) continue;   // islands do not propagate across static bodies
                for (ContactEdge* ce = b->contactList; ce; ce = ce->next) {
                    Contact* contact = ce->contact;
                    if (contact->islandFlag) continue;
                    if (!contact->enabled || !contact->touching) continue;
                    island.contacts.push_back(contact);
                    contact->islandFlag = true;
                    Body* other = ce->other;
                    if (other->islandFlag) continue;
                    stack.push_back(other);
                    other->islandFlag = true;
                }
                for (JointEdge* je = b->jointList; je; je = je->next) {
                    if (je->joint->islandFlag) continue;
                    Body* other = je->other;
                    if (!other->enabled) continue;
                    island.joints.push_back(je->joint);
                    je->joint->islandFlag = true;
                    if (other->islandFlag) continue;
                    stack.push_back(other);
                    other->islandFlag = true;
                }
            }
            if (opt.canonicalContactOrder) std::sort(island.contacts.begin(), island.contacts.end(), canonicalLess);
            solveIsland(island, step);
            for (Body* b : island.bodies) if (b->type == STATIC_BODY) b->islandFlag = false;   // static bodies may join other islands
        }
        for (auto& b : bodies) {
            if (!b->islandFlag) continue;
            if (b->type == STATIC_BODY) continue;
            b->synchronizeFixtures();
        }
        findNewContacts();
    }

    // ---- b2World::SolveTOI + b2Island::SolveTOI ------------------------------------------------------
    void solveTOIIsland(Island& island, const TimeStep& subStep, int toiIndexA, int toiIndexB) {
        const int n = (int)island.bodies.size();
        island.positions.resize(n); island.velocities.resize(n);
        for (int i = 0; i < n; ++i) {
            Body* b = island.bodies[i];
            island.positions[i].c = b->sweep.c; island.positions[i].a = b->sweep.a;
            island.velocities[i].v = b->linearVelocity; island.velocities[i].w = b->angularVelocity;
        }
        ContactSolver contactSolver(subStep, island.contacts, island.positions.data(), island.velocities.data());
        for (int i = 0; i < subStep.positionIterations; ++i) if (contactSolver.solveTOIPositionConstraints(toiIndexA, toiIndexB)) break;
        // leap of faith to the new safe state
        island.bodies[toiIndexA]->sweep.c0 = island.positions[toiIndexA].c; island.bodies[toiIndexA]->sweep.a0 = island.positions[toiIndexA].a;
        island.bodies[toiIndexB]->sweep.c0 = island.positions[toiIndexB].c; island.bodies[toiIndexB]->sweep.a0 = island.positions[toiIndexB].a;
        // no warm starting: the impulses of the first pass are stale
        contactSolver.initializeVelocityConstraints();
        for (int i = 0; i < subStep.velocityIterations; ++i) contactSolver.solveVelocityConstraints();
        // impulses are not stored for the next step
        const float h = subStep.dt;
        for (int i = 0; i < n; ++i) {
            Vec2 c = island.positions[i].c; float a = island.positions[i].a;
            Vec2 v = island.velocities[i].v; float w = island.velocities[i].w;
            const Vec2 translation = h * v;
            if (dot(translation, translation) > kMaxTranslationSquared) { const float ratio = kMaxTranslation / translation.length(); v *= ratio; }
            const float rotation = h * w;
            if (rotation * rotation > kMaxRotationSquared) { const float ratio = kMaxRotation / std::fabs(rotation); w *= ratio; }
            c += h * v;
            a += h * w;
            island.positions[i].c = c; island.positions[i].a = a;
            island.velocities[i].v = v; island.velocities[i].w = w;
            Body* body = island.bodies[i];
            body->sweep.c = c; body->sweep.a = a;
            body->linearVelocity = v; body->angularVelocity = w;
            body->synchronizeTransform();
        }
    }

    void solveTOI(const TimeStep& step) {
        Island island;
        if (stepComplete) {
            for (auto& b : bodies) { b->islandFlag = false; b->sweep.alpha0 = 0.0f; }
            for (Contact* c : contactList) { c->toiFlag = false; c->islandFlag = false; c->toiCount = 0; c->toi = 1.0f; }
        }
        for (;;) {   // find TOI events and solve them
            Contact* minContact = nullptr;
            float minAlpha = 1.0f;
            for (Contact* c : contactList) {
                if (!c->enabled) continue;                 // disabled by the user
                if (c->toiCount > kMaxSubSteps) continue;  // prevent excessive sub-stepping
                float alpha = 1.0f;
                if (c->toiFlag) {
                    alpha = c->toi;   // cached
                } else {
                    Fixture* fA = c->fixtureA; Fixture* fB = c->fixtureB;
                    Body* bA = fA->body; Body* bB = fB->body;
                    const BodyType typeA = bA->type, typeB = bB->type;
                    const bool activeA = bA->awake && typeA != STATIC_BODY;
                    const bool activeB = bB->awake && typeB != STATIC_BODY;
                    if (!activeA && !activeB) continue;   // is at least one body active (awake and dynamic or kinematic)?
                    const bool collideA = bA->bullet || typeA != DYNAMIC_BODY;
                    const bool collideB = bB->bullet || typeB != DYNAMIC_BODY;
                    if (!collideA && !collideB) continue;   // are these two non-bullet dynamic bodies?
                    // put the sweeps onto the same time interval
                    float alpha0 = bA->sweep.alpha0;
                    if (bA->sweep.alpha0 < bB->sweep.alpha0) { alpha0 = bB->sweep.alpha0; bA->sweep.advance(alpha0); }
                    else if (bB->sweep.alpha0 < bA->sweep.alpha0) { alpha0 = bA->sweep.alpha0; bB->sweep.advance(alpha0); }
                    TOIInput input;
                    input.proxyA.set(fA->shape.get());
                    input.proxyB.set(fB->shape.get());
                    input.sweepA = bA->sweep; input.sweepB = bB->sweep;
                    input.tMax = 1.0f;
                    TOIOutput output;
                    timeOfImpact(&output, &input);
                    const float beta = output.t;   // fraction of the remaining part of the step
                    if (output.state == TOIOutput::TOUCHING) alpha = std::min(alpha0 + (1.0f - alpha0) * beta, 1.0f);
                    else alpha = 1.0f;
                    c->toi = alpha;
                    c->toiFlag = true;
                }
                if (alpha < minAlpha) { minContact = c; minAlpha = alpha; }
            }
            if (minContact == nullptr || 1.0f - 10.0f * kEpsilon < minAlpha) { stepComplete = true; break; }   // no more TOI events
            // advance the bodies to the TOI
            Fixture* fA = minContact->fixtureA; Fixture* fB = minContact->fixtureB;
            Body* bA = fA->body; Body* bB = fB->body;
            const Sweep backup1 = bA->sweep, backup2 = bB->sweep;
            bA->advance(minAlpha);
            bB->advance(minAlpha);
            updateContact(minContact);   // the TOI contact likely has some new contact points
            minContact->toiFlag = false;
            ++minContact->toiCount;
            if (!minContact->enabled || !minContact->touching) {   // is the contact solid?
                minContact->enabled = false;
                bA->sweep = backup1; bB->sweep = backup2;
                bA->synchronizeTransform(); bB->synchronizeTransform();
                continue;
            }
            ++lastToiEvents;
            bA->setAwake(true);
            bB->setAwake(true);
            island.clear();
            island.add(bA);
            island.add(bB);
            island.contacts.push_back(minContact);
            bA->islandFlag = true; bB->islandFlag = true;
            minContact->islandFlag = true;
            Body* pair[2] = {bA, bB};
            for (int i = 0; i < 2; ++i) {   // contacts of the two bodies
                Body* body = pair[i];
                if (body->type != DYNAMIC_BODY) continue;
                for (ContactEdge* ce = body->contactList; ce; ce = ce->next) {
                    Contact* contact = ce->contact;
                    if (contact->islandFlag) continue;
                    Body* other = ce->other;
                    if (other->type == DYNAMIC_BODY && !body->bullet && !other->bullet) continue;   // only static, kinematic or bullet bodies join
                    const Sweep backup = other->sweep;   // tentatively advance the body to the TOI
                    if (!other->islandFlag) other->advance(minAlpha);
                    updateContact(contact);
                    if (!contact->enabled) { other->sweep = backup; other->synchronizeTransform(); continue; }
                    if (!contact->touching) { other->sweep = backup; other->synchronizeTransform(); continue; }
                    contact->islandFlag = true;
                    island.contacts.push_back(contact);
                    if (other->islandFlag) continue;
                    other->islandFlag = true;
                    if (other->type != STATIC_BODY) other->setAwake(true);
                    island.add(other);
                }
            }
            TimeStep subStep;
            subStep.dt = (1.0f - minAlpha) * step.dt;
            subStep.inv_dt = 1.0f / subStep.dt;
            subStep.dtRatio = 1.0f;
            subStep.positionIterations = opt.toiPositionIterations;
            subStep.velocityIterations = opt.toiVelocityIterations;
            subStep.warmStarting = false;
            solveTOIIsland(island, subStep, bA->islandIndex, bB->islandIndex);
            for (Body* body : island.bodies) {   // reset island flags and synchronise broad-phase proxies
                body->islandFlag = false;
                if (body->type != DYNAMIC_BODY) continue;
                body->synchronizeFixtures();
                for (ContactEdge* ce = body->contactList; ce; ce = ce->next) { ce->contact->toiFlag = false; ce->contact->islandFlag = false; }   // invalidate all contact TOIs on this displaced body
            }
            findNewContacts();   // commit fixture proxy movements; also some contacts can be destroyed
        }
    }

    // World.Step(dt, ref SolverIterations)
    void step(float dt, int velocityIterations, int positionIterations) {
        SinCosScope scope(opt.sincos);
        lastToiEvents = 0;
        if (newFixture) { findNewContacts(); newFixture = false; }
        TimeStep step;
        step.dt = dt;
        step.velocityIterations = velocityIterations;
        step.positionIterations = positionIterations;
        step.inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
        step.dtRatio = inv_dt0 * dt;
        step.warmStarting = opt.warmStarting;
        collide();   // update contacts; this is where some contacts are destroyed
        if (stepComplete && step.dt > 0.0f) solve(step);
        if (opt.continuousPhysics && step.dt > 0.0f) solveTOI(step);
        if (step.dt > 0.0f) inv_dt0 = step.inv_dt;
        for (auto& b : bodies) { b->force = Vec2(0.0f, 0.0f); b->torque = 0.0f; }   // ClearForces (AutoClearForces)
    }
};

// ---- Body members that need World ------------------------------------------------------------------
inline void Body::setTransform(const Vec2& position, float angle) {
    World::SinCosScope scope(world->opt.sincos);
    xf.q.set(angle);
    xf.p = position;
    sweep.c = mul(xf, sweep.localCenter);
    sweep.a = angle;
    sweep.c0 = sweep.c;
    sweep.a0 = angle;
    for (auto& f : fixtures) {
        const AABB aabb = f->shape->computeAABB(xf);
        world->moveProxy(f->proxyId, aabb, Vec2(0.0f, 0.0f));
    }
    if (world->opt.findContactsOnSetTransform) world->findNewContacts();
}
inline void Body::setType(BodyType t) {
    if (type == t) return;
    type = t;
    resetMassData();
    if (type == STATIC_BODY) {
        linearVelocity = Vec2(0.0f, 0.0f); angularVelocity = 0.0f;
        sweep.a0 = sweep.a; sweep.c0 = sweep.c;
        synchronizeFixtures();
    }
    setAwake(true);
    force = Vec2(0.0f, 0.0f); torque = 0.0f;
    // (existing contacts of the body would be deleted here; a body changes type before it has any)
    for (auto& f : fixtures) world->touchProxy(f->proxyId);
}
inline Fixture* Body::createFixture(std::unique_ptr<Shape> shape) {
    World::SinCosScope scope(world->opt.sincos);
    fixtures.emplace_back(new Fixture());
    Fixture* f = fixtures.back().get();
    f->body = this;
    f->shape = std::move(shape);
    f->proxyId = world->createProxy(f, f->shape->computeAABB(xf));
    if (f->shape->density > 0.0f) resetMassData();
    world->newFixture = true;
    return f;
}
inline void Body::resetMassData() {
    mass = 0.0f; invMass = 0.0f; inertia = 0.0f; invI = 0.0f;
    sweep.localCenter = Vec2(0.0f, 0.0f);
    if (type == STATIC_BODY || type == KINEMATIC_BODY) {
        sweep.c0 = xf.p; sweep.c = xf.p; sweep.a0 = sweep.a;
        return;
    }
    Vec2 localCenter(0.0f, 0.0f);
    for (auto& f : fixtures) {
        if (f->shape->density == 0.0f) continue;
        const MassData md = f->shape->massData();
        mass += md.mass;
        localCenter += md.mass * md.center;
        inertia += md.inertia;
    }
    if (mass > 0.0f) { invMass = 1.0f / mass; localCenter *= invMass; }
    else { mass = 1.0f; invMass = 1.0f; }   // dynamic bodies are forced to have a positive mass
    if (inertia > 0.0f && !fixedRotation) {
        inertia -= mass * dot(localCenter, localCenter);   // centre the inertia about the centre of mass
        invI = 1.0f / inertia;
    } else {
        inertia = 0.0f; invI = 0.0f;
    }
    const Vec2 oldCenter = sweep.c;
    sweep.localCenter = localCenter;
    sweep.c0 = sweep.c = mul(xf, sweep.localCenter);
    linearVelocity += cross(angularVelocity, sweep.c - oldCenter);
}
inline void Body::synchronizeFixtures() {
    Transform xf1;
    xf1.q.set(sweep.a0);
    xf1.p = sweep.c0 - mul(xf1.q, sweep.localCenter);
    for (auto& f : fixtures) {
        const AABB aabb1 = f->shape->computeAABB(xf1), aabb2 = f->shape->computeAABB(xf);
        AABB aabb;
        aabb.combine(aabb1, aabb2);
        world->moveProxy(f->proxyId, aabb, xf.p - xf1.p);
    }
}
inline bool Body::shouldCollide(const Body* other) const {
    if (type != DYNAMIC_BODY && other->type != DYNAMIC_BODY) return false;   // at least one body should be dynamic
    for (JointEdge* jn = jointList; jn; jn = jn->next) if (jn->other == other && !jn->joint->collideConnected) return false;
    return true;
}

}  // namespace w2d

// TEST INFRASTRUCTURE ONLY.  Nothing under gym.net_b200/ may include, link or load this.
//
// LunarLanderEnv (src/Gym.Environments/Envs/Aether/LunarLanderEnv.cs) restated on top of the GENERIC engine in
// world2d.hpp, call by call: the C# builds an Aether World out of CreateBody / CreateFixture / RevoluteJoint /
// ContactManager delegates and steps it; so does this file.  There is no lander-specific physics here -- no
// hard-coded mass data, no fixed contact slots, no one-body solver -- which is what makes it an independent
// check of gym.net_b200/csrc/lunar_core.cuh (a fixed-topology specialisation written for the GPU) and of its
// CPU twin oracle/lunar.hpp.
//
// The random draws (NumSharp's generator is unreproducible, see oracle/philox.hpp) are INPUTS: reset takes its
// 14 uniforms, step its two dispersion uniforms, so the caller can feed the engine's Philox stream.
//
// export_state / import_state translate between the generic world and the kernel's per-lander state layout
// (gym.net_b200/csrc/lunar.cuh: 80 float + 29 int32 words) -- the only place that knows that layout.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "../detmath.hpp"
#include "world2d.hpp"

using namespace w2d;

namespace {

constexpr float SCALE = 30.0f;                 // LunarLanderEnv.cs:156
constexpr float MAIN_ENGINE_POWER = 13.0f;     // :157
constexpr float SIDE_ENGINE_POWER = 0.6f;      // :158
constexpr int CHUNKS = 11;                     // :503
constexpr int KMAXC = 6;                       // contact slots in the kernel's layout
constexpr int KMAXP = 12;                      // broad-phase pair bytes in the kernel's layout
constexpr int KSTATE = 80, KAUX = 29;
enum : int32_t { F_GAME_OVER = 1, F_LEG0 = 2, F_LEG1 = 4, F_FUSELAGE = 8, F_AWAKE = 16, F_FIRST_STEP = 32, F_CONTINUOUS = 64 };

void sincos_det_hook(float a, float* s, float* c) { oracle::det::sincosf_det(a, s, c); }

struct Component { Body* unit = nullptr; bool contact = false; bool is(Body* b1, Body* b2) const { return unit == b1 || unit == b2; } };

struct LunarSim {
    // constructor arguments (:381)
    bool continuous = false;
    float gravity = -10.0f;
    bool useWind = false;
    float windPower = 15.0f, turbulencePower = 1.5f;
    int windIdx = 0, torqueIdx = 0;   // :409-410, drawn by the caller
    WorldOptions wopt;
    // Body.ApplyForce(Vector2) in the Farseer lineage applies the force at the body ORIGIN (`ApplyForce(ref force, ref _xf.p)`),
    // not at the centre of mass: the initial kick of Reset (:496) then also spins the fuselage.  false = at the centre.
    bool forceAtOrigin = true;

    // fields of the C# class
    int FPS = 50;
    float SIDE_ENGINE_HEIGHT = 14.0f, SIDE_ENGINE_AWAY = 12.0f;
    int VIEWPORT_W = 600, VIEWPORT_H = 400;
    float LEG_W = 2.0f, LEG_H = 8.0f, LEG_AWAY = 20.0f, LEG_DOWN = 18.0f, LEG_SPRING_TORQUE = 40.0f;
    bool gameOver = false;
    std::unique_ptr<World> world;
    Component fuselage, legs[2];
    RevoluteJoint* legJoint[2] = {nullptr, nullptr};
    Body* moon = nullptr;
    float helipadY = 0.0f, prevShaping = 0.0f;
    float smoothY[CHUNKS];
    int toiEvents = 0;

    // ContactDetector (:305-346)
    bool beginContact(Contact* contact) {
        Body* a = contact->fixtureA->body; Body* b = contact->fixtureB->body;
        fuselage.contact = false;
        if (fuselage.is(a, b)) fuselage.contact = true;
        for (int i = 0; i < 2; ++i) {
            legs[i].contact = false;
            if (legs[i].is(a, b)) legs[i].contact = true;
        }
        return false;
    }
    void endContact(Contact* contact) {
        Body* a = contact->fixtureA->body; Body* b = contact->fixtureB->body;
        for (int i = 0; i < 2; ++i) if (legs[i].is(a, b)) legs[i].contact = false;
    }

    void installCallbacks() {
        world->beginContact = [this](Contact* c) { return beginContact(c); };
        world->endContact = [this](Contact* c) { endContact(c); };
    }

    // LunarLanderBody.CreateFuselage (:227-246)
    void createFuselage() {
        Body* b = world->createBody();
        b->setRotation(0.0f);
        b->setType(DYNAMIC_BODY);
        const float poly[6][2] = {{-14.0f, 17.0f}, {-17.0f, 0.0f}, {-17.0f, -10.0f}, {17.0f, -10.0f}, {17.0f, 0.0f}, {14.0f, 17.0f}};   // LANDER_POLY (:189)
        std::vector<Vec2> v;
        for (auto& p : poly) v.push_back(Vec2(p[0] / SCALE, p[1] / SCALE));
        Fixture* f = b->createFixture(std::unique_ptr<Shape>(new PolygonShape(v, 5.0f)));
        f->friction = 0.1f;
        f->category = 1u << 15;      // Category.Cat16
        f->collidesWith = 1u << 0;   // Category.Cat1
        f->restitution = 0.0f;
        b->userIndex = 0;
        fuselage.unit = b;
        fuselage.contact = false;
    }
    // LunarLanderBody.CreateLeg (:255-282)
    void createLeg(int iLeg) {
        Body* b = world->createBody();
        b->setRotation(iLeg == 0 ? -0.05f : 0.05f);
        b->setType(DYNAMIC_BODY);
        b->setPosition(Vec2((iLeg == 0 ? -1.0f : 1.0f) * LEG_AWAY / SCALE, 0.0f));
        std::vector<Vec2> v = {Vec2(0.0f, 0.0f), Vec2(LEG_W / SCALE, 0.0f), Vec2(LEG_W / SCALE, LEG_H / SCALE), Vec2(0.0f, LEG_H / SCALE)};
        Fixture* f = b->createFixture(std::unique_ptr<Shape>(new PolygonShape(v, 1.0f)));
        f->restitution = 0.0f;
        f->category = 1u << 19;      // Category.Cat20
        f->collidesWith = 1u << 0;   // Category.Cat1
        std::unique_ptr<RevoluteJoint> rj(new RevoluteJoint(fuselage.unit, b, Vec2(0.0f, 0.0f), Vec2((iLeg == 0 ? -1.0f : 1.0f) * LEG_AWAY / SCALE, LEG_DOWN / SCALE)));
        rj->enableMotor = true;
        rj->enableLimit = true;
        rj->maxMotorTorque = LEG_SPRING_TORQUE;
        rj->motorSpeed = (iLeg == 0 ? -1.0f : 1.0f) * 0.3f;
        rj->upperAngle = (iLeg == 0 ? 1.0f : -1.0f) * 0.9f + (iLeg == 0 ? 0.0f : 0.5f);
        rj->lowerAngle = (iLeg == 0 ? 1.0f : -1.0f) * 0.9f + (iLeg == 0 ? -0.5f : 0.0f);
        b->userIndex = 1 + iLeg;
        legs[iLeg].unit = b;
        legs[iLeg].contact = false;
        legJoint[iLeg] = world->addJoint(std::move(rj));
    }

    void buildTerrainBody() {
        const float w = (float)VIEWPORT_W / SCALE;
        float chunkX[CHUNKS];
        for (int i = 0; i < CHUNKS; ++i) chunkX[i] = w / (float)(CHUNKS - 1) * (float)i;   // :511-513
        moon = world->createBody();
        moon->setType(STATIC_BODY);
        moon->setPosition(Vec2(0.0f, 0.0f));
        Fixture* base = moon->createFixture(std::unique_ptr<Shape>(new EdgeShape(Vec2(0.0f, 0.0f), Vec2(w, 0.0f))));   // :541
        base->userIndex = 10;
        for (int i = 0; i < CHUNKS - 1; ++i) {                                                                          // :545-557
            const Vec2 p1(chunkX[i], smoothY[i]), p2(chunkX[i + 1], smoothY[i + 1]);
            Fixture* f = moon->createFixture(std::unique_ptr<Shape>(new EdgeShape(p1, p2)));
            f->friction = 0.1f;
            f->collidesWith = 0xffffffffu;   // Category.All
            f->category = 0xffffffffu;
            f->userIndex = i;
        }
        moon->userIndex = 3;
    }

    // Reset (:489-572).  draws: fx, fy (already scaled to +-INITIAL_RANDOM), then 12 heights in [0, h/2).
    void reset(const float draws[14]) {
        World::SinCosScope scope(wopt.sincos);
        world.reset(new World(Vec2(0.0f, gravity), wopt));
        installCallbacks();
        createFuselage();
        createLeg(0);
        createLeg(1);
        {
            const Vec2 f(draws[0], draws[1]);
            Body* fu = fuselage.unit;
            const Vec2 point = forceAtOrigin ? fu->xf.p : fu->sweep.c;
            fu->applyForce(f, point);                                                          // :496
        }
        gameOver = false;
        prevShaping = -FLT_MAX;                                                                // float.MinValue (:498)
        const float h = (float)VIEWPORT_H / SCALE;
        float height[CHUNKS + 1];
        for (int i = 0; i < CHUNKS + 1; ++i) height[i] = draws[2 + i];                         // :505-508
        const int mid = CHUNKS / 2;
        helipadY = h / 4.0f;
        height[mid - 2] = helipadY; height[mid - 1] = helipadY; height[mid] = helipadY; height[mid + 1] = helipadY; height[mid + 2] = helipadY;
        for (int i = 0; i < CHUNKS; ++i) {
            float h1 = 0.0f;
            if (i > 0) h1 = height[i - 1];
            smoothY[i] = 0.33f * (h1 + height[i] + height[i + 1]);
            if (smoothY[i] > h) smoothY[i] = h / 4.0f;
        }
        buildTerrainBody();
        const Vec2 dv((float)VIEWPORT_W / SCALE / 2.0f, (float)VIEWPORT_H / SCALE);           // :560
        fuselage.unit->setPosition(fuselage.unit->position() + dv);
        legs[0].unit->setPosition(legs[0].unit->position() + dv);
        legs[1].unit->setPosition(legs[1].unit->position() + dv);
        // (the zero step of :567-571 is issued by the caller: w2d_lunar_reset)
    }

    struct StepOut { float reward; int done; };

    // Step (:574-774).  disp: the two uniform(-1, 1) draws of :611-612 (before the division by SCALE).
    StepOut step(int iAction, const float cActionIn[2], const float dispIn[2], float obs[8]) {
        World::SinCosScope scope(wopt.sincos);
        Body* fu = fuselage.unit;
        if (useWind && !(legs[0].contact || legs[1].contact)) {                                // :588-596
            const float windMag = (float)(std::tanh(std::sin(0.02 * windIdx) + std::sin(3.14159265358979323846 * 0.01 * windIdx))) * windPower;
            windIdx++;
            fu->applyForce(Vec2(windMag, 0.0f), forceAtOrigin ? fu->xf.p : fu->sweep.c);
            const float torqueMag = (float)(std::tanh(std::sin(0.02 * torqueIdx) + std::sin(3.14159265358979323846 * 0.01 * torqueIdx))) * turbulencePower;
            torqueIdx++;
            fu->applyTorque(torqueMag);
        }
        float a0 = 0.0f, a1 = 0.0f;
        if (continuous) { a0 = clampf(cActionIn[0], -1.0f, 1.0f); a1 = clampf(cActionIn[1], -1.0f, 1.0f); }   // :600
        float sn, cs;
        current_sincos()(fu->rotation(), &sn, &cs);
        const Vec2 tip(sn, cs);                                                                // :609
        const Vec2 side(-tip.y, tip.x);                                                        // :610
        const float dispX = dispIn[0] / SCALE, dispY = dispIn[1] / SCALE;                      // :611-612
        bool fireMain = false, fireThruster = false;
        if (continuous) { if (a0 > 0.0f) fireMain = true; if (std::fabs(a1) > 0.5f) fireThruster = true; }
        else { fireMain = iAction == 2; fireThruster = iAction == 1 || iAction == 3; }
        float mPower = 0.0f;
        if (fireMain) {
            mPower = continuous ? (clampf(a0, 0.0f, 1.0f) + 1.0f) * 0.5f : 1.0f;
            const float ox = tip.x * (4.0f / SCALE + 2.0f * dispX) + side.x * dispY;
            const float oy = -tip.y * (4.0f / SCALE + 2.0f * dispX) - side.y * dispY;
            const Vec2 impulsePos = fu->position() + Vec2(ox, oy);
            // (the exhaust particle of :658-664 has Category.None on both masks: it can touch nothing, so it is not created)
            const Vec2 impulse(-ox * MAIN_ENGINE_POWER * mPower, -oy * MAIN_ENGINE_POWER * mPower);
            fu->applyLinearImpulse(impulse, impulsePos);
        }
        float sPower = 0.0f;
        if (fireThruster) {
            float direction;
            if (continuous) { direction = a1 < 0.0f ? -1.0f : 1.0f; sPower = clampf(std::fabs(a1), 0.5f, 1.0f); }
            else { direction = (float)iAction - 2.0f; sPower = 1.0f; }
            const float ox = tip.x * dispX + side.x * (3.0f * dispY + direction * SIDE_ENGINE_AWAY / SCALE);
            const float oy = -tip.y * dispX - side.y * (3.0f * dispY + direction * SIDE_ENGINE_AWAY / SCALE);
            const Vec2 impulsePos = fu->position() + Vec2(ox - tip.x * 17.0f / SCALE, oy + tip.y * SIDE_ENGINE_HEIGHT / SCALE);
            const Vec2 impulse(-ox * SIDE_ENGINE_POWER * sPower, -oy * SIDE_ENGINE_POWER * sPower);
            fu->applyLinearImpulse(impulse, impulsePos);
        }
        const float dt = 1.0f / (float)FPS;
        world->step(dt, 6 * 30, 2 * 30);                                                       // :721-725
        toiEvents += world->lastToiEvents;
        if (fuselage.contact) gameOver = true;                                                 // :726-729
        observe(obs);
        const float px = obs[0], py = obs[1], vx = obs[2], vy = obs[3];
        float reward = 0.0f;
        float shaping = -100.0f * std::sqrt(px * px + py * py);
        shaping += -100.0f * std::sqrt(vx * vx + vy * vy);
        shaping += -100.0f * std::fabs(obs[4]);
        shaping += 10.0f * (legs[0].contact ? 1.0f : 0.0f);
        shaping += 10.0f * (legs[1].contact ? 1.0f : 0.0f);
        if (prevShaping != -FLT_MAX) reward = shaping - prevShaping;
        prevShaping = shaping;
        reward -= mPower * 0.3f;
        reward -= sPower * 0.03f;
        int done = 0;
        if (gameOver || px > 1.0f) { done = 1; reward = -100.0f; }                             // :762 (one-sided)
        if (!fu->awake) { done = 1; reward = 100.0f; }                                         // :767
        return StepOut{reward, done};
    }

    void observe(float obs[8]) const {                                                         // :733-747
        const Body* fu = fuselage.unit;
        Vec2 pos = fu->position();
        pos.x = (pos.x - (float)VIEWPORT_W / SCALE / 2.0f) / ((float)VIEWPORT_W / SCALE / 2.0f);
        pos.y = (pos.y - (helipadY + LEG_DOWN / SCALE)) / ((float)VIEWPORT_H / SCALE / 2.0f);
        Vec2 vel = fu->linearVelocity;
        vel.x *= ((float)VIEWPORT_W / SCALE / 2.0f) / (float)FPS;
        vel.y *= ((float)VIEWPORT_H / SCALE / 2.0f) / (float)FPS;
        obs[0] = pos.x; obs[1] = pos.y; obs[2] = vel.x; obs[3] = vel.y;
        obs[4] = fu->rotation();
        obs[5] = 20.0f * fu->angularVelocity / (float)FPS;
        obs[6] = legs[0].contact ? 1.0f : 0.0f;
        obs[7] = legs[1].contact ? 1.0f : 0.0f;
    }

    // ---- translation to / from the kernel's per-lander layout ---------------------------------------------
    Body* bodyOf(int k) const { return k == 0 ? fuselage.unit : legs[k - 1].unit; }

    // touching contacts in creation order (the order the kernel keeps its slots in)
    std::vector<Contact*> touchingSorted() const {
        std::vector<Contact*> v;
        for (Contact* c : world->contactList) if (c->touching && (int)v.size() < KMAXC) v.push_back(c);
        return v;
    }

    void exportState(float* s, int32_t* a) const {
        int k = 0;
        for (int i = 0; i < 3; ++i) {
            const Body* b = bodyOf(i);
            s[k++] = b->sweep.c.x; s[k++] = b->sweep.c.y; s[k++] = b->sweep.a;
            s[k++] = b->linearVelocity.x; s[k++] = b->linearVelocity.y; s[k++] = b->angularVelocity; s[k++] = b->sleepTime;
        }
        for (int i = 0; i < 2; ++i) { const RevoluteJoint* j = legJoint[i]; s[k++] = j->impulse.x; s[k++] = j->impulse.y; s[k++] = j->impulse.z; s[k++] = j->motorImpulse; }
        const std::vector<Contact*> tc = touchingSorted();
        for (int slot = 0; slot < KMAXC; ++slot) {
            for (int p = 0; p < 2; ++p) {
                const bool live = slot < (int)tc.size() && p < tc[slot]->manifold.pointCount;
                s[k++] = live ? tc[slot]->manifold.points[p].normalImpulse : 0.0f;
                s[k++] = live ? tc[slot]->manifold.points[p].tangentImpulse : 0.0f;
            }
        }
        for (int i = 0; i < CHUNKS; ++i) s[k++] = smoothY[i];
        s[k++] = prevShaping;
        s[k++] = fuselage.unit->force.x; s[k++] = fuselage.unit->force.y; s[k++] = fuselage.unit->torque;
        for (int i = 0; i < 3; ++i) {
            const AABB& f = world->proxies[bodyOf(i)->fixtures[0]->proxyId].fat;
            s[k++] = f.lo.x; s[k++] = f.lo.y; s[k++] = f.hi.x; s[k++] = f.hi.y;
        }
        k = 0;
        uint32_t touch[3] = {0u, 0u, 0u};
        for (Contact* c : tc) touch[c->fixtureB->body->userIndex] |= 1u << c->fixtureA->userIndex;
        for (int i = 0; i < 3; ++i) a[k++] = (int32_t)touch[i];
        int32_t flags = 0;
        if (gameOver) flags |= F_GAME_OVER;
        if (legs[0].contact) flags |= F_LEG0;
        if (legs[1].contact) flags |= F_LEG1;
        if (fuselage.contact) flags |= F_FUSELAGE;
        if (fuselage.unit->awake) flags |= F_AWAKE;
        if (world->inv_dt0 == 0.0f) flags |= F_FIRST_STEP;
        if (continuous) flags |= F_CONTINUOUS;
        a[k++] = flags;
        a[k++] = (int32_t)legJoint[0]->limitState; a[k++] = (int32_t)legJoint[1]->limitState;
        for (int slot = 0; slot < KMAXC; ++slot) {
            if (slot < (int)tc.size()) {
                const Contact* c = tc[slot];
                a[k++] = c->fixtureB->body->userIndex * 16 + c->fixtureA->userIndex;
                for (int p = 0; p < 2; ++p) a[k++] = p < c->manifold.pointCount ? (int32_t)c->manifold.points[p].id.key() : -1;
            } else {
                a[k++] = -1; a[k++] = -1; a[k++] = -1;
            }
        }
        a[k++] = windIdx; a[k++] = torqueIdx;
        uint32_t pw[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu};
        int np = 0;
        for (Contact* c : world->contactList) {
            if (np >= KMAXP) break;
            const uint32_t pr = (uint32_t)(c->fixtureB->body->userIndex * 16 + c->fixtureA->userIndex);
            pw[np >> 2] = (pw[np >> 2] & ~(0xffu << (8 * (np & 3)))) | (pr << (8 * (np & 3)));
            ++np;
        }
        for (int i = 0; i < 3; ++i) a[k++] = (int32_t)pw[i];
    }

    // Builds a world in the given state: bodies, joints and terrain as Reset makes them, then poses, velocities,
    // accumulated impulses and the touching contacts (ids + impulses) of the previous step.
    void importState(const float* s, const int32_t* a) {
        World::SinCosScope scope(wopt.sincos);
        world.reset(new World(Vec2(0.0f, gravity), wopt));
        installCallbacks();
        createFuselage();
        createLeg(0);
        createLeg(1);
        const float h = (float)VIEWPORT_H / SCALE;
        helipadY = h / 4.0f;
        for (int i = 0; i < CHUNKS; ++i) smoothY[i] = s[53 + i];
        buildTerrainBody();
        int k = 0;
        for (int i = 0; i < 3; ++i) {
            Body* b = bodyOf(i);
            const Vec2 c(s[k], s[k + 1]); const float ang = s[k + 2];
            k += 3;
            // place the body so that its centre of mass is c: origin = c - R(ang) * localCenter
            const Rot q(ang);
            b->setTransform(c - mul(q, b->sweep.localCenter), ang);
            b->sweep.c = c; b->sweep.c0 = c;   // exactly the stored centre (setTransform recomputes it from the origin)
            b->linearVelocity = Vec2(s[k], s[k + 1]); b->angularVelocity = s[k + 2]; b->sleepTime = s[k + 3];
            k += 4;
        }
        for (int i = 0; i < 2; ++i) { RevoluteJoint* j = legJoint[i]; j->impulse = Vec3(s[k], s[k + 1], s[k + 2]); j->motorImpulse = s[k + 3]; k += 4; }
        const float* slotImp = s + k;
        k += 4 * KMAXC + CHUNKS;
        prevShaping = s[k++];
        fuselage.unit->force = Vec2(s[k], s[k + 1]); fuselage.unit->torque = s[k + 2];
        const int32_t flags = a[3];
        gameOver = (flags & F_GAME_OVER) != 0;
        legs[0].contact = (flags & F_LEG0) != 0;
        legs[1].contact = (flags & F_LEG1) != 0;
        fuselage.contact = (flags & F_FUSELAGE) != 0;
        const bool awake = (flags & F_AWAKE) != 0;
        for (int i = 0; i < 3; ++i) bodyOf(i)->awake = awake;
        world->inv_dt0 = (flags & F_FIRST_STEP) ? 0.0f : 1.0f / (1.0f / (float)FPS);
        continuous = (flags & F_CONTINUOUS) != 0;
        legJoint[0]->limitState = (LimitState)a[4]; legJoint[1]->limitState = (LimitState)a[5];
        windIdx = a[6 + 3 * KMAXC]; torqueIdx = a[7 + 3 * KMAXC];
        // broad phase: the stored proxy boxes, and the contacts that exist, re-created in their creation order
        while (!world->contactList.empty()) world->destroyContact(world->contactList.back());   // pairs met while the bodies were being placed
        world->moveBuffer.clear();
        world->newFixture = false;
        for (int i = 0; i < 3; ++i) {
            AABB& f = world->proxies[bodyOf(i)->fixtures[0]->proxyId].fat;
            const float* q = s + 68 + 4 * i;
            f.lo = Vec2(q[0], q[1]); f.hi = Vec2(q[2], q[3]);
        }
        for (int n = 0; n < KMAXP; ++n) {
            const uint32_t pr = ((uint32_t)a[26 + (n >> 2)] >> (8 * (n & 3))) & 0xffu;
            if (pr == 0xffu) break;
            const int body = (int)(pr >> 4), edge = (int)(pr & 15u);
            Fixture* fe = nullptr;
            for (auto& f : moon->fixtures) if (f->userIndex == edge) fe = f.get();
            world->addPair(bodyOf(body)->fixtures[0].get(), fe);
        }
        for (int i = 0; i < 3; ++i) bodyOf(i)->awake = awake;   // (creating a contact wakes its bodies)
        for (int slot = 0; slot < KMAXC; ++slot) {
            const int32_t pair = a[6 + 3 * slot];
            if (pair < 0) continue;
            const int body = pair / 16, edge = pair % 16;
            Contact* found = nullptr;
            for (Contact* c : world->contactList) if (c->fixtureB->body->userIndex == body && c->fixtureA->userIndex == edge) found = c;
            if (!found) {
                Fixture* fe = nullptr;
                for (auto& f : moon->fixtures) if (f->userIndex == edge) fe = f.get();
                world->addPair(bodyOf(body)->fixtures[0].get(), fe);
                for (Contact* c : world->contactList) if (c->fixtureB->body->userIndex == body && c->fixtureA->userIndex == edge) found = c;
            }
            found->touching = true;
            int n = 0;
            for (int p = 0; p < 2; ++p) {
                const uint32_t key = (uint32_t)a[6 + 3 * slot + 1 + p];
                if (key == 0xffffffffu) continue;
                ManifoldPoint& mp = found->manifold.points[n++];
                mp.id.indexA = (uint8_t)(key & 0xff); mp.id.indexB = (uint8_t)((key >> 8) & 0xff);
                mp.id.typeA = (uint8_t)((key >> 16) & 0xff); mp.id.typeB = (uint8_t)((key >> 24) & 0xff);
                mp.normalImpulse = slotImp[4 * slot + 2 * p]; mp.tangentImpulse = slotImp[4 * slot + 2 * p + 1];
            }
            found->manifold.pointCount = n;
        }
    }
};

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// C API (ctypes: tests/world2d_lib.py)
// ------------------------------------------------------------------------------------------------------------
extern "C" {

struct w2d_lunar_options {
    int32_t continuous;
    float gravity;
    int32_t use_wind;
    float wind_power, turbulence_power;
    int32_t wind_idx, torque_idx;
    int32_t begin_contact_false;     // WorldOptions::beginContactFalse
    int32_t continuous_physics;      // TOI pass
    int32_t reverse_seed_order;
    int32_t contact_list_head_insertion;
    int32_t det_sincos;              // 1: the engine's deterministic float32 sincos (oracle/detmath.hpp); 0: (float)sin((double)a)
    int32_t force_at_origin;         // Body.ApplyForce(Vector2) applies at the body origin (Farseer lineage)
    int32_t canonical_contact_order; // diagnostic: WorldOptions::canonicalContactOrder
};

void w2d_lunar_default_options(w2d_lunar_options* o) {
    o->continuous = 0; o->gravity = -10.0f; o->use_wind = 0; o->wind_power = 15.0f; o->turbulence_power = 1.5f;
    o->wind_idx = 0; o->torque_idx = 0;
    o->begin_contact_false = 0; o->continuous_physics = 0; o->reverse_seed_order = 0; o->contact_list_head_insertion = 0;
    o->det_sincos = 0; o->force_at_origin = 1; o->canonical_contact_order = 0;
}

void* w2d_lunar_create(const w2d_lunar_options* o) {
    LunarSim* s = new LunarSim();
    s->continuous = o->continuous != 0; s->gravity = o->gravity; s->useWind = o->use_wind != 0;
    s->windPower = o->wind_power; s->turbulencePower = o->turbulence_power; s->windIdx = o->wind_idx; s->torqueIdx = o->torque_idx;
    s->wopt.beginContactFalse = o->begin_contact_false;
    s->wopt.continuousPhysics = o->continuous_physics != 0;
    s->wopt.reverseSeedOrder = o->reverse_seed_order != 0;
    s->wopt.contactListHeadInsertion = o->contact_list_head_insertion != 0;
    s->wopt.sincos = o->det_sincos ? sincos_det_hook : sincos_libm_double;
    s->forceAtOrigin = o->force_at_origin != 0;
    s->wopt.canonicalContactOrder = o->canonical_contact_order != 0;
    return s;
}
void w2d_lunar_destroy(void* h) { delete static_cast<LunarSim*>(h); }

/* Reset (:489-572) including its zero step (:567-571).  reset_draws[14] = fx, fy, 12 heights (already mapped to their
 * ranges); step_draws[2] = the zero step's dispersion uniforms in [-1, 1). */
void w2d_lunar_reset(void* h, const float* reset_draws, const float* step_draws, float* obs) {
    LunarSim* s = static_cast<LunarSim*>(h);
    s->reset(reset_draws);
    const float zero[2] = {0.0f, 0.0f};
    s->step(0, zero, step_draws, obs);
}
void w2d_lunar_step(void* h, int32_t i_action, const float* c_action, const float* step_draws, float* obs, float* reward, int32_t* done) {
    LunarSim* s = static_cast<LunarSim*>(h);
    const float zero[2] = {0.0f, 0.0f};
    const LunarSim::StepOut r = s->step(i_action, c_action ? c_action : zero, step_draws, obs);
    *reward = r.reward; *done = r.done;
}
void w2d_lunar_export(void* h, float* state68, int32_t* aux26) { static_cast<LunarSim*>(h)->exportState(state68, aux26); }
void w2d_lunar_import(void* h, const float* state68, const int32_t* aux26) { static_cast<LunarSim*>(h)->importState(state68, aux26); }
/* analysis: record the iteration at which the velocity / position sweeps first repeat (out[6]: vel at, vel period, pos at, pos period, pos iterations run, contacts touching) */
void w2d_lunar_record_cycles(void* h, int32_t on) { static_cast<LunarSim*>(h)->world->recordCycles = on != 0; }
void w2d_lunar_cycle_info(void* h, int32_t* out) {
    World* w = static_cast<LunarSim*>(h)->world.get();
    out[0] = w->velCycleAt; out[1] = w->velCyclePeriod; out[2] = w->posCycleAt; out[3] = w->posCyclePeriod; out[4] = w->posIterationsRun;
    int t = 0; for (Contact* c : w->contactList) if (c->touching) ++t;
    out[5] = t;
}
int32_t w2d_lunar_toi_events(void* h) { return static_cast<LunarSim*>(h)->toiEvents; }
int32_t w2d_lunar_num_contacts(void* h) { return (int32_t)static_cast<LunarSim*>(h)->world->contactList.size(); }

/* mass data the generic engine derives from the vertices: [body] -> mass, inv_mass, inertia (about the centre), inv_inertia,
 * local centre x, y, shape centroid x, y */
void w2d_lunar_mass_data(void* h, float* out24) {
    LunarSim* s = static_cast<LunarSim*>(h);
    for (int i = 0; i < 3; ++i) {
        const Body* b = s->bodyOf(i);
        const PolygonShape* p = static_cast<const PolygonShape*>(b->fixtures[0]->shape.get());
        float* o = out24 + 8 * i;
        o[0] = b->mass; o[1] = b->invMass; o[2] = b->inertia; o[3] = b->invI;
        o[4] = b->sweep.localCenter.x; o[5] = b->sweep.localCenter.y; o[6] = p->md.center.x; o[7] = p->md.center.y;
    }
}
/* polygon of body i as the hull builder ordered it: returns the vertex count; verts/normals [8][2] */
int32_t w2d_lunar_polygon(void* h, int32_t body, float* verts, float* normals) {
    LunarSim* s = static_cast<LunarSim*>(h);
    const PolygonShape* p = static_cast<const PolygonShape*>(s->bodyOf(body)->fixtures[0]->shape.get());
    for (size_t i = 0; i < p->vertices.size(); ++i) { verts[2 * i] = p->vertices[i].x; verts[2 * i + 1] = p->vertices[i].y; normals[2 * i] = p->normals[i].x; normals[2 * i + 1] = p->normals[i].y; }
    return (int32_t)p->vertices.size();
}

}  // extern "C"

// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Container + C API of the CPU oracle.  Instances are stepped in a serial loop, the structure of
// the reference's VecEnvWrapper.Step (src/Gym/Envs/VecEnvWrapper.cs:22-24); with threads > 1 the
// loop is split over host threads the way the reference's DistributedScheduler pool
// (src/Gym/Internal/Threading/DistributedScheduler.cs:12-13) would spread independent envs.
//
// Vector semantics that the reference does not have (auto-reset, time limit, the Philox stream)
// are the engine's own and are defined in DESIGN.md; this file is their executable restatement.
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "classic.hpp"
#include "philox.hpp"
#ifdef ORACLE_WITH_LUNAR
#include "lunar.hpp"
#endif

using namespace oracle;

struct KindInfo { int sd, ad, od, actd, actn, default_limit; };
static const KindInfo KINDS[] = {
    /* CARTPOLE         */ {4, 3, 4, 1, 2, 0},
    /* PENDULUM         */ {2, 3, 3, 1, 0, 200},
    /* MOUNTAINCAR      */ {2, 3, 2, 1, 3, 200},
    /* MOUNTAINCAR_CONT */ {2, 3, 2, 1, 0, 999},
    /* ACROBOT          */ {4, 3, 6, 1, 3, 500},
#ifdef ORACLE_WITH_LUNAR
    /* LUNARLANDER      */ {lunar::STATE_DIM, lunar::AUX_DIM, 8, 1, 4, 0},
    /* LUNARLANDER_CONT */ {lunar::STATE_DIM, lunar::AUX_DIM, 8, 2, 0, 0},
#endif
};
static const int NUM_KINDS = (int)(sizeof(KINDS) / sizeof(KINDS[0]));

struct oracle_env {
    int kind, n, mode, limit, threads;
    uint64_t seed, t;
    uint32_t off, flags;
    KindInfo ki;
    std::vector<double> sd;      // F64 modes
    std::vector<float> sf;       // F32 mode
    std::vector<int32_t> aux;    // [n][ad]: classic = {steps_beyond_done, episode_step, episode ordinal}
    std::vector<int32_t> seeds;  // optional per-env seeds (VecEnv.Seed(int[]), src/Gym/Envs/VecEnv.cs:48-53)
#ifdef ORACLE_WITH_LUNAR
    std::vector<lunar::Lander> landers;
    float gravity = -10.0f, wind_power = 15.0f, turbulence_power = 1.5f;
    int use_wind = 0;
#endif
    int EPT() const { return is_lunar() ? ki.ad - 2 : 1; }
    int ORD() const { return is_lunar() ? ki.ad - 1 : 2; }
    bool is_lunar() const { return kind >= ORACLE_LUNARLANDER; }
    uint64_t seed_of(int i) const { return seeds.empty() ? seed : (uint64_t)(uint32_t)seeds[i]; }
};

static void write_obs(oracle_env* e, int i, float* obs) {
    if (!obs) return;
    float* o = obs + (size_t)i * e->ki.od;
#ifdef ORACLE_WITH_LUNAR
    if (e->is_lunar()) { lunar::observe(e->landers[i], o); return; }
#endif
    const bool f32 = e->mode == ORACLE_MODE_F32;
    const double* d = f32 ? nullptr : &e->sd[(size_t)i * e->ki.sd];
    const float* f = f32 ? &e->sf[(size_t)i * e->ki.sd] : nullptr;
    switch (e->kind) {
        case ORACLE_CARTPOLE:
        case ORACLE_MOUNTAINCAR:
        case ORACLE_MOUNTAINCAR_CONT:
            for (int k = 0; k < e->ki.sd; ++k) o[k] = f32 ? f[k] : (float)d[k];
            break;
        case ORACLE_PENDULUM:
            if (f32) { float s, c; det::sincosf_det(f[0], &s, &c); o[0] = c; o[1] = s; o[2] = f[1]; }
            else { o[0] = (float)std::cos(d[0]); o[1] = (float)std::sin(d[0]); o[2] = (float)d[1]; }
            break;
        case ORACLE_ACROBOT:
            if (f32) {
                float s1, c1, s2, c2;
                det::sincosf_det(f[0], &s1, &c1); det::sincosf_det(f[1], &s2, &c2);
                o[0] = c1; o[1] = s1; o[2] = c2; o[3] = s2; o[4] = f[2]; o[5] = f[3];
            } else {
                o[0] = (float)std::cos(d[0]); o[1] = (float)std::sin(d[0]);
                o[2] = (float)std::cos(d[1]); o[3] = (float)std::sin(d[1]);
                o[4] = (float)d[2]; o[5] = (float)d[3];
            }
            break;
    }
}

// Reset of instance i: the RESET draw is indexed by the env's episode ordinal (RNG spec v1).
static void reset_one(oracle_env* e, int i, uint64_t next_t) {
    int32_t* auxp = &e->aux[(size_t)i * e->ki.ad];
    const uint64_t index = (uint64_t)(uint32_t)auxp[e->ORD()];
    auxp[e->ORD()] += 1;
    auxp[e->EPT()] = 0;
    const uint32_t gid = e->off + (uint32_t)i;
    const uint64_t seed = e->seed_of(i);
#ifdef ORACLE_WITH_LUNAR
    if (e->is_lunar()) {
        lunar::reset(e->landers[i], seed, gid, index, e->kind == ORACLE_LUNARLANDER_CONT, next_t,
                     e->gravity, e->use_wind, e->wind_power, e->turbulence_power);
        return;
    }
#endif
    (void)next_t;
    Block b = draw(seed, gid, index, STREAM_RESET);
    float v[4] = {0, 0, 0, 0};
    const float PI_F = 3.1415927410125732f;
    switch (e->kind) {
        case ORACLE_CARTPOLE:   // CartPoleEnv.cs:65  uniform(-0.05, 0.05, 4)
            for (int k = 0; k < 4; ++k) v[k] = uniformf(-0.05f, 0.05f, b.w[k]);
            break;
        case ORACLE_PENDULUM:   // upstream: uniform(-[pi, 1], [pi, 1])
            v[0] = uniformf(-PI_F, PI_F, b.w[0]); v[1] = uniformf(-1.0f, 1.0f, b.w[1]);
            break;
        case ORACLE_MOUNTAINCAR:
        case ORACLE_MOUNTAINCAR_CONT:   // upstream: [uniform(-0.6, -0.4), 0]
            v[0] = uniformf(-0.6f, -0.4f, b.w[0]); v[1] = 0.0f;
            break;
        case ORACLE_ACROBOT:    // upstream: uniform(-0.1, 0.1, 4)
            for (int k = 0; k < 4; ++k) v[k] = uniformf(-0.1f, 0.1f, b.w[k]);
            break;
    }
    for (int k = 0; k < e->ki.sd; ++k) {
        if (e->mode == ORACLE_MODE_F32) e->sf[(size_t)i * e->ki.sd + k] = v[k];
        else e->sd[(size_t)i * e->ki.sd + k] = (double)v[k];
    }
    auxp[0] = -1;   // steps_beyond_done = -1 (CartPoleEnv.cs:64)
}

static bool action_valid(const oracle_env* e, int a) { return a >= 0 && a < e->ki.actn; }

// One instance, one step.  Returns 1 if the action was invalid (instance left untouched).
static int step_one_t(oracle_env* e, int i, const void* actions, float* obs, float* reward, uint8_t* done, uint64_t now) {
    const KindInfo& ki = e->ki;
    int ia = 0; const float* fa = nullptr;
    if (ki.actn > 0) ia = ((const int32_t*)actions)[i];
    else fa = (const float*)actions + (size_t)i * ki.actd;
    StepOut r{0.0f, 0};
    int invalid = 0;
    int32_t* aux = &e->aux[(size_t)i * ki.ad];
#ifdef ORACLE_WITH_LUNAR
    if (e->is_lunar()) {
        if (ki.actn > 0 && !action_valid(e, ia)) invalid = 1;   // LunarLanderEnv.cs:604-607 throws InvalidActionError
        else {
            lunar::StepResult lr = lunar::step(e->landers[i], e->seed_of(i), e->off + (uint32_t)i, now, ia, fa);
            r.reward = lr.reward; r.done = lr.done;
            if (e->limit > 0) aux[e->EPT()] += 1;
        }
    } else
#endif
    {
        // CartPole accepts anything in Release builds (Debug.Assert only, CartPoleEnv.cs:139) and
        // treats != 1 as "left"; the upstream-spec envs assert -> instance not stepped.
        if (ki.actn > 0 && e->kind != ORACLE_CARTPOLE && !action_valid(e, ia)) invalid = 1;
        else if (ki.actn == 0 && !(fa[0] == fa[0])) invalid = 1;   // NaN torque/force
        else {
            const bool f32 = e->mode == ORACLE_MODE_F32;
            double* d = f32 ? nullptr : &e->sd[(size_t)i * ki.sd];
            float* f = f32 ? &e->sf[(size_t)i * ki.sd] : nullptr;
            switch (e->kind) {
                case ORACLE_CARTPOLE:
                    r = f32 ? cartpole_step_f32(f, ia, &aux[0]) : cartpole_step_f64(d, ia, &aux[0]); break;
                case ORACLE_PENDULUM:
                    r = f32 ? pendulum_step_f32(f, fa[0]) : pendulum_step_f64(d, fa[0]); break;
                case ORACLE_MOUNTAINCAR:
                    r = f32 ? mountaincar_any_step_f32(f, false, ia, 0.0f) : mountaincar_step_f64(d, ia); break;
                case ORACLE_MOUNTAINCAR_CONT:
                    r = f32 ? mountaincar_any_step_f32(f, true, 0, fa[0]) : mountaincar_cont_step_f64(d, fa[0]); break;
                case ORACLE_ACROBOT:
                    r = f32 ? acrobot_step_f32(f, ia) : acrobot_step_f64(d, ia); break;
            }
            if (e->mode == ORACLE_MODE_F64_F32STORE)
                for (int k = 0; k < ki.sd; ++k) d[k] = (double)(float)d[k];
            if (e->limit > 0) aux[1] += 1;   // the episode-step counter exists only under a time limit
        }
    }
    if (!invalid && e->limit > 0 && aux[e->EPT()] >= e->limit && !r.done)   // truncation folded into done
        r.done = (e->flags & ORACLE_FLAG_DONE_BITS) ? 2 : 1;
    if (!invalid && r.done && (e->flags & ORACLE_FLAG_AUTO_RESET)) reset_one(e, i, now + 1);
    write_obs(e, i, obs);
    if (reward) reward[i] = r.reward;
    if (done) done[i] = r.done;
    return invalid;
}

static int step_one(oracle_env* e, int i, const void* actions, float* obs, float* reward, uint8_t* done) {
    return step_one_t(e, i, actions, obs, reward, done, e->t);
}
static int step_one_at(oracle_env* e, int i, const void* actions, uint64_t now) {
    return step_one_t(e, i, actions, nullptr, nullptr, nullptr, now);
}

template <class F>
static void parallel_for(int n, int threads, F f) {
    if (threads <= 1 || n < 2 * threads) { f(0, n); return; }
    std::vector<std::thread> pool;
    int chunk = (n + threads - 1) / threads;
    for (int k = 0; k < threads; ++k) {
        int lo = k * chunk, hi = std::min(n, lo + chunk);
        if (lo >= hi) break;
        pool.emplace_back([=] { f(lo, hi); });
    }
    for (auto& th : pool) th.join();
}

#ifdef ORACLE_WITH_LUNAR
// LunarLanderEnv ctor (:409-410): _wind_idx / _torque_idx = randint(-9999, 9999), drawn once per generator
static void lunar_ctor_draws(oracle_env* e) {
    for (int i = 0; i < e->n; ++i) {
        Block b = draw(e->seed_of(i), e->off + (uint32_t)i, 0, STREAM_CTOR);
        e->landers[i].wind_idx = -9999 + (int32_t)(((uint64_t)b.w[0] * 19998u) >> 32);
        e->landers[i].torque_idx = -9999 + (int32_t)(((uint64_t)b.w[1] * 19998u) >> 32);
    }
}
#endif

extern "C" {

void oracle_set_lunar_params(oracle_env* e, float gravity, int use_wind, float wind_power, float turbulence_power) {
#ifdef ORACLE_WITH_LUNAR
    e->gravity = gravity; e->use_wind = use_wind; e->wind_power = wind_power; e->turbulence_power = turbulence_power;
#else
    (void)e; (void)gravity; (void)use_wind; (void)wind_power; (void)turbulence_power;
#endif
}

int oracle_dims(int kind, int* sd, int* ad, int* od, int* actd, int* actn) {
    if (kind < 0 || kind >= NUM_KINDS) return -1;
    if (sd) *sd = KINDS[kind].sd;
    if (ad) *ad = KINDS[kind].ad;
    if (od) *od = KINDS[kind].od;
    if (actd) *actd = KINDS[kind].actd;
    if (actn) *actn = KINDS[kind].actn;
    return 0;
}

oracle_env* oracle_create(int kind, int n, uint64_t seed, uint32_t off, uint32_t flags, int time_limit, int mode) {
    if (kind < 0 || kind >= NUM_KINDS || n <= 0 || mode < 0 || mode > 2) return nullptr;
    oracle_env* e = new oracle_env();
    e->kind = kind; e->n = n; e->mode = mode; e->seed = seed; e->off = off; e->flags = flags;
    e->t = 0; e->threads = 1; e->ki = KINDS[kind];
    e->limit = time_limit == 0 ? e->ki.default_limit : (time_limit < 0 ? 0 : time_limit);
    if (mode == ORACLE_MODE_F32) e->sf.assign((size_t)n * e->ki.sd, 0.0f);
    else e->sd.assign((size_t)n * e->ki.sd, 0.0);
    e->aux.assign((size_t)n * e->ki.ad, 0);
    if (!e->is_lunar()) for (int i = 0; i < n; ++i) e->aux[(size_t)i * e->ki.ad] = -1;
#ifdef ORACLE_WITH_LUNAR
    if (e->is_lunar()) {
        e->landers.resize(n);
        for (int i = 0; i < n; ++i) std::memset(&e->landers[i], 0, sizeof(lunar::Lander));
        lunar_ctor_draws(e);
    }
#endif
    return e;
}

void oracle_destroy(oracle_env* e) { delete e; }
static void restart_streams(oracle_env* e) {   // a new generator restarts every stream (CartPoleEnv.cs:197)
    for (int i = 0; i < e->n; ++i) e->aux[(size_t)i * e->ki.ad + e->ORD()] = 0;
    e->t = 0;
#ifdef ORACLE_WITH_LUNAR
    if (e->is_lunar()) lunar_ctor_draws(e);
#endif
}
void oracle_seed(oracle_env* e, uint64_t seed) { e->seed = seed; e->seeds.clear(); restart_streams(e); }
void oracle_seed_each(oracle_env* e, const int32_t* seeds) { e->seeds.assign(seeds, seeds + e->n); restart_streams(e); }
void oracle_set_threads(oracle_env* e, int threads) { e->threads = threads < 1 ? 1 : threads; }

void oracle_reset(oracle_env* e, float* obs) {
    parallel_for(e->n, e->threads, [=](int lo, int hi) {
        for (int i = lo; i < hi; ++i) { reset_one(e, i, e->t); write_obs(e, i, obs); }
    });
}

void oracle_reset_masked(oracle_env* e, const uint8_t* mask, float* obs) {
    for (int i = 0; i < e->n; ++i) {
        if (mask[i]) reset_one(e, i, e->t);
        write_obs(e, i, obs);
    }
}

int oracle_step(oracle_env* e, const void* actions, float* obs, float* reward, uint8_t* done) {
    std::atomic<int> bad{0};
    parallel_for(e->n, e->threads, [&](int lo, int hi) {
        int c = 0;
        for (int i = lo; i < hi; ++i) c += step_one(e, i, actions, obs, reward, done);
        bad += c;
    });
    e->t += 1;
    return bad.load();
}

// K steps with pre-generated actions [K][n][act_dim], no outputs: each host thread runs its slice of
// independent envs through all K steps (one thread launch per call) -- the CPU-baseline timing loop.
int oracle_step_many(oracle_env* e, int K, const void* actions) {
    std::atomic<int> bad{0};
    const size_t stride = (size_t)e->n * (e->ki.actn > 0 ? 1 : e->ki.actd) * 4;
    const uint64_t t0 = e->t;
    parallel_for(e->n, e->threads, [&](int lo, int hi) {
        int c = 0;
        for (int k = 0; k < K; ++k) {
            const char* act = (const char*)actions + (size_t)k * stride;
            for (int i = lo; i < hi; ++i) c += step_one_at(e, i, act, t0 + (uint64_t)k);
        }
        bad += c;
    });
    e->t = t0 + (uint64_t)K;
    return bad.load();
}

void oracle_rollout_random(oracle_env* e, int K, float* obs, float* reward, uint8_t* done, void* actions) {
    const KindInfo ki = e->ki;
    const int n = e->n;
    std::vector<int32_t> ia; std::vector<float> fa;
    if (ki.actn > 0) ia.resize(n); else fa.resize((size_t)n * ki.actd);
    for (int k = 0; k < K; ++k) {
        const uint64_t t = e->t;
        for (int i = 0; i < n; ++i) {
            const uint32_t gid = e->off + (uint32_t)i;
            const uint64_t seed = e->seed_of(i);
            switch (e->kind) {
                case ORACLE_CARTPOLE: ia[i] = action_discrete2(seed, gid, t); break;
                case ORACLE_MOUNTAINCAR:
                case ORACLE_ACROBOT: ia[i] = action_discrete3(seed, gid, t); break;
                case ORACLE_LUNARLANDER: ia[i] = action_discrete4(seed, gid, t); break;
                case ORACLE_PENDULUM: fa[i] = action_box1(seed, gid, t, -2.0f, 2.0f); break;
                case ORACLE_MOUNTAINCAR_CONT: fa[i] = action_box1(seed, gid, t, -1.0f, 1.0f); break;
                case ORACLE_LUNARLANDER_CONT: action_box2(seed, gid, t, -1.0f, 1.0f, &fa[(size_t)i * 2]); break;
            }
        }
        const void* act = ki.actn > 0 ? (const void*)ia.data() : (const void*)fa.data();
        if (actions) {
            if (ki.actn > 0) std::memcpy((int32_t*)actions + (size_t)k * n, ia.data(), sizeof(int32_t) * n);
            else std::memcpy((float*)actions + (size_t)k * n * ki.actd, fa.data(), sizeof(float) * n * ki.actd);
        }
        oracle_step(e, act, obs ? obs + (size_t)k * n * ki.od : nullptr,
                    reward ? reward + (size_t)k * n : nullptr, done ? done + (size_t)k * n : nullptr);
    }
}

// ActionSpace.Sample() at the current step index (Discrete.cs:17-28 with a mask, Box.cs:84): the ACTION-stream
// draw of step t; with a mask, word 0 of sub-block 1 picks uniformly among the entries equal to 1.
void oracle_sample_actions(oracle_env* e, const uint8_t* mask, void* out) {
    const KindInfo ki = e->ki;
    for (int i = 0; i < e->n; ++i) {
        const uint32_t gid = e->off + (uint32_t)i;
        const uint64_t seed = e->seed_of(i), t = e->t;
        if (ki.actn > 0) {
            int a = 0;
            switch (ki.actn) {
                case 2: a = action_discrete2(seed, gid, t); break;
                case 3: a = action_discrete3(seed, gid, t); break;
                default: a = action_discrete4(seed, gid, t); break;
            }
            if (mask) {
                const uint8_t* m = mask + (size_t)i * ki.actn;
                int valid = 0;
                for (int k = 0; k < ki.actn; ++k) valid += m[k] == 1;
                a = 0;
                if (valid > 0) {
                    Block b = draw(seed, gid, t, STREAM_ACTION, 1);
                    int j = (int)(((uint64_t)b.w[0] * (uint32_t)valid) >> 32);
                    for (int k = 0; k < ki.actn; ++k) if (m[k] == 1) { if (j == 0) { a = k; break; } --j; }
                }
            }
            ((int32_t*)out)[i] = a;
        } else if (ki.actd == 1) {
            const float lo = e->kind == ORACLE_PENDULUM ? -2.0f : -1.0f;
            ((float*)out)[i] = action_box1(seed, gid, t, lo, -lo);
        } else {
            action_box2(seed, gid, t, -1.0f, 1.0f, (float*)out + (size_t)i * 2);
        }
    }
}

void oracle_get_state(oracle_env* e, double* state, int32_t* aux, uint64_t* t) {
    const size_t m = (size_t)e->n * e->ki.sd;
#ifdef ORACLE_WITH_LUNAR
    if (e->is_lunar()) {
        for (int i = 0; i < e->n; ++i) {
            lunar::get_state(e->landers[i], state ? state + (size_t)i * e->ki.sd : nullptr, aux ? aux + (size_t)i * e->ki.ad : nullptr);
            if (aux) {
                aux[(size_t)i * e->ki.ad + e->EPT()] = e->aux[(size_t)i * e->ki.ad + e->EPT()];
                aux[(size_t)i * e->ki.ad + e->ORD()] = e->aux[(size_t)i * e->ki.ad + e->ORD()];
            }
        }
        if (t) *t = e->t;
        return;
    }
#endif
    if (state) for (size_t k = 0; k < m; ++k) state[k] = e->mode == ORACLE_MODE_F32 ? (double)e->sf[k] : e->sd[k];
    if (aux) std::memcpy(aux, e->aux.data(), sizeof(int32_t) * e->aux.size());
    if (t) *t = e->t;
}

void oracle_set_state(oracle_env* e, const double* state, const int32_t* aux, uint64_t t) {
    const size_t m = (size_t)e->n * e->ki.sd;
#ifdef ORACLE_WITH_LUNAR
    if (e->is_lunar()) {
        for (int i = 0; i < e->n; ++i) {
            lunar::set_state(e->landers[i], state + (size_t)i * e->ki.sd, aux + (size_t)i * e->ki.ad);
            e->aux[(size_t)i * e->ki.ad + e->EPT()] = aux[(size_t)i * e->ki.ad + e->EPT()];
            e->aux[(size_t)i * e->ki.ad + e->ORD()] = aux[(size_t)i * e->ki.ad + e->ORD()];
        }
        e->t = t;
        return;
    }
#endif
    if (state) for (size_t k = 0; k < m; ++k) { if (e->mode == ORACLE_MODE_F32) e->sf[k] = (float)state[k]; else e->sd[k] = state[k]; }
    if (aux) std::memcpy(e->aux.data(), aux, sizeof(int32_t) * e->aux.size());
    e->t = t;
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    Block b = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    std::memcpy(out, b.w, sizeof(b.w));
}
void oracle_draw(uint64_t seed, uint32_t env_id, uint64_t index, uint32_t stream, uint32_t sub, uint32_t out[4]) {
    Block b = draw(seed, env_id, index, stream, sub);
    std::memcpy(out, b.w, sizeof(b.w));
}
void oracle_sincosf(const float* x, float* s, float* c, size_t n) {
    for (size_t i = 0; i < n; ++i) det::sincosf_det(x[i], &s[i], &c[i]);
}

}  // extern "C"

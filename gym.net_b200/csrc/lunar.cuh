// Env traits of LunarLander for the generic step / rollout / reset kernels (kernels.cuh), on top of
// the per-lander physics in lunar_core.cuh.
#pragma once
#include "env_classic.cuh"
#include "lunar_core.cuh"

namespace gymcuda {

namespace lunar {

constexpr int SD = 21 + 8 + 4 * MAXC + CHUNKS + 1 + 3 + 12;   // float words per lander in HBM
constexpr int AUXD = 3 + 1 + 2 + 3 * MAXC + 2 + 3;        // int32 words per lander in HBM

// field-major gather / scatter: word f of env i lives at base[f * n + i]
// SLOTS = false: the lander is known to have no contact pair, so its contact slots are empty and are neither read nor written
template <bool SLOTS = true>
__device__ __forceinline__ void load_lander(Lander& L, const float* __restrict__ s, const int32_t* __restrict__ a, size_t n, size_t i) {
    int k = 0;
#define LD() s[(size_t)(k++) * n + i]
    for (int b = 0; b < 3; ++b) { Body& B = L.b[b]; B.c.x = LD(); B.c.y = LD(); B.a = LD(); B.v.x = LD(); B.v.y = LD(); B.w = LD(); B.sleep_time = LD(); }
    for (int j = 0; j < 2; ++j) { L.j[j].ix = LD(); L.j[j].iy = LD(); L.j[j].iz = LD(); L.j[j].motor = LD(); }
    if (SLOTS) { for (int c = 0; c < MAXC; ++c) { L.c[c].ni[0] = LD(); L.c[c].ti[0] = LD(); L.c[c].ni[1] = LD(); L.c[c].ti[1] = LD(); } }
    else { k += 4 * MAXC; for (int c = 0; c < MAXC; ++c) { L.c[c].ni[0] = L.c[c].ti[0] = L.c[c].ni[1] = L.c[c].ti[1] = 0.0f; } }
    for (int t = 0; t < CHUNKS; ++t) L.terrain[t] = LD();
    L.prev_shaping = LD(); L.force.x = LD(); L.force.y = LD(); L.torque = LD();
    for (int b = 0; b < 3; ++b) for (int j = 0; j < 4; ++j) L.fat[b][j] = LD();
#undef LD
    k = 0;
#define LA() a[(size_t)(k++) * n + i]
    if (SLOTS) { for (int b = 0; b < 3; ++b) L.touch[b] = (uint32_t)LA(); } else { k += 3; L.touch[0] = L.touch[1] = L.touch[2] = 0u; }
    L.flags = LA(); L.j[0].limit_state = LA(); L.j[1].limit_state = LA();
    if (SLOTS) { for (int c = 0; c < MAXC; ++c) { L.c[c].pair = LA(); L.c[c].key[0] = (uint32_t)LA(); L.c[c].key[1] = (uint32_t)LA(); } }
    else { k += 3 * MAXC; for (int c = 0; c < MAXC; ++c) { L.c[c].pair = -1; L.c[c].key[0] = L.c[c].key[1] = NO_KEY; } }
    L.wind_idx = LA(); L.torque_idx = LA();
    for (int w = 0; w < 3; ++w) L.pairs[w] = (uint32_t)LA();
#undef LA
}

template <bool SLOTS = true>
__device__ __forceinline__ void store_lander(const Lander& L, float* __restrict__ s, int32_t* __restrict__ a, size_t n, size_t i) {
    int k = 0;
#define ST(v) s[(size_t)(k++) * n + i] = (v)
    for (int b = 0; b < 3; ++b) { const Body& B = L.b[b]; ST(B.c.x); ST(B.c.y); ST(B.a); ST(B.v.x); ST(B.v.y); ST(B.w); ST(B.sleep_time); }
    for (int j = 0; j < 2; ++j) { ST(L.j[j].ix); ST(L.j[j].iy); ST(L.j[j].iz); ST(L.j[j].motor); }
    if (SLOTS) { for (int c = 0; c < MAXC; ++c) { ST(L.c[c].ni[0]); ST(L.c[c].ti[0]); ST(L.c[c].ni[1]); ST(L.c[c].ti[1]); } } else k += 4 * MAXC;
    for (int t = 0; t < CHUNKS; ++t) ST(L.terrain[t]);
    ST(L.prev_shaping); ST(L.force.x); ST(L.force.y); ST(L.torque);
    for (int b = 0; b < 3; ++b) for (int j = 0; j < 4; ++j) ST(L.fat[b][j]);
#undef ST
    k = 0;
#define SA(v) a[(size_t)(k++) * n + i] = (int32_t)(v)
    if (SLOTS) { for (int b = 0; b < 3; ++b) SA(L.touch[b]); } else k += 3;
    SA(L.flags); SA(L.j[0].limit_state); SA(L.j[1].limit_state);
    if (SLOTS) { for (int c = 0; c < MAXC; ++c) { SA(L.c[c].pair); SA(L.c[c].key[0]); SA(L.c[c].key[1]); } } else k += 3 * MAXC;
    SA(L.wind_idx); SA(L.torque_idx);
    for (int w = 0; w < 3; ++w) SA(L.pairs[w]);
#undef SA
}

}  // namespace lunar

// ------------------------------------------------------------------------------------------------
// LunarLander-v2, discrete (Discrete(4), LunarLanderEnv.cs:421) and continuous (Box(-1, 1, (2,)), :417)
// ------------------------------------------------------------------------------------------------
// HAS_PAIRS = false: the traits of the free-flight class of the contact partition (kernels.cuh)
// TRIO_ = true (contact class only): three lanes per lander inside the solver's iteration loops (lunar_core.cuh "TRIO");
// step_kernel maps ten landers to a warp and lets the first lane of each trio do the global stores
template <bool CONTINUOUS, bool HAS_PAIRS = true, bool TRIO_ = false>
struct LunarLanderT {
    static constexpr bool TRIO = TRIO_;
    static constexpr int SD = lunar::SD, AUX = lunar::AUXD + 2, AUXW = lunar::AUXD;
    static constexpr int OD = 8, AD = CONTINUOUS ? 2 : 1, ACTN = CONTINUOUS ? 0 : 4, DEFAULT_LIMIT = 0;
    static constexpr bool HAS_SBD = false;
    static constexpr bool REJECT_INVALID = true;    // InvalidActionError (LunarLanderEnv.cs:604-607)
    static constexpr bool HAS_SMALL = false;
    static constexpr bool FUSED_RESET = true;       // step_autoreset: a crash and the zero step of the next episode share one solve
    static constexpr bool ROLLOUT_CHUNK = false;    // one step is thousands of instructions: no unrolling
    static constexpr bool PREGEN_RESET = false;     // resets are rare and a full zero step: done in place
    static constexpr float ACT_LOW = -1.0f, ACT_HIGH = 1.0f;
    using Act = typename std::conditional<CONTINUOUS, float2, int32_t>::type;
    using S = lunar::Lander;

    __device__ static __forceinline__ S load(const void* base, const int32_t* aux, int n, int i, const EnvParams& prm) {
        S L;
        lunar::load_lander<HAS_PAIRS>(L, reinterpret_cast<const float*>(base), aux, (size_t)n, (size_t)i);
        L.gravity = prm.gravity; L.use_wind = prm.use_wind; L.wind_power = prm.wind_power; L.turbulence_power = prm.turbulence_power;
        return L;
    }
    __device__ static __forceinline__ void store(void* base, int32_t* aux, int n, int i, const S& L) {
        lunar::store_lander<HAS_PAIRS>(L, reinterpret_cast<float*>(base), aux, (size_t)n, (size_t)i);
    }
    __device__ static __forceinline__ void reset(S& L, uint64_t seed, uint32_t gid, uint32_t ordinal, uint64_t t, const EnvParams& prm) {
        lunar::reset(L, seed, gid, (uint64_t)ordinal, CONTINUOUS, t, prm.gravity, prm.use_wind, prm.wind_power, prm.turbulence_power);
    }
    __device__ static __forceinline__ bool valid(int32_t a) { return a >= 0 && a < 4; }
    __device__ static __forceinline__ bool valid(float2 a) { return a.x == a.x && a.y == a.y; }
    __device__ static __forceinline__ StepOut step(S& L, int32_t a, int32_t&, uint64_t seed, uint32_t gid, uint64_t t) {
        const float none[2] = {0.0f, 0.0f};
        const lunar::StepResult r = lunar::step<HAS_PAIRS>(L, seed, gid, t, (int)a, none);
        return StepOut{r.reward, (unsigned)(r.done != 0)};
    }
    __device__ static __forceinline__ StepOut step(S& L, float2 a, int32_t&, uint64_t seed, uint32_t gid, uint64_t t) {
        const float act[2] = {a.x, a.y};
        const lunar::StepResult r = lunar::step<HAS_PAIRS>(L, seed, gid, t, 0, act);
        return StepOut{r.reward, (unsigned)(r.done != 0)};
    }
    // step with the auto-reset folded in (lunar_core.cuh step_autoreset); `ordinal` = the env's next RESET draw index
    __device__ static __forceinline__ StepOut step_ar(S& L, int32_t a, uint64_t seed, uint32_t gid, uint64_t t, bool allow, uint32_t ordinal) {
        const float none[2] = {0.0f, 0.0f};
        const lunar::StepResult r = lunar::step_autoreset<HAS_PAIRS, TRIO_>(L, seed, gid, t, (int)a, none, allow, (uint64_t)ordinal);
        return StepOut{r.reward, (unsigned)(r.done != 0), (unsigned)r.did_reset};
    }
    __device__ static __forceinline__ StepOut step_ar(S& L, float2 a, uint64_t seed, uint32_t gid, uint64_t t, bool allow, uint32_t ordinal) {
        const float act[2] = {a.x, a.y};
        const lunar::StepResult r = lunar::step_autoreset<HAS_PAIRS, TRIO_>(L, seed, gid, t, 0, act, allow, (uint64_t)ordinal);
        return StepOut{r.reward, (unsigned)(r.done != 0), (unsigned)r.did_reset};
    }
    __device__ static __forceinline__ void obs(const S& L, float* o) { lunar::observe(L, o); }
    // LunarLanderEnv ctor (:409-410): _wind_idx / _torque_idx = randint(-9999, 9999), once per generator
    __device__ static __forceinline__ void ctor(void*, int32_t* aux, int n, int i, uint64_t seed, uint32_t gid) {
        const Block b = draw(seed, gid, 0, STREAM_CTOR);
        aux[(size_t)(lunar::AUXD - 5) * n + i] = -9999 + (int32_t)__umulhi(b.w0, 19998u);
        aux[(size_t)(lunar::AUXD - 4) * n + i] = -9999 + (int32_t)__umulhi(b.w1, 19998u);
    }
};
using LunarLander = LunarLanderT<false>;
using LunarLanderCont = LunarLanderT<true>;

}  // namespace gymcuda

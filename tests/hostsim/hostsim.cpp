// TEST INFRASTRUCTURE ONLY.  The engine's device headers compiled for the HOST: every __device__ function of
// gym.net_b200/csrc (detmath, Philox, the five classic envs, the LunarLander solver) becomes plain C++ through
// stubs/cuda_runtime.h, and the per-thread body of step_kernel / reset_kernel (kernels.cuh) is replayed here
// one env at a time.  tests/test_hostsim_cpu.py compares it bit for bit with the CPU oracle: the kernel SOURCE
// is then checked on machines without a GPU.  It proves nothing about speed and is never part of the product.
//
//   g++ -O2 -std=c++17 -ffp-contract=off -Itests/hostsim/stubs -Igym.net_b200/csrc -shared -fPIC ...
#include <cuda_runtime.h>   // the stub

#include "env_classic.cuh"
#include "lunar.cuh"
#include "kernels.cuh"

using namespace gymcuda;

namespace {

struct Bufs {
    void* state; int32_t* aux; int32_t* sbd; int32_t* ep_t; int32_t* episode;
};

template <class E, class ActT>
int sim_step(const Bufs& b, const ActT* actions, float* obs, float* reward, uint8_t* done, int n, uint32_t env_off,
             uint64_t seed, uint64_t t, int limit, int auto_reset, const EnvParams& prm) {
    int invalid = 0;
    for (int i = 0; i < n; ++i) {
        typename E::S s = E::load(b.state, b.aux, n, i, prm);
        const typename E::Act a = actions[i];
        int32_t sbd = (E::HAS_SBD && !auto_reset) ? b.sbd[i] : -1;
        int32_t ept = limit > 0 ? b.ep_t[i] : 0;
        StepOut r{0.0f, 0u};
        if (E::REJECT_INVALID && !E::valid(a)) {
            invalid += 1;
        } else {
            const uint32_t gid = env_off + (uint32_t)i;
            r = E::step(s, a, sbd, seed, gid, t);
            if (limit > 0) { ept += 1; if (ept >= limit && !r.done) r.done = 1u; }
            if (auto_reset && r.done) {
                const int32_t ep = b.episode[i];
                E::reset(s, seed, gid, (uint32_t)ep, t + 1, prm);
                b.episode[i] = ep + 1;
                sbd = -1;
                ept = 0;
            }
            E::store(b.state, b.aux, n, i, s);
            if (E::HAS_SBD && !auto_reset) b.sbd[i] = sbd;
            if (limit > 0) b.ep_t[i] = ept;
        }
        float o[E::OD];
        E::obs(s, o);
        for (int k = 0; k < E::OD; ++k) obs[(size_t)i * E::OD + k] = o[k];
        reward[i] = r.reward;
        done[i] = (uint8_t)(r.done != 0);
    }
    return invalid;
}

template <class E>
void sim_reset(const Bufs& b, float* obs, int n, uint32_t env_off, uint64_t seed, uint64_t t, const EnvParams& prm) {
    for (int i = 0; i < n; ++i) {
        typename E::S s = E::load(b.state, b.aux, n, i, prm);
        const int32_t ep = b.episode[i];
        E::reset(s, seed, env_off + (uint32_t)i, (uint32_t)ep, t, prm);
        E::store(b.state, b.aux, n, i, s);
        b.episode[i] = ep + 1;
        b.sbd[i] = -1;
        b.ep_t[i] = 0;
        float o[E::OD];
        E::obs(s, o);
        for (int k = 0; k < E::OD; ++k) obs[(size_t)i * E::OD + k] = o[k];
    }
}

}  // namespace

// The rollout kernel itself (kernels.cuh), executed one thread at a time: a "warp" here has one lane, so warp votes are
// per-thread and the staged observation store (which needs the 32 lanes of a warp side by side) must be off: callers
// pass n % 4 != 0 for 3- and 6-float observations.  Everything else is the kernel's own code: action generator, head /
// unrolled chunks / tail, reduced-range and limit-free chunk variants, pre-generated resets, 32-bit row index.
template <class E, bool AR, bool LIM, bool ALL_OUT, int BLOCK>
static void run_rollout(const RolloutArgs& a) {
    const int grid = (a.n + BLOCK - 1) / BLOCK;
    gridDim.x = (unsigned)grid; blockDim.x = (unsigned)BLOCK;
    for (int b = 0; b < grid; ++b)
        for (int tid = 0; tid < BLOCK; ++tid) {
            blockIdx.x = (unsigned)b; threadIdx.x = (unsigned)tid;
            rollout_kernel<E, AR, LIM, ALL_OUT, BLOCK>(a);
        }
}

template <class E>
static int rollout_dispatch(const RolloutArgs& a, int auto_reset, int all_out, int block) {
    const bool lim = a.limit > 0;
#define HS_CASE(AR, LIM, AO, BL) if (auto_reset == AR && lim == LIM && all_out == AO && block == BL) { run_rollout<E, AR, LIM, AO, BL>(a); return 0; }
    HS_CASE(1, 0, 1, 64) HS_CASE(1, 1, 1, 64) HS_CASE(1, 0, 0, 64) HS_CASE(1, 1, 0, 64) HS_CASE(0, 0, 0, 64) HS_CASE(0, 1, 0, 64)
    HS_CASE(0, 0, 1, 64) HS_CASE(0, 1, 1, 64)
    if constexpr (E::ROLLOUT_CHUNK) { HS_CASE(1, 0, 1, 512) HS_CASE(1, 1, 1, 512) }
#undef HS_CASE
    return -2;
}


extern "C" {

// kinds as in include/gymcuda.h: 0 CartPole, 1 Pendulum, 2 MountainCar, 3 MountainCarContinuous, 4 Acrobot,
// 5 LunarLander, 6 LunarLanderContinuous.  Layouts are the DEVICE layouts (kernels.cuh header comment).
int hostsim_step(int kind, void* state, int32_t* aux, int32_t* sbd, int32_t* ep_t, int32_t* episode, const void* actions,
                 float* obs, float* reward, uint8_t* done, int n, uint32_t env_off, uint64_t seed, uint64_t t, int limit,
                 int auto_reset, float gravity, float wind_power, float turbulence_power, int use_wind) {
    const Bufs b{state, aux, sbd, ep_t, episode};
    const EnvParams prm{gravity, wind_power, turbulence_power, use_wind};
    switch (kind) {
        case 0: return sim_step<CartPole>(b, (const int32_t*)actions, obs, reward, done, n, env_off, seed, t, limit, auto_reset, prm);
        case 1: return sim_step<Pendulum>(b, (const float*)actions, obs, reward, done, n, env_off, seed, t, limit, auto_reset, prm);
        case 2: return sim_step<MountainCar>(b, (const int32_t*)actions, obs, reward, done, n, env_off, seed, t, limit, auto_reset, prm);
        case 3: return sim_step<MountainCarCont>(b, (const float*)actions, obs, reward, done, n, env_off, seed, t, limit, auto_reset, prm);
        case 4: return sim_step<Acrobot>(b, (const int32_t*)actions, obs, reward, done, n, env_off, seed, t, limit, auto_reset, prm);
        case 5: return sim_step<LunarLander>(b, (const int32_t*)actions, obs, reward, done, n, env_off, seed, t, limit, auto_reset, prm);
        case 6: return sim_step<LunarLanderCont>(b, (const float2*)actions, obs, reward, done, n, env_off, seed, t, limit, auto_reset, prm);
    }
    return -1;
}

int hostsim_reset(int kind, void* state, int32_t* aux, int32_t* sbd, int32_t* ep_t, int32_t* episode, float* obs, int n,
                  uint32_t env_off, uint64_t seed, uint64_t t, float gravity, float wind_power, float turbulence_power, int use_wind) {
    const Bufs b{state, aux, sbd, ep_t, episode};
    const EnvParams prm{gravity, wind_power, turbulence_power, use_wind};
    switch (kind) {
        case 0: sim_reset<CartPole>(b, obs, n, env_off, seed, t, prm); return 0;
        case 1: sim_reset<Pendulum>(b, obs, n, env_off, seed, t, prm); return 0;
        case 2: sim_reset<MountainCar>(b, obs, n, env_off, seed, t, prm); return 0;
        case 3: sim_reset<MountainCarCont>(b, obs, n, env_off, seed, t, prm); return 0;
        case 4: sim_reset<Acrobot>(b, obs, n, env_off, seed, t, prm); return 0;
        case 5: sim_reset<LunarLander>(b, obs, n, env_off, seed, t, prm); return 0;
        case 6: sim_reset<LunarLanderCont>(b, obs, n, env_off, seed, t, prm); return 0;
    }
    return -1;
}

// LunarLander constructor draws (wind phase): ctor_kernel
int hostsim_ctor(int kind, void* state, int32_t* aux, int n, uint32_t env_off, uint64_t seed) {
    if (kind != 5 && kind != 6) return 0;
    for (int i = 0; i < n; ++i) {
        if (kind == 5) LunarLander::ctor(state, aux, n, i, seed, env_off + (uint32_t)i);
        else LunarLanderCont::ctor(state, aux, n, i, seed, env_off + (uint32_t)i);
    }
    return 0;
}

// the division core on its own: div_inrange(x, y) for arrays (the host reciprocal estimate is 1.0f / y)
void hostsim_div_inrange(const float* x, const float* y, float* q, size_t n) {
    for (size_t i = 0; i < n; ++i) q[i] = div_inrange(x[i], y[i]);
}

void hostsim_sincos(const float* x, float* s, float* c, size_t n) {
    for (size_t i = 0; i < n; ++i) sincosf_det(x[i], &s[i], &c[i]);
}

int hostsim_rollout(int kind, void* state, int32_t* aux, int32_t* sbd, int32_t* ep_t, int32_t* episode, float* obs, float* reward,
                    uint8_t* done, void* actions, unsigned long long* stats, int n, int k_steps, uint32_t env_off, uint64_t seed,
                    uint64_t t, int limit, int auto_reset, int all_out, int block, float gravity, float wind_power,
                    float turbulence_power, int use_wind) {
    RolloutArgs a{};
    a.state = state; a.aux = aux; a.sbd = sbd; a.ep_t = ep_t; a.episode = episode; a.seeds = nullptr; a.perm = nullptr;
    a.obs = obs; a.reward = reward; a.done = done; a.actions = actions; a.stats = stats; a.ep_ret = nullptr; a.sums = nullptr;
    a.done_bits = 0; a.n = n; a.k_steps = k_steps; a.env_off = env_off; a.seed = seed; a.t = t; a.limit = limit;
    a.prm = EnvParams{gravity, wind_power, turbulence_power, use_wind};
    switch (kind) {
        case 0: return rollout_dispatch<CartPole>(a, auto_reset, all_out, block);
        case 1: return rollout_dispatch<Pendulum>(a, auto_reset, all_out, block);
        case 2: return rollout_dispatch<MountainCar>(a, auto_reset, all_out, block);
        case 3: return rollout_dispatch<MountainCarCont>(a, auto_reset, all_out, block);
        case 4: return rollout_dispatch<Acrobot>(a, auto_reset, all_out, block);
        case 5: return rollout_dispatch<LunarLander>(a, auto_reset, all_out, block);
        case 6: return rollout_dispatch<LunarLanderCont>(a, auto_reset, all_out, block);
    }
    return -1;
}

}  // extern "C"

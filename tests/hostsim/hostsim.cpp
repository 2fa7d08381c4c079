// TEST INFRASTRUCTURE ONLY.  The engine's device headers compiled for the HOST: every __device__ function of
// gym.net_b200/csrc (detmath, Philox, the five classic envs, the LunarLander solver) becomes plain C++ through
// stubs/cuda_runtime.h, and the per-thread body of step_kernel / reset_kernel (kernels.cuh) is replayed here
// one env at a time; the kernels of kernels.cuh themselves are launched either one thread at a time or, through
// simt.hpp, as CTAs of fibers with real warp / block primitives.  tests/test_hostsim_cpu.py and
// tests/test_hostsim_simt_cpu.py compare all of it bit for bit with the CPU oracle: the kernel SOURCE is then checked
// on machines without a GPU.  It proves nothing about speed and is never part of the product.
//
//   g++ -O2 -std=c++17 -ffp-contract=off -Itests/hostsim/stubs -Igym.net_b200/csrc -shared -fPIC ...
#include <cuda_runtime.h>   // the stub

#include <functional>
#include <vector>

#include "env_classic.cuh"
#include "lunar.cuh"
#include "kernels.cuh"
#include "normalize.cuh"

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

// Kernel launches on the host, two ways (g_simt, set by hostsim_set_simt):
//   0  one thread at a time: a "warp" has one lane, so warp votes are per-thread and the staged observation store of
//      the rollout kernel (which needs the 32 lanes of a warp side by side) must be off -- callers pass n % 4 != 0 for
//      3- and 6-float observations.  Fast; everything per-thread is the kernel's own code.
//   1  simt.hpp: the threads of a CTA are fibers, ballots / votes / shuffles / barriers are real rendezvous, `static`
//      stands in for __shared__.  The kernels run as written, warp-cooperative parts included.
static int g_simt = 0;
static const int32_t* g_seeds = nullptr;
static float* g_terminal_obs = nullptr;
static const void* g_actions_in = nullptr;
static int* g_rollout_invalid = nullptr;
static void set_block(unsigned b, unsigned grid, unsigned block) { blockIdx.x = b; gridDim.x = grid; blockDim.x = block; }
static void set_thread(unsigned t) { threadIdx.x = t; }

template <class F>
static void launch(int grid, int block, F&& kernel_call) {
    if (g_simt) { simt::launch(grid, block, set_block, set_thread, std::function<void()>(kernel_call)); return; }
    for (int b = 0; b < grid; ++b)
        for (int tid = 0; tid < block; ++tid) { set_block((unsigned)b, (unsigned)grid, (unsigned)block); set_thread((unsigned)tid); kernel_call(); }
}

template <class E, bool AR, bool LIM, bool ALL_OUT, int BLOCK>
static void run_rollout(const RolloutArgs& a) {
    // all_out + actions_in: the SUPPLIED variant (gymcuda_step_many* with every output requested)
    if constexpr (ALL_OUT) {
        if (a.actions_in) { launch((a.n + BLOCK - 1) / BLOCK, BLOCK, [&] { rollout_kernel<E, AR, LIM, true, BLOCK, true>(a); }); return; }
    }
    launch((a.n + BLOCK - 1) / BLOCK, BLOCK, [&] { rollout_kernel<E, AR, LIM, ALL_OUT, BLOCK>(a); });
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


static int g_trio = 0;   // LunarLander step launches run as the TRIO variant (three lanes per lander; SIMT executor only)

template <class E>
static int k_step(const StepArgs& a, int auto_reset) {
    int grid = (a.n + STEP_BLOCK - 1) / STEP_BLOCK;
    if constexpr (TrioEnv<E>::value) grid = ((a.n + TRIO_ENVS_PER_WARP - 1) / TRIO_ENVS_PER_WARP * 32 + STEP_BLOCK - 1) / STEP_BLOCK;
    const bool lim = a.limit > 0;
    if (auto_reset && lim) launch(grid, STEP_BLOCK, [&] { step_kernel<E, true, true>(a); });
    else if (auto_reset) launch(grid, STEP_BLOCK, [&] { step_kernel<E, true, false>(a); });
    else if (lim) launch(grid, STEP_BLOCK, [&] { step_kernel<E, false, true>(a); });
    else launch(grid, STEP_BLOCK, [&] { step_kernel<E, false, false>(a); });
    return 0;
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
        const uint64_t sd = seed_of(g_seeds, seed, i);
        if (kind == 5) LunarLander::ctor(state, aux, n, i, sd, env_off + (uint32_t)i);
        else LunarLanderCont::ctor(state, aux, n, i, sd, env_off + (uint32_t)i);
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
                    uint8_t* done, void* actions, unsigned long long* stats, float* ep_ret, double* sums, int done_bits, int n,
                    int k_steps, uint32_t env_off, uint64_t seed, uint64_t t, int limit, int auto_reset, int all_out, int block,
                    float gravity, float wind_power, float turbulence_power, int use_wind) {
    RolloutArgs a{};
    a.state = state; a.aux = aux; a.sbd = sbd; a.ep_t = ep_t; a.episode = episode; a.seeds = g_seeds; a.perm = nullptr;
    a.obs = obs; a.reward = reward; a.done = done; a.actions = actions; a.stats = stats; a.ep_ret = ep_ret; a.sums = sums;
    a.done_bits = done_bits; a.n = n; a.k_steps = k_steps; a.env_off = env_off; a.seed = seed; a.t = t; a.limit = limit;
    a.prm = EnvParams{gravity, wind_power, turbulence_power, use_wind};
    a.actions_in = g_actions_in; a.host_invalid = g_rollout_invalid; g_actions_in = nullptr; g_rollout_invalid = nullptr;
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


// ---- the kernels themselves (not their per-thread bodies): meant for hostsim_set_simt(1) --------------------------
// per-env seeds of VecEnv.Seed(int[]) (VecEnv.cs:48-53) for every kernel launched from now on; null = the handle's one seed
void hostsim_set_seeds(const int32_t* seeds) { g_seeds = seeds; }
void hostsim_set_simt(int on) { g_simt = on ? 1 : 0; }
void hostsim_set_trio(int on) { g_trio = on ? 1 : 0; }
// thread scheduling order of the SIMT executor: 0 ascending, 1 descending, 2 pseudo-random (seeded)
void hostsim_set_schedule(int policy, uint64_t seed) { simt::g_policy = policy; simt::g_rng = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull; }

// The NEXT hostsim_step_kernel call runs as rank `rank` of a fused step + observation gather (gymcuda_step_gather_device):
// peer_obs[r] / peer_flags[r] stand for the cudaIpc-mapped gather buffer and arrival flags of rank r.
static struct { int world, rank; uint32_t gseq; float* peer_obs[MAX_PEERS]; uint32_t* peer_flags[MAX_PEERS]; unsigned* block_counter; } g_gather;
void hostsim_set_gather(int world, int rank, uint32_t gseq, float** peer_obs, uint32_t** peer_flags, unsigned* block_counter) {
    g_gather.world = world; g_gather.rank = rank; g_gather.gseq = gseq; g_gather.block_counter = block_counter;
    for (int r = 0; r < world; ++r) { g_gather.peer_obs[r] = peer_obs[r]; g_gather.peer_flags[r] = peer_flags[r]; }
}

// The NEXT hostsim_step_kernel call writes terminal observations here (gymcuda_set_terminal_obs); the NEXT hostsim_rollout
// call reads its actions from `actions_in` (gymcuda_step_many*) and raises *host_invalid when it rejects one.
void hostsim_set_terminal_obs(float* buf) { g_terminal_obs = buf; }
void hostsim_set_actions_in(const void* actions_in, int* host_invalid) { g_actions_in = actions_in; g_rollout_invalid = host_invalid; }

// gather_wait_kernel: returns the timeout flag (0, or 1 + the rank that never published)
int hostsim_gather_wait(const uint32_t* flags, int world, uint32_t gseq) {
    int timeout_flag = 0;
    launch(1, 32, [&] { gather_wait_kernel(flags, world, gseq, &timeout_flag); });
    return timeout_flag;
}

// step_kernel as launched by gymcuda_step_device: done compaction, statistics, packed done bytes, optional perm
int hostsim_step_kernel(int kind, void* state, int32_t* aux, int32_t* sbd, int32_t* ep_t, int32_t* episode, const int32_t* perm,
                        const void* actions, float* obs, float* reward, uint8_t* done, int32_t* done_idx, int32_t* done_count,
                        unsigned long long* stats, float* ep_ret, double* sums, int done_bits, int* host_invalid, int n,
                        uint32_t env_off, uint64_t seed, uint64_t t, int limit, int auto_reset, int use_bcast, int32_t bcast_action,
                        uint32_t seq, float gravity, float wind_power, float turbulence_power, int use_wind) {
    StepArgs a{};
    a.state = state; a.aux = aux; a.sbd = sbd; a.ep_t = ep_t; a.episode = episode; a.seeds = g_seeds; a.perm = perm; a.actions = actions;
    a.obs = obs; a.reward = reward; a.done = done; a.done_idx = done_idx; a.done_count = done_count; a.stats = stats; a.ep_ret = ep_ret;
    a.sums = sums; a.done_bits = done_bits; a.host_invalid = host_invalid; a.n = n; a.env_off = env_off; a.seed = seed; a.t = t;
    a.limit = limit; a.use_bcast = use_bcast; a.bcast_action = bcast_action; a.seq = seq; a.fold_prev = 1;
    a.prm = EnvParams{gravity, wind_power, turbulence_power, use_wind};
    a.world = g_gather.world; a.rank = g_gather.rank; a.gseq = g_gather.gseq; a.block_counter = g_gather.block_counter;
    for (int r = 0; r < g_gather.world; ++r) { a.peer_obs[r] = g_gather.peer_obs[r]; a.peer_flags[r] = g_gather.peer_flags[r]; }
    const bool block_epilogue = kind < 5 || g_gather.world > 0;   // LunarLander without the fused gather: warp-granular epilogue, writes done_idx itself
    g_gather.world = 0;   // one launch
    a.terminal_obs = g_terminal_obs; g_terminal_obs = nullptr;
    // the CTA-level epilogue leaves per-CTA counts and sub-lists; the done list is built from them afterwards, as gymcuda_done_indices does
    const int nb = (n + STEP_BLOCK - 1) / STEP_BLOCK;
    std::vector<int32_t> blk_cnt((size_t)nb + 1, 0), tmp_idx((size_t)nb * STEP_BLOCK, -1);
    a.blk_cnt = blk_cnt.data(); a.tmp_idx = tmp_idx.data();
    int rc = -1;
    switch (kind) {
        case 0: rc = k_step<CartPole>(a, auto_reset); break;
        case 1: rc = k_step<Pendulum>(a, auto_reset); break;
        case 2: rc = k_step<MountainCar>(a, auto_reset); break;
        case 3: rc = k_step<MountainCarCont>(a, auto_reset); break;
        case 4: rc = k_step<Acrobot>(a, auto_reset); break;
        case 5: rc = g_trio ? k_step<LunarLanderT<false, true, true>>(a, auto_reset) : k_step<LunarLander>(a, auto_reset); break;
        case 6: rc = g_trio ? k_step<LunarLanderT<true, true, true>>(a, auto_reset) : k_step<LunarLanderCont>(a, auto_reset); break;
    }
    if (rc == 0 && block_epilogue && done_idx != nullptr) {
        int32_t* cnt = blk_cnt.data(); const int32_t* tmp = tmp_idx.data();
        launch(1, 1024, [&] { partition_scan_kernel(cnt, nb); });
        launch((nb + 7) / 8, 256, [&] { done_list_scatter_kernel(cnt, tmp, nb, STEP_BLOCK, done_idx); });
    }
    return rc;
}

// reset_kernel (mask may be null = all)
int hostsim_reset_kernel(int kind, void* state, int32_t* aux, int32_t* sbd, int32_t* ep_t, int32_t* episode, const uint8_t* mask,
                         float* obs, float* ep_ret, int n, uint32_t env_off, uint64_t seed, uint64_t t, float gravity,
                         float wind_power, float turbulence_power, int use_wind) {
    ResetArgs a{};
    a.state = state; a.aux = aux; a.sbd = sbd; a.ep_t = ep_t; a.episode = episode; a.seeds = g_seeds; a.mask = mask; a.obs = obs;
    a.ep_ret = ep_ret; a.n = n; a.env_off = env_off; a.seed = seed; a.t = t; a.prm = EnvParams{gravity, wind_power, turbulence_power, use_wind};
    const int grid = (n + 127) / 128;
    switch (kind) {
        case 0: launch(grid, 128, [&] { reset_kernel<CartPole>(a); }); return 0;
        case 1: launch(grid, 128, [&] { reset_kernel<Pendulum>(a); }); return 0;
        case 2: launch(grid, 128, [&] { reset_kernel<MountainCar>(a); }); return 0;
        case 3: launch(grid, 128, [&] { reset_kernel<MountainCarCont>(a); }); return 0;
        case 4: launch(grid, 128, [&] { reset_kernel<Acrobot>(a); }); return 0;
        case 5: launch(grid, 128, [&] { reset_kernel<LunarLander>(a); }); return 0;
        case 6: launch(grid, 128, [&] { reset_kernel<LunarLanderCont>(a); }); return 0;
    }
    return -1;
}

// sample_kernel: ActionSpace.Sample per env (mask: Discrete only, [n][ACTN] bytes, may be null)
int hostsim_sample_kernel(int kind, const uint8_t* mask, void* out, int n, uint32_t env_off, uint64_t seed, uint64_t t) {
    SampleArgs a{g_seeds, mask, out, n, env_off, seed, t};
    const int grid = (n + 127) / 128;
    switch (kind) {
        case 0: launch(grid, 128, [&] { sample_kernel<CartPole>(a); }); return 0;
        case 1: launch(grid, 128, [&] { sample_kernel<Pendulum>(a); }); return 0;
        case 2: launch(grid, 128, [&] { sample_kernel<MountainCar>(a); }); return 0;
        case 3: launch(grid, 128, [&] { sample_kernel<MountainCarCont>(a); }); return 0;
        case 4: launch(grid, 128, [&] { sample_kernel<Acrobot>(a); }); return 0;
        case 5: launch(grid, 128, [&] { sample_kernel<LunarLander>(a); }); return 0;
        case 6: launch(grid, 128, [&] { sample_kernel<LunarLanderCont>(a); }); return 0;
    }
    return -1;
}

// the three launches of contact_partition() (gymcuda.cu): aux field-major, block_free [nb + 1] scratch, perm [n] out
int hostsim_partition(const int32_t* aux, int n, int32_t* block_free, int32_t* perm) {
    const int nb = (n + PART_BLOCK - 1) / PART_BLOCK;
    launch(nb, PART_BLOCK, [&] { partition_count_kernel(aux, n, block_free); });
    launch(1, 1024, [&] { partition_scan_kernel(block_free, nb); });
    launch(nb, PART_BLOCK, [&] { partition_scatter_kernel(aux, n, block_free, nb, perm); });
    return nb;
}


// normalize.cuh: the two launches of gymcuda_normalize_device
int hostsim_normalize(float* obs, float* reward, const uint8_t* done, float* ret, double* acc, int n, int od, float gamma, float eps,
                      float clip_obs, float clip_reward, int update) {
    NormArgs a{};
    a.obs = obs; a.reward = reward; a.done = done; a.ret = ret; a.acc = acc; a.n = n; a.od = od;
    a.gamma = gamma; a.eps = eps; a.clip_obs = clip_obs; a.clip_reward = clip_reward;
    const int grid = (n + NORM_BLOCK - 1) / NORM_BLOCK;
    if (update) launch(grid, NORM_BLOCK, [&] { norm_update_kernel(a); });
    launch(grid, NORM_BLOCK, [&] { norm_apply_kernel(a); });
    return 0;
}

}  // extern "C"

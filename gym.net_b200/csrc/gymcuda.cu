// libgymcuda: C-ABI implementation (include/gymcuda.h).  Handle lifecycle, device memory,
// kernel dispatch, host<->device copies, NCCL all-gather.  There is no CPU path in this file:
// every compute entry point launches a CUDA kernel or fails with GYMCUDA_ECUDA.
#include "../../include/gymcuda.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

#include <algorithm>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: the ranges cost a null-pointer test unless a profiler injects itself

#include "kernels.cuh"
#include "normalize.cuh"
#include "box_sample.cuh"
#include "render.cuh"
#ifdef GYMCUDA_WITH_LUNAR
#include "lunar_launch.h"
#endif

using namespace gymcuda;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CU_TRY(expr)                                                                                  \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return fail(_e == cudaErrorMemoryAllocation ? GYMCUDA_ENOMEM : GYMCUDA_ECUDA,             \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);  \
    } while (0)

// ------------------------------------------------------------------------------------------------
// NCCL, loaded lazily with dlopen so the library has no link-time dependency on it
// ------------------------------------------------------------------------------------------------
struct Id128 { char b[128]; };   // ncclUniqueId, passed by value
namespace {
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
}  // namespace
static NcclApi g_nccl;
static std::mutex g_nccl_mutex;   // handles may initialise their communicators from different host threads

static int nccl_load(const char* path) {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.lib) return GYMCUDA_OK;
    const char* candidates[] = {path, getenv("GYMCUDA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* c : candidates) {
        if (!c || !*c) continue;
        h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(GYMCUDA_ENCCL, "cannot dlopen NCCL (tried path argument, $GYMCUDA_NCCL_LIB, libnccl.so.2): %s", dlerror());
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) {
        dlclose(h);
        return fail(GYMCUDA_ENCCL, "NCCL library is missing required symbols");
    }
    g_nccl.lib = h;
    return GYMCUDA_OK;
}

#define NCCL_TRY(expr)                                                                              \
    do {                                                                                            \
        int _r = (expr);                                                                            \
        if (_r != 0)                                                                                \
            return fail(GYMCUDA_ENCCL, "%s failed: %s", #expr,                                      \
                        g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error");          \
    } while (0)

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
struct KindInfo { int sd, aux, od, ad, actn, default_limit; };

static bool kind_info(int kind, KindInfo* ki) {
    switch (kind) {
        case GYMCUDA_CARTPOLE: *ki = {CartPole::SD, 3, CartPole::OD, CartPole::AD, CartPole::ACTN, CartPole::DEFAULT_LIMIT}; return true;
        case GYMCUDA_PENDULUM: *ki = {Pendulum::SD, 3, Pendulum::OD, Pendulum::AD, Pendulum::ACTN, Pendulum::DEFAULT_LIMIT}; return true;
        case GYMCUDA_MOUNTAINCAR: *ki = {MountainCar::SD, 3, MountainCar::OD, MountainCar::AD, MountainCar::ACTN, MountainCar::DEFAULT_LIMIT}; return true;
        case GYMCUDA_MOUNTAINCAR_CONT: *ki = {MountainCarCont::SD, 3, MountainCarCont::OD, MountainCarCont::AD, MountainCarCont::ACTN, MountainCarCont::DEFAULT_LIMIT}; return true;
        case GYMCUDA_ACROBOT: *ki = {Acrobot::SD, 3, Acrobot::OD, Acrobot::AD, Acrobot::ACTN, Acrobot::DEFAULT_LIMIT}; return true;
#ifdef GYMCUDA_WITH_LUNAR
        case GYMCUDA_LUNARLANDER: *ki = {LUNAR_STATE_DIM, LUNAR_AUX_DIM, 8, 1, 4, 0}; return true;
        case GYMCUDA_LUNARLANDER_CONT: *ki = {LUNAR_STATE_DIM, LUNAR_AUX_DIM, 8, 2, 0, 0}; return true;
#endif
        default: return false;
    }
}

struct gymcuda_env {
    gymcuda_config cfg;
    KindInfo ki;
    int n, limit;
    bool auto_reset, has_state;
    uint64_t seed, t;
    uint32_t seq;
    cudaStream_t own_stream, stream;
    // state
    void* d_state;
    int32_t *d_sbd, *d_ept, *d_episode, *d_seeds, *d_aux;
    EnvParams prm;
    int32_t *d_perm, *d_block_free;   // LunarLander contact partition
    cudaStream_t side_stream;         // LunarLander: the kernel of the second partition class runs here, concurrently
    cudaStream_t in_stream, out_stream;   // host-buffer k-step calls: action chunks travel in, trajectory chunks travel out while the next chunk is stepped (created on first use)
    std::vector<cudaEvent_t> pipe_events; // their events (grow-only)
    cudaEvent_t ev_fork, ev_join;
    int sm_count;                     // multiprocessors of the device (launch heuristics)
    int auxw;   // int32 words per env in d_aux (LunarLander only)
    // I/O staging for the host-buffer entry points
    void* d_actions;
    uint8_t* d_out;          // obs | reward | done
    float *d_obs, *d_reward;
    uint8_t *d_done, *d_mask, *d_sample_mask;
    void* scratch[4]; size_t scratch_cap[4];   // device scratch of gymcuda_rollout_random (host buffers)
    volatile int* h_invalid; // mapped pinned flags: [0] set by the step kernel when it rejects an action,
    int* d_invalid_flag;     //                      [1] set by gather_wait_kernel on a timeout (1 + missing rank)
    int32_t *d_done_idx, *d_done_count;
    int32_t *d_blk_cnt, *d_blk_off, *d_tmp_idx;   // per-CTA counts (and their exclusive scan) / sub-lists of the last step launch (the done list is built on demand)
    bool done_list_stale;             // the last step launch left blk_cnt / tmp_idx, d_done_idx has not been built from them yet
    unsigned long long* d_stats;
    float* d_ep_ret;          // GYMCUDA_FLAG_EPISODE_STATS
    double* d_sums;
    bool ep_stats, done_bits;
    const float* last_obs;   // device pointer of the most recent observations
    // terminal observations under auto-reset (gymcuda_set_terminal_obs)
    float* term_dev;         // what the step kernel writes: the caller's device / mapped buffer, or d_term_own
    float* term_host;        // pageable caller buffer: d_term_own is copied into it by the host-buffer step calls
    float* d_term_own;
    // observation / reward normalisation (normalize.cuh), allocated by the first call
    double* d_norm_acc; float* d_norm_ret;
    float norm_gamma, norm_eps, norm_clip_obs, norm_clip_reward;
    // pinned scratch: [0..1] stats, [2] done_count
    unsigned long long* h_small;
    unsigned long long invalid_seen, env_steps;
    unsigned long long* d_clock;   // {t, seq} on the device (gymcuda_set_device_clock)
    bool device_clock;
    bool stats_pending;      // the episode count of the latest step launch is still in d_done_count, not in d_stats[0]
    bool async_steps;        // *_device steps were enqueued since the last synchronisation: their rejected actions are not yet reported
    // nccl
    void* comm;
    int rank, world;
    // fused gather over peer memory
    int g_world, g_rank;
    uint32_t g_seq;
    uint8_t* g_local;                 // [2][world][n][od] floats, then flags[world] u32, then block counter
    size_t g_flags_off, g_counter_off;
    void* g_peer[MAX_PEERS];          // mapped base of every rank's g_local (own entry = g_local)
    size_t act_bytes() const { return (size_t)n * ki.ad * 4; }
    size_t obs_bytes() const { return (size_t)n * ki.od * 4; }
};

// NVTX range around an entry point (tracing): visible in Nsight Systems timelines as gymcuda/<name>
namespace {
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace
#define TRACE(name) NvtxRange _nvtx_range("gymcuda/" name)

static int check(const gymcuda_env* e) {
    if (!e) return fail(GYMCUDA_EINVAL, "null gymcuda_env handle");
    return GYMCUDA_OK;
}
#define ENTER(e)                                        \
    do {                                                \
        int _c = check(e);                              \
        if (_c) return _c;                              \
        CU_TRY(cudaSetDevice((e)->cfg.device));         \
    } while (0)

// ------------------------------------------------------------------------------------------------
// kernel dispatch
// ------------------------------------------------------------------------------------------------
template <class E>
static cudaError_t launch_step(gymcuda_env* e, const StepArgs& a) {
    const bool ar = e->auto_reset, lim = e->limit > 0;
    if (e->n >= STEP_BIG_BATCH) {   // a million envs and more (kernels.cuh)
        const int grid = (e->n + STEP_BLOCK_BIG - 1) / STEP_BLOCK_BIG;
        if (ar && lim) step_kernel<E, true, true, STEP_BLOCK_BIG><<<grid, STEP_BLOCK_BIG, 0, e->stream>>>(a);
        else if (ar) step_kernel<E, true, false, STEP_BLOCK_BIG><<<grid, STEP_BLOCK_BIG, 0, e->stream>>>(a);
        else if (lim) step_kernel<E, false, true, STEP_BLOCK_BIG><<<grid, STEP_BLOCK_BIG, 0, e->stream>>>(a);
        else step_kernel<E, false, false, STEP_BLOCK_BIG><<<grid, STEP_BLOCK_BIG, 0, e->stream>>>(a);
        return cudaGetLastError();
    }
    const int grid = (e->n + STEP_BLOCK - 1) / STEP_BLOCK;
    if (ar && lim) step_kernel<E, true, true><<<grid, STEP_BLOCK, 0, e->stream>>>(a);
    else if (ar) step_kernel<E, true, false><<<grid, STEP_BLOCK, 0, e->stream>>>(a);
    else if (lim) step_kernel<E, false, true><<<grid, STEP_BLOCK, 0, e->stream>>>(a);
    else step_kernel<E, false, false><<<grid, STEP_BLOCK, 0, e->stream>>>(a);
    return cudaGetLastError();
}

template <class E, bool ALL_OUT, int BLOCK, bool SUPPLIED = false>
static void launch_rollout_variant(gymcuda_env* e, const RolloutArgs& a) {
    const int grid = (e->n + BLOCK - 1) / BLOCK;
    const bool ar = e->auto_reset, lim = e->limit > 0;
    if (ar && lim) rollout_kernel<E, true, true, ALL_OUT, BLOCK, SUPPLIED><<<grid, BLOCK, 0, e->stream>>>(a);
    else if (ar) rollout_kernel<E, true, false, ALL_OUT, BLOCK, SUPPLIED><<<grid, BLOCK, 0, e->stream>>>(a);
    else if (lim) rollout_kernel<E, false, true, ALL_OUT, BLOCK, SUPPLIED><<<grid, BLOCK, 0, e->stream>>>(a);
    else rollout_kernel<E, false, false, ALL_OUT, BLOCK, SUPPLIED><<<grid, BLOCK, 0, e->stream>>>(a);
}

template <class E>
static cudaError_t launch_rollout(gymcuda_env* e, const RolloutArgs& a) {
    // ALL_OUT addresses the trajectory with one 32-bit row index: every row * envs + env must fit
    const bool idx32 = (unsigned long long)a.k_steps * (unsigned long long)a.n <= 0xffffffffull;
    if (a.actions_in && a.obs && a.reward && a.done && !a.ep_ret && !a.done_bits && idx32) {
        // gymcuda_step_many* with every output requested: the chunked all-outputs kernel reading the caller's actions
        const long long wave = (long long)e->sm_count * ROLLOUT_BLOCK_WAVE;
        bool one_wave = false;
        if constexpr (E::ROLLOUT_CHUNK) one_wave = (long long)e->n <= wave && 4ll * e->n >= 3ll * wave;
        if constexpr (E::ROLLOUT_CHUNK) { if (one_wave) { launch_rollout_variant<E, true, ROLLOUT_BLOCK_WAVE, true>(e, a); return cudaGetLastError(); } }
        launch_rollout_variant<E, true, ROLLOUT_BLOCK, true>(e, a);
    } else if (!a.actions_in && a.obs && a.reward && a.done && a.actions && !a.ep_ret && !a.done_bits && idx32) {
        // one wave of 16-warp CTAs, one per SM, when the batch fits and fills at least 3/4 of the SMs (kernels.cuh)
        const long long wave = (long long)e->sm_count * ROLLOUT_BLOCK_WAVE;
        bool one_wave = false;
        if constexpr (E::ROLLOUT_CHUNK) one_wave = (long long)e->n <= wave && 4ll * e->n >= 3ll * wave;
        if constexpr (E::ROLLOUT_CHUNK) { if (one_wave) { launch_rollout_variant<E, true, ROLLOUT_BLOCK_WAVE>(e, a); return cudaGetLastError(); } }
        launch_rollout_variant<E, true, ROLLOUT_BLOCK>(e, a);
    } else {
        launch_rollout_variant<E, false, ROLLOUT_BLOCK>(e, a);
    }
    return cudaGetLastError();
}

template <class E>
static cudaError_t launch_reset(gymcuda_env* e, const ResetArgs& a) {
    reset_kernel<E><<<(e->n + 127) / 128, 128, 0, e->stream>>>(a);
    return cudaGetLastError();
}

static bool is_lunar(const gymcuda_env* e) { return e->cfg.env_kind == GYMCUDA_LUNARLANDER || e->cfg.env_kind == GYMCUDA_LUNARLANDER_CONT; }
static bool is_lunar_cont(const gymcuda_env* e) { return e->cfg.env_kind == GYMCUDA_LUNARLANDER_CONT; }

#define DISPATCH(FN, ...)                                                                \
    switch (e->cfg.env_kind) {                                                           \
        case GYMCUDA_CARTPOLE: return launch_##FN<CartPole>(e, __VA_ARGS__);             \
        case GYMCUDA_PENDULUM: return launch_##FN<Pendulum>(e, __VA_ARGS__);             \
        case GYMCUDA_MOUNTAINCAR: return launch_##FN<MountainCar>(e, __VA_ARGS__);       \
        case GYMCUDA_MOUNTAINCAR_CONT: return launch_##FN<MountainCarCont>(e, __VA_ARGS__); \
        case GYMCUDA_ACROBOT: return launch_##FN<Acrobot>(e, __VA_ARGS__);               \
        default: return cudaErrorInvalidValue;                                           \
    }

// LunarLander: stable partition of the env ids by "a broad-phase contact pair exists" (kernels.cuh): d_perm lists the
// free-flight landers first, d_block_free[nb] is their number.  The two classes are stepped by two kernels, the second
// on a side stream (fork / join with events: legal under stream capture), so the long dependent chains of the landers
// near the ground overlap the bulk of the batch instead of following it.
static cudaError_t lunar_step(gymcuda_env* e, StepArgs a) {
#ifdef GYMCUDA_WITH_LUNAR
    const bool cont = is_lunar_cont(e), ar = e->auto_reset, lim = e->limit > 0;
    const int grid = (e->n + STEP_BLOCK - 1) / STEP_BLOCK;
    if (a.world > 0 || e->n < 4 * STEP_BLOCK) {   // fused gather (one kernel signals the peers) or a tiny batch: one launch, general kernel
        a.perm = nullptr; a.part = 0;
        return lunar_launch_step(cont, true, ar, lim, grid, e->stream, a);
    }
    const int nb = (e->n + PART_BLOCK - 1) / PART_BLOCK;
    partition_count_kernel<<<nb, PART_BLOCK, 0, e->stream>>>(e->d_aux, e->n, e->d_block_free);
    partition_scan_kernel<<<1, 1024, 0, e->stream>>>(e->d_block_free, nb);
    partition_scatter_kernel<<<nb, PART_BLOCK, 0, e->stream>>>(e->d_aux, e->n, e->d_block_free, nb, e->d_perm);
    a.perm = e->d_perm; a.split = e->d_block_free + nb;
    cudaError_t ce = cudaEventRecord(e->ev_fork, e->stream);
    if (ce != cudaSuccess) return ce;
    if ((ce = cudaStreamWaitEvent(e->side_stream, e->ev_fork, 0)) != cudaSuccess) return ce;
    a.part = 2;   // landers with a contact pair: few, long.  GYMCUDA_LUNAR_TRIO=1: three lanes per lander (lunar_core.cuh; measured: no gain, off)
    static const bool trio = [] { const char* v = getenv("GYMCUDA_LUNAR_TRIO"); return v && v[0] == '1'; }();
    ce = trio ? lunar_launch_step_trio(cont, ar, lim, e->side_stream, a) : lunar_launch_step(cont, true, ar, lim, grid, e->side_stream, a);
    if (ce != cudaSuccess) return ce;
    if ((ce = cudaEventRecord(e->ev_join, e->side_stream)) != cudaSuccess) return ce;
    a.part = 1;   // free flight: the bulk
    if ((ce = lunar_launch_step(cont, false, ar, lim, grid, e->stream, a)) != cudaSuccess) return ce;
    return cudaStreamWaitEvent(e->stream, e->ev_join, 0);
#else
    (void)e; (void)a;
    return cudaErrorInvalidValue;
#endif
}

static cudaError_t dispatch_step(gymcuda_env* e, const StepArgs& a) {
    if (is_lunar(e)) return lunar_step(e, a);
    DISPATCH(step, a)
}
static cudaError_t dispatch_rollout(gymcuda_env* e, const RolloutArgs& a) { DISPATCH(rollout, a) }   // (LunarLander: gymcuda_rollout_random_device loops over steps)
template <class E>
static cudaError_t launch_sample(gymcuda_env* e, const SampleArgs& a) {
    sample_kernel<E><<<(e->n + 127) / 128, 128, 0, e->stream>>>(a);
    return cudaGetLastError();
}
static cudaError_t dispatch_reset(gymcuda_env* e, const ResetArgs& a) {
#ifdef GYMCUDA_WITH_LUNAR
    if (is_lunar(e)) return lunar_launch_reset(is_lunar_cont(e), (e->n + 127) / 128, e->stream, a);
#endif
    DISPATCH(reset, a)
}
static cudaError_t dispatch_sample(gymcuda_env* e, const SampleArgs& a) {
#ifdef GYMCUDA_WITH_LUNAR
    if (is_lunar(e)) return lunar_launch_sample(is_lunar_cont(e), (e->n + 127) / 128, e->stream, a);
#endif
    DISPATCH(sample, a)
}

// constructor draws (LunarLander only): at create and whenever the generator is replaced by Seed()
static cudaError_t dispatch_ctor(gymcuda_env* e, const ResetArgs& a) {
#ifdef GYMCUDA_WITH_LUNAR
    if (is_lunar(e)) return lunar_launch_ctor(is_lunar_cont(e), (e->n + 127) / 128, e->stream, a);
#endif
    (void)e; (void)a;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// library
// ------------------------------------------------------------------------------------------------
extern "C" {

static int clock_pull(gymcuda_env* e);
static int clock_push(gymcuda_env* e);

int gymcuda_version(void) { return GYMCUDA_VERSION; }
const char* gymcuda_last_error(void) { return g_last_error.c_str(); }

int gymcuda_device_count(int* count) {
    if (!count) return fail(GYMCUDA_EINVAL, "count is null");
    CU_TRY(cudaGetDeviceCount(count));
    return GYMCUDA_OK;
}

int gymcuda_config_default(gymcuda_config* cfg, int env_kind, int num_envs) {
    if (!cfg) return fail(GYMCUDA_EINVAL, "cfg is null");
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = (uint32_t)sizeof(*cfg);
    cfg->env_kind = env_kind;
    cfg->num_envs = num_envs;
    cfg->device = 0;
    cfg->seed = 0;
    cfg->flags = 0;
    cfg->time_limit = 0;
    cfg->gravity = -10.0f;          // LunarLanderEnv.cs:351
    cfg->enable_wind = 0;
    cfg->wind_power = 15.0f;        // :353
    cfg->turbulence_power = 1.5f;   // :354
    return GYMCUDA_OK;
}

int gymcuda_destroy(gymcuda_env* e) {
    if (!e) return GYMCUDA_OK;
    cudaSetDevice(e->cfg.device);
    if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
    for (int r = 0; r < e->g_world; ++r) if (e->g_peer[r] && r != e->g_rank) cudaIpcCloseMemHandle(e->g_peer[r]);
    cudaFree(e->g_local);
    if (e->own_stream) cudaStreamSynchronize(e->own_stream);
    if (e->side_stream) { cudaStreamSynchronize(e->side_stream); cudaStreamDestroy(e->side_stream); }
    if (e->in_stream) { cudaStreamSynchronize(e->in_stream); cudaStreamDestroy(e->in_stream); }
    if (e->out_stream) { cudaStreamSynchronize(e->out_stream); cudaStreamDestroy(e->out_stream); }
    for (cudaEvent_t ev : e->pipe_events) cudaEventDestroy(ev);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    cudaFree(e->d_state); cudaFree(e->d_sbd); cudaFree(e->d_ept); cudaFree(e->d_episode); cudaFree(e->d_seeds); cudaFree(e->d_aux); cudaFree(e->d_perm); cudaFree(e->d_block_free);
    cudaFree(e->d_actions); cudaFree(e->d_out); cudaFree(e->d_mask); cudaFree(e->d_sample_mask);
    if (e->h_invalid) cudaFreeHost((void*)e->h_invalid);
    for (int k = 0; k < 4; ++k) cudaFree(e->scratch[k]);
    cudaFree(e->d_norm_acc); cudaFree(e->d_norm_ret); cudaFree(e->d_term_own); cudaFree(e->d_clock);
    cudaFree(e->d_done_idx); cudaFree(e->d_done_count); cudaFree(e->d_blk_cnt); cudaFree(e->d_blk_off); cudaFree(e->d_tmp_idx); cudaFree(e->d_stats); cudaFree(e->d_ep_ret); cudaFree(e->d_sums);
    if (e->h_small) cudaFreeHost(e->h_small);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    delete e;
    return GYMCUDA_OK;
}

static int create_impl(const gymcuda_config* cfg, gymcuda_env* e) {
    CU_TRY(cudaSetDevice(cfg->device));
    CU_TRY(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
    CU_TRY(cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, cfg->device));
    e->stream = e->own_stream;
    const size_t n = (size_t)e->n;
    const size_t state_bytes = n * (size_t)e->ki.sd * 4;
    CU_TRY(cudaMalloc(&e->d_state, state_bytes));
    if (e->auxw > 0) {
        CU_TRY(cudaMalloc(&e->d_aux, n * (size_t)e->auxw * 4));
        CU_TRY(cudaMemsetAsync(e->d_aux, 0, n * (size_t)e->auxw * 4, e->stream));
        CU_TRY(cudaMalloc(&e->d_perm, n * 4));
        CU_TRY(cudaMalloc(&e->d_block_free, ((n + PART_BLOCK - 1) / PART_BLOCK + 1) * 4));
        CU_TRY(cudaStreamCreateWithFlags(&e->side_stream, cudaStreamNonBlocking));
        CU_TRY(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
    }
    CU_TRY(cudaMalloc(&e->d_sbd, n * 4));
    CU_TRY(cudaMalloc(&e->d_ept, n * 4));
    CU_TRY(cudaMalloc(&e->d_episode, n * 4));
    CU_TRY(cudaMalloc(&e->d_actions, e->act_bytes()));
    // obs | reward | done live in ONE allocation, in the order of the C ABI's out-parameters, so that a
    // caller whose three host buffers are adjacent gets them with a single DMA
    CU_TRY(cudaMalloc(&e->d_out, e->obs_bytes() + n * 4 + n));
    e->d_obs = reinterpret_cast<float*>(e->d_out);
    e->d_reward = reinterpret_cast<float*>(e->d_out + e->obs_bytes());
    e->d_done = reinterpret_cast<uint8_t*>(e->d_out + e->obs_bytes() + n * 4);
    CU_TRY(cudaHostAlloc((void**)&e->h_invalid, 2 * sizeof(int), cudaHostAllocMapped));
    e->h_invalid[0] = 0; e->h_invalid[1] = 0;
    CU_TRY(cudaHostGetDevicePointer((void**)&e->d_invalid_flag, (void*)e->h_invalid, 0));
    CU_TRY(cudaMalloc(&e->d_mask, n));
    CU_TRY(cudaMalloc(&e->d_done_idx, n * 4));
    CU_TRY(cudaMalloc(&e->d_done_count, 2 * sizeof(int32_t)));
    {
        const size_t nb = (n + STEP_BLOCK - 1) / STEP_BLOCK;
        CU_TRY(cudaMalloc(&e->d_blk_cnt, (nb + 1) * sizeof(int32_t)));
        CU_TRY(cudaMalloc(&e->d_blk_off, (nb + 1) * sizeof(int32_t)));
        CU_TRY(cudaMalloc(&e->d_tmp_idx, nb * STEP_BLOCK * sizeof(int32_t)));
    }
    CU_TRY(cudaMalloc(&e->d_stats, 2 * sizeof(unsigned long long)));
    CU_TRY(cudaHostAlloc((void**)&e->h_small, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
    if (e->ep_stats) {
        CU_TRY(cudaMalloc(&e->d_ep_ret, n * 4));
        CU_TRY(cudaMalloc(&e->d_sums, 2 * sizeof(double)));
        CU_TRY(cudaMemsetAsync(e->d_ep_ret, 0, n * 4, e->stream));
        CU_TRY(cudaMemsetAsync(e->d_sums, 0, 2 * sizeof(double), e->stream));
    }
    CU_TRY(cudaMemsetAsync(e->d_state, 0, state_bytes, e->stream));
    CU_TRY(cudaMemsetAsync(e->d_sbd, 0xff, n * 4, e->stream));
    CU_TRY(cudaMemsetAsync(e->d_ept, 0, n * 4, e->stream));
    CU_TRY(cudaMemsetAsync(e->d_episode, 0, n * 4, e->stream));
    CU_TRY(cudaMemsetAsync(e->d_done_count, 0, 2 * sizeof(int32_t), e->stream));
    CU_TRY(cudaMemsetAsync(e->d_stats, 0, 2 * sizeof(unsigned long long), e->stream));
    CU_TRY(cudaMemsetAsync(e->d_obs, 0, e->obs_bytes(), e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return GYMCUDA_OK;
}

static int run_ctor(gymcuda_env* e) {
    ResetArgs a{};
    a.state = e->d_state; a.aux = e->d_aux; a.prm = e->prm; a.sbd = e->d_sbd; a.ep_t = e->d_ept; a.episode = e->d_episode; a.seeds = e->d_seeds;
    a.n = e->n; a.env_off = e->cfg.env_id_offset; a.seed = e->seed; a.t = e->t;
    CU_TRY(dispatch_ctor(e, a));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return GYMCUDA_OK;
}

int gymcuda_create(const gymcuda_config* cfg, gymcuda_env** out) {
    if (!cfg || !out) return fail(GYMCUDA_EINVAL, "cfg/out is null");
    *out = nullptr;
    if (cfg->struct_size != sizeof(gymcuda_config))
        return fail(GYMCUDA_EINVAL, "gymcuda_config.struct_size %u != %zu (use gymcuda_config_default)", cfg->struct_size, sizeof(gymcuda_config));
    KindInfo ki;
    if (!kind_info(cfg->env_kind, &ki)) return fail(GYMCUDA_EINVAL, "unknown env_kind %d", cfg->env_kind);
    if (cfg->num_envs <= 0) return fail(GYMCUDA_EINVAL, "num_envs must be > 0 (got %d)", cfg->num_envs);
    if ((uint64_t)cfg->env_id_offset + (uint64_t)cfg->num_envs > 0x100000000ull)
        return fail(GYMCUDA_EINVAL, "env_id_offset + num_envs exceeds 2^32");
    // LunarLanderEnv ctor range checks (LunarLanderEnv.cs:396-407)
    if (cfg->gravity < -12.0f || cfg->gravity > 0.0f) return fail(GYMCUDA_EINVAL, "Gravity must be between -12 and 0");
    if (cfg->wind_power < 0.0f || cfg->wind_power > 20.0f) return fail(GYMCUDA_EINVAL, "wind_power value is recommended to be between 0.0 and 20.0");
    if (cfg->turbulence_power < 0.0f || cfg->turbulence_power > 2.0f) return fail(GYMCUDA_EINVAL, "turbulence_power value is recommended to be between 0.0 and 2.0");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(GYMCUDA_ECUDA, "no CUDA device available (%s); libgymcuda has no CPU path", ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(GYMCUDA_EINVAL, "device %d out of range [0, %d)", cfg->device, ndev);

    gymcuda_env* e = new (std::nothrow) gymcuda_env();
    if (!e) return fail(GYMCUDA_ENOMEM, "host allocation failed");
    std::memset(static_cast<void*>(e), 0, sizeof(*e));
    e->cfg = *cfg;
    e->ki = ki;
    e->n = cfg->num_envs;
    e->limit = cfg->time_limit == 0 ? ki.default_limit : (cfg->time_limit < 0 ? 0 : cfg->time_limit);
    e->auto_reset = (cfg->flags & GYMCUDA_FLAG_AUTO_RESET) != 0;
    e->ep_stats = (cfg->flags & GYMCUDA_FLAG_EPISODE_STATS) != 0;
    e->done_bits = (cfg->flags & GYMCUDA_FLAG_DONE_BITS) != 0;
    if (e->ep_stats && e->limit == 0) e->limit = 0x7fffffff;   // the episode-step counter doubles as the episode length
    e->seed = cfg->seed;
    e->last_obs = nullptr;
    e->norm_gamma = 0.99f; e->norm_eps = 1e-8f; e->norm_clip_obs = 10.0f; e->norm_clip_reward = 10.0f;
    e->prm = EnvParams{cfg->gravity, cfg->wind_power, cfg->turbulence_power, cfg->enable_wind ? 1 : 0};
    e->auxw = cfg->env_kind >= GYMCUDA_LUNARLANDER ? ki.aux - 2 : 0;
    int rc = create_impl(cfg, e);
    if (rc == GYMCUDA_OK) rc = run_ctor(e);
    if (rc != GYMCUDA_OK) { std::string keep = g_last_error; gymcuda_destroy(e); g_last_error = keep; return rc; }
    e->last_obs = e->d_obs;
    *out = e;
    return GYMCUDA_OK;
}

int gymcuda_num_envs(const gymcuda_env* e) { return e ? e->n : fail(GYMCUDA_EINVAL, "null handle"); }

int gymcuda_space(const gymcuda_env* e, gymcuda_space_info* o) {
    if (!e || !o) return fail(GYMCUDA_EINVAL, "null argument");
    std::memset(o, 0, sizeof(*o));
    o->obs_dim = e->ki.od; o->act_dim = e->ki.ad; o->act_n = e->ki.actn;
    o->state_dim = e->ki.sd; o->aux_dim = e->ki.aux; o->time_limit = e->limit == 0x7fffffff ? 0 : e->limit;
    const float FMAX = std::numeric_limits<float>::max();
    const float PI_F = 3.1415927410125732f;
    auto set_obs = [&](std::initializer_list<float> hi) { int k = 0; for (float v : hi) { o->obs_high[k] = v; o->obs_low[k] = -v; ++k; } };
    switch (e->cfg.env_kind) {
        case GYMCUDA_CARTPOLE:   // CartPoleEnv.cs:46-48: high = [x_thr*2, float.Max, theta_thr*2, float.Max]
            set_obs({4.800000190734863f, FMAX, 0.41887903213500977f, FMAX});
            break;
        case GYMCUDA_PENDULUM:
            set_obs({1.0f, 1.0f, 8.0f});
            o->act_low[0] = -2.0f; o->act_high[0] = 2.0f;
            break;
        case GYMCUDA_MOUNTAINCAR:
        case GYMCUDA_MOUNTAINCAR_CONT:
            o->obs_low[0] = -1.2f; o->obs_high[0] = 0.6f; o->obs_low[1] = -0.07f; o->obs_high[1] = 0.07f;
            if (e->cfg.env_kind == GYMCUDA_MOUNTAINCAR_CONT) { o->act_low[0] = -1.0f; o->act_high[0] = 1.0f; }
            break;
        case GYMCUDA_ACROBOT:
            set_obs({1.0f, 1.0f, 1.0f, 1.0f, 4.0f * PI_F, 9.0f * PI_F});
            break;
        case GYMCUDA_LUNARLANDER:
        case GYMCUDA_LUNARLANDER_CONT: {   // LunarLanderEnv.cs:412-414
            const float lo[8] = {-1.5f, -1.5f, -5.0f, -5.0f, -PI_F, -5.0f, 0.0f, 0.0f};
            const float hi[8] = {1.5f, 1.5f, 5.0f, 5.0f, PI_F, 5.0f, 1.0f, 1.0f};
            std::memcpy(o->obs_low, lo, sizeof(lo)); std::memcpy(o->obs_high, hi, sizeof(hi));
            if (e->cfg.env_kind == GYMCUDA_LUNARLANDER_CONT) { o->act_low[0] = o->act_low[1] = -1.0f; o->act_high[0] = o->act_high[1] = 1.0f; }
            break;
        }
    }
    return GYMCUDA_OK;
}

// ------------------------------------------------------------------------------------------------
// seeding
// ------------------------------------------------------------------------------------------------
int gymcuda_seed(gymcuda_env* e, uint64_t seed) {
    ENTER(e);
    if (int _c = clock_pull(e)) return _c;
    e->seed = seed;
    CU_TRY(cudaStreamSynchronize(e->stream));
    if (e->d_seeds) { CU_TRY(cudaFree(e->d_seeds)); e->d_seeds = nullptr; }
    // a new generator restarts every stream (CartPoleEnv.cs:197 replaces the RandomState)
    CU_TRY(cudaMemsetAsync(e->d_episode, 0, (size_t)e->n * 4, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    e->t = 0;
    if (int _c = clock_push(e)) return _c;
    return run_ctor(e);
}

int gymcuda_seed_each(gymcuda_env* e, const int32_t* seeds, int n) {
    ENTER(e);
    if (int _c = clock_pull(e)) return _c;
    if (!seeds) return fail(GYMCUDA_EINVAL, "seeds is null");
    // VecEnv.Seed(int[]) throws ArgumentException on a length mismatch (VecEnv.cs:49)
    if (n != e->n) return fail(GYMCUDA_EINVAL, "Number of seeds passed should be equals to number of environments");
    if (!e->d_seeds) CU_TRY(cudaMalloc(&e->d_seeds, (size_t)n * 4));
    CU_TRY(cudaMemcpyAsync(e->d_seeds, seeds, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
    CU_TRY(cudaMemsetAsync(e->d_episode, 0, (size_t)e->n * 4, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    e->t = 0;
    if (int _c = clock_push(e)) return _c;
    return run_ctor(e);
}

// ------------------------------------------------------------------------------------------------
// reset
// ------------------------------------------------------------------------------------------------
static int reset_impl(gymcuda_env* e, const uint8_t* d_mask, float* obs_host) {
    ResetArgs a{};
    a.state = e->d_state; a.aux = e->d_aux; a.prm = e->prm; a.sbd = e->d_sbd; a.ep_t = e->d_ept; a.episode = e->d_episode; a.seeds = e->d_seeds;
    a.mask = d_mask; a.obs = e->d_obs; a.ep_ret = e->d_ep_ret; a.n = e->n; a.env_off = e->cfg.env_id_offset;
    a.seed = e->seed; a.t = e->t;
    CU_TRY(dispatch_reset(e, a));
    e->last_obs = e->d_obs;
    if (obs_host) CU_TRY(cudaMemcpyAsync(obs_host, e->d_obs, e->obs_bytes(), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return GYMCUDA_OK;
}

int gymcuda_reset(gymcuda_env* e, float* obs_out) {
    ENTER(e);
    if (int _c = clock_pull(e)) return _c;
    TRACE("reset");
    int rc = reset_impl(e, nullptr, obs_out);
    if (rc == GYMCUDA_OK) e->has_state = true;
    return rc;
}

int gymcuda_reset_masked(gymcuda_env* e, const uint8_t* mask, float* obs_out) {
    ENTER(e);
    if (int _c = clock_pull(e)) return _c;
    TRACE("reset_masked");
    if (!mask) return fail(GYMCUDA_EINVAL, "mask is null");
    if (!e->has_state) return fail(GYMCUDA_ESTATE, "reset_masked before the first full Reset()");
    CU_TRY(cudaMemcpyAsync(e->d_mask, mask, (size_t)e->n, cudaMemcpyHostToDevice, e->stream));
    return reset_impl(e, e->d_mask, obs_out);
}

// ------------------------------------------------------------------------------------------------
// step
// ------------------------------------------------------------------------------------------------
// ---- device-resident clock (gymcuda_set_device_clock) --------------------------------------------------------------
// In this mode the step kernels read t and the launch sequence number from d_clock and a one-thread kernel advances it after
// every step, so a captured step can be replayed; the host mirrors (e->t, e->seq) are exact only after clock_pull.
static int clock_pull(gymcuda_env* e) {   // device -> host mirrors (synchronises; not legal while the stream is being captured)
    if (!e->device_clock) return GYMCUDA_OK;
    CU_TRY(cudaMemcpyAsync(e->h_small + 6, e->d_clock, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    e->t = e->h_small[6]; e->seq = (uint32_t)e->h_small[7];
    return GYMCUDA_OK;
}
static int clock_push(gymcuda_env* e) {   // host mirrors -> device, after an entry point that advanced them on the host
    if (!e->device_clock) return GYMCUDA_OK;
    e->h_small[6] = e->t; e->h_small[7] = e->seq;
    CU_TRY(cudaMemcpyAsync(e->d_clock, e->h_small + 6, 2 * sizeof(unsigned long long), cudaMemcpyHostToDevice, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return GYMCUDA_OK;
}

// d_actions == nullptr (and no broadcast): the random policy -- ActionSpace.Sample() of step t evaluated in the kernel, also written to sampled_out
static int step_launch(gymcuda_env* e, const void* d_actions, int use_bcast, int32_t bcast, float* d_obs,
                       float* d_reward, uint8_t* d_done, bool gather = false, void* sampled_out = nullptr) {
    if (!e->has_state) return fail(GYMCUDA_ESTATE, "Step() before Reset(): the reference dereferences a null state here (CartPoleEnv.cs:40,141)");
    StepArgs a{};
    a.state = e->d_state; a.aux = e->d_aux; a.prm = e->prm; a.sbd = e->d_sbd; a.ep_t = e->d_ept; a.episode = e->d_episode; a.seeds = e->d_seeds;
    a.actions = d_actions; a.obs = d_obs; a.reward = d_reward; a.done = d_done;
    a.done_idx = e->d_done_idx; a.blk_cnt = e->d_blk_cnt; a.tmp_idx = e->d_tmp_idx; a.done_count = e->d_done_count; a.stats = e->d_stats; a.host_invalid = e->d_invalid_flag; a.ep_ret = e->d_ep_ret; a.sums = e->d_sums; a.done_bits = e->done_bits ? 1 : 0;
    a.n = e->n; a.env_off = e->cfg.env_id_offset; a.seed = e->seed; a.t = e->t; a.limit = e->limit;
    a.use_bcast = use_bcast; a.bcast_action = bcast; a.seq = e->seq; a.fold_prev = e->stats_pending ? 1 : 0;
    if (e->device_clock) { a.clock = e->d_clock; a.fold_prev = 1; }   // (entering the mode zeroed the other counter if nothing was pending)
    a.terminal_obs = e->auto_reset ? e->term_dev : nullptr;
    if (!d_actions && !use_bcast) { a.sample = 1; a.act_out = sampled_out; }
    if (gather) {
        e->g_seq += 1;
        a.world = e->g_world; a.rank = e->g_rank; a.gseq = e->g_seq;
        for (int r = 0; r < e->g_world; ++r) {
            a.peer_obs[r] = reinterpret_cast<float*>(e->g_peer[r]);
            a.peer_flags[r] = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(e->g_peer[r]) + e->g_flags_off);
        }
        a.block_counter = reinterpret_cast<unsigned*>(e->g_local + e->g_counter_off);
    }
    CU_TRY(dispatch_step(e, a));
    // which epilogue ran: the warp-granular one of the partitioned LunarLander step writes d_done_idx itself
    e->done_list_stale = !(is_lunar(e) && a.world == 0);
    if (e->device_clock) { clock_tick_kernel<<<1, 1, 0, e->stream>>>(e->d_clock, 1ull); CU_TRY(cudaGetLastError()); }
    e->stats_pending = true;   // this launch's episode count sits in done_count[seq & 1] until the next launch folds it into stats[0]
    e->t += 1;
    e->seq += 1;
    e->env_steps += (unsigned long long)e->n;
    e->last_obs = d_obs;
    return GYMCUDA_OK;
}

static bool is_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

// What the kernels' vector accesses need: observations leave as 16 B (float4) or 8 B (float2) vectors, a 2-D Box action
// is read as one float2.  (A GCHandle-pinned managed float[] is only 8-byte aligned.)
constexpr size_t OBS_ALIGN = 16;
static size_t act_align(const gymcuda_env* e) { return (size_t)e->ki.ad * 4; }

static int check_device_buffers(const gymcuda_env* e, const void* actions, const float* obs, const float* reward) {
    if (actions && !is_aligned(actions, act_align(e))) return fail(GYMCUDA_EINVAL, "device action buffer must be %zu-byte aligned", act_align(e));
    if (obs && !is_aligned(obs, OBS_ALIGN)) return fail(GYMCUDA_EINVAL, "device observation buffer must be 16-byte aligned (it is written with vector stores)");
    if (reward && !is_aligned(reward, 4)) return fail(GYMCUDA_EINVAL, "device reward buffer must be 4-byte aligned");
    return GYMCUDA_OK;
}

// Device-visible alias of a host pointer if (and only if) it is page-locked memory mapped into the device address space
// (cudaHostAlloc / cudaHostRegister / gymcuda_host_alloc), else null.
// `align`: a mapped buffer the kernel could not address with its vector accesses counts as not mapped (it is staged).
// Asked anew on every call (a fraction of a microsecond next to a launch): a cached answer would outlive a buffer that was
// unpinned or freed behind the library's back (raw cudaFreeHost / cudaHostUnregister, a torch tensor going away) and whose
// address came back as pageable memory -- the kernel would then store through a stale device alias (ADVICE, round 1).
static void* mapped_alias(gymcuda_env*, const void* host, int, size_t align) {
    if (!host || !is_aligned(host, align)) return nullptr;
    cudaPointerAttributes at;
    void* dev = nullptr;
    if (cudaPointerGetAttributes(&at, host) == cudaSuccess && at.type == cudaMemoryTypeHost) dev = at.devicePointer;
    else cudaGetLastError();   // pageable memory reports an error on old drivers: clear it
    if (dev && !is_aligned(dev, align)) dev = nullptr;
    return dev;
}

// Before a host-buffer step: rejected actions of earlier asynchronous *_device steps belong to THOSE calls (reported by
// gymcuda_sync); if the caller never asked, they are dropped here rather than blamed on the step that follows.
static int drain_async_invalid(gymcuda_env* e) {
    if (!e->async_steps) return GYMCUDA_OK;
    CU_TRY(cudaStreamSynchronize(e->stream));
    e->async_steps = false;
    if (*e->h_invalid) {
        *e->h_invalid = 0;
        CU_TRY(cudaMemcpyAsync(e->h_small, e->d_stats, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaStreamSynchronize(e->stream));
        e->invalid_seen = e->h_small[1];
    }
    return GYMCUDA_OK;
}

static int step_finish_host(gymcuda_env* e, float* obs, float* reward, uint8_t* done) {
    const size_t ob = e->obs_bytes(), n = (size_t)e->n;
    const bool adjacent = obs && reward && done && reinterpret_cast<uint8_t*>(obs) + ob == reinterpret_cast<uint8_t*>(reward) &&
                          reinterpret_cast<uint8_t*>(reward) + n * 4 == done;
    if (adjacent) {
        CU_TRY(cudaMemcpyAsync(obs, e->d_out, ob + n * 4 + n, cudaMemcpyDeviceToHost, e->stream));
    } else {
        if (obs) CU_TRY(cudaMemcpyAsync(obs, e->d_obs, ob, cudaMemcpyDeviceToHost, e->stream));
        if (reward) CU_TRY(cudaMemcpyAsync(reward, e->d_reward, n * 4, cudaMemcpyDeviceToHost, e->stream));
        if (done) CU_TRY(cudaMemcpyAsync(done, e->d_done, n, cudaMemcpyDeviceToHost, e->stream));
    }
    if (e->term_host) CU_TRY(cudaMemcpyAsync(e->term_host, e->d_term_own, e->obs_bytes(), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    if (*e->h_invalid) {   // written through mapped memory by the kernel: no extra copy on the common path
        *e->h_invalid = 0;
        CU_TRY(cudaMemcpyAsync(e->h_small, e->d_stats, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaStreamSynchronize(e->stream));
        unsigned long long bad = e->h_small[1] - e->invalid_seen;
        e->invalid_seen = e->h_small[1];
        return fail(GYMCUDA_EACTION, "%llu action(s) invalid for this action space; those envs were not stepped", bad);
    }
    return GYMCUDA_OK;
}

int gymcuda_step(gymcuda_env* e, const void* actions, float* obs, float* reward, uint8_t* done) {
    ENTER(e);
    TRACE("step");
    if (!actions) return fail(GYMCUDA_EINVAL, "actions is null");
    if (!e->has_state) return fail(GYMCUDA_ESTATE, "Step() before Reset(): the reference dereferences a null state here (CartPoleEnv.cs:40,141)");
    if (int rc = drain_async_invalid(e)) return rc;
    // Zero-copy path: when all four host buffers are page-locked, the kernel reads the actions and
    // writes obs / reward / done straight through their device aliases -- the H2D and D2H traffic
    // crosses PCIe inside the step kernel, overlapped with the math, with no DMA set-up per buffer.
    void* za = mapped_alias(e, actions, 0, act_align(e));
    void* zo = mapped_alias(e, obs, 1, OBS_ALIGN);
    void* zr = mapped_alias(e, reward, 2, 4);
    void* zd = mapped_alias(e, done, 3, 1);
    if (za && zo && zr && zd) {
        int rc = step_launch(e, za, 0, 0, (float*)zo, (float*)zr, (uint8_t*)zd);
        if (rc) return rc;
        e->last_obs = nullptr;   // the observations were written to host memory only
        if (e->term_host) CU_TRY(cudaMemcpyAsync(e->term_host, e->d_term_own, e->obs_bytes(), cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaStreamSynchronize(e->stream));
        if (*e->h_invalid) return step_finish_host(e, nullptr, nullptr, nullptr);
        return GYMCUDA_OK;
    }
    // mixed / pageable buffers: every page-locked one is still accessed in place, the others are staged
    if (!za) CU_TRY(cudaMemcpyAsync(e->d_actions, actions, e->act_bytes(), cudaMemcpyHostToDevice, e->stream));
    const bool any_mapped_out = zo || zr || zd;
    int rc = step_launch(e, za ? za : e->d_actions, 0, 0, zo ? (float*)zo : e->d_obs, zr ? (float*)zr : e->d_reward,
                         zd ? (uint8_t*)zd : e->d_done);
    if (rc) return rc;
    if (zo) e->last_obs = nullptr;
    if (!any_mapped_out) return step_finish_host(e, obs, reward, done);   // one DMA when the three are adjacent
    return step_finish_host(e, zo ? nullptr : obs, zr ? nullptr : reward, zd ? nullptr : done);
}

int gymcuda_step_broadcast(gymcuda_env* e, int32_t action, float* obs, float* reward, uint8_t* done) {
    ENTER(e);
    TRACE("step_broadcast");
    if (e->ki.actn == 0) return fail(GYMCUDA_EINVAL, "IVecEnv.Step(int action) needs a Discrete action space");
    if (int rc0 = drain_async_invalid(e)) return rc0;
    int rc = step_launch(e, nullptr, 1, action, e->d_obs, e->d_reward, e->d_done);
    if (rc) return rc;
    return step_finish_host(e, obs, reward, done);
}

int gymcuda_step_device(gymcuda_env* e, const void* d_actions, float* d_obs, float* d_reward, uint8_t* d_done) {
    ENTER(e);
    TRACE("step_device");
    if (!d_actions) return fail(GYMCUDA_EINVAL, "d_actions is null");
    const bool no_obs = d_obs == GYMCUDA_NO_OBS;   // no observation copy: the caller reads the state in place (gymcuda_obs_view_device) or asks gymcuda_observe
    if (no_obs) d_obs = nullptr;
    if (int rc = check_device_buffers(e, d_actions, d_obs, d_reward)) return rc;
    e->async_steps = true;
    return step_launch(e, d_actions, 0, 0, no_obs ? nullptr : (d_obs ? d_obs : e->d_obs), d_reward ? d_reward : e->d_reward,
                       d_done ? d_done : e->d_done);
}

int gymcuda_obs_view_device(gymcuda_env* e, const float** d_obs) {
    ENTER(e);
    if (!d_obs) return fail(GYMCUDA_EINVAL, "d_obs is null");
    const int k = e->cfg.env_kind;
    if (k != GYMCUDA_CARTPOLE && k != GYMCUDA_MOUNTAINCAR && k != GYMCUDA_MOUNTAINCAR_CONT)
        return fail(GYMCUDA_EINVAL, "the observation of this env kind is computed from its state: there is nothing to view in place");
    *d_obs = reinterpret_cast<const float*>(e->d_state);
    return GYMCUDA_OK;
}

int gymcuda_set_terminal_obs(gymcuda_env* e, float* buffer) {
    ENTER(e);
    CU_TRY(cudaStreamSynchronize(e->stream));   // steps in flight still write the previous buffer
    e->term_dev = nullptr; e->term_host = nullptr;
    if (!buffer) return GYMCUDA_OK;
    if (!e->auto_reset) return fail(GYMCUDA_EINVAL, "terminal observations exist only under GYMCUDA_FLAG_AUTO_RESET (without it Step returns the terminal observation itself)");
    cudaPointerAttributes at;
    cudaError_t ce = cudaPointerGetAttributes(&at, buffer);
    if (ce != cudaSuccess) { cudaGetLastError(); at.type = cudaMemoryTypeUnregistered; at.devicePointer = nullptr; }
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) {
        if (!is_aligned(buffer, OBS_ALIGN)) return fail(GYMCUDA_EINVAL, "device terminal-observation buffer must be 16-byte aligned");
        e->term_dev = buffer;
    } else if (at.type == cudaMemoryTypeHost && at.devicePointer && is_aligned(at.devicePointer, OBS_ALIGN)) {
        e->term_dev = static_cast<float*>(at.devicePointer);   // page-locked and mapped: written in place over PCIe
    } else {   // pageable (or unaligned page-locked) host memory: staged
        if (!e->d_term_own) CU_TRY(cudaMalloc(&e->d_term_own, e->obs_bytes()));
        // the staging copy starts as the caller's buffer: rows of envs that have not finished an episode keep their contents
        CU_TRY(cudaMemcpyAsync(e->d_term_own, buffer, e->obs_bytes(), cudaMemcpyHostToDevice, e->stream));
        CU_TRY(cudaStreamSynchronize(e->stream));
        e->term_dev = e->d_term_own;
        e->term_host = buffer;
    }
    return GYMCUDA_OK;
}

// k steps with caller-supplied actions: the rollout kernel's generic variant reading `actions_in` (classic envs), k step
// launches for LunarLander
int gymcuda_step_many_device(gymcuda_env* e, int k_steps, const void* d_actions, float* d_obs, float* d_reward, uint8_t* d_done) {
    ENTER(e);
    if (int _c = clock_pull(e)) return _c;
    TRACE("step_many_device");
    if (k_steps <= 0) return fail(GYMCUDA_EINVAL, "k_steps must be > 0");
    if (!d_actions) return fail(GYMCUDA_EINVAL, "d_actions is null");
    if (!e->has_state) return fail(GYMCUDA_ESTATE, "Step() before Reset(): the reference dereferences a null state here (CartPoleEnv.cs:40,141)");
    if (int rc = check_device_buffers(e, d_actions, d_obs, d_reward)) return rc;
    e->async_steps = true;
    const size_t n = (size_t)e->n;
    if (is_lunar(e)) {
        for (int j = 0; j < k_steps; ++j) {
            int rc = step_launch(e, reinterpret_cast<const uint8_t*>(d_actions) + (size_t)j * n * e->ki.ad * 4, 0, 0,
                                 d_obs ? d_obs + (size_t)j * n * e->ki.od : e->d_obs, d_reward ? d_reward + (size_t)j * n : e->d_reward,
                                 d_done ? d_done + (size_t)j * n : e->d_done);
            if (rc) return rc;
        }
        return GYMCUDA_OK;
    }
    RolloutArgs a{};
    a.state = e->d_state; a.aux = e->d_aux; a.prm = e->prm; a.sbd = e->d_sbd; a.ep_t = e->d_ept; a.episode = e->d_episode; a.seeds = e->d_seeds;
    a.obs = d_obs; a.reward = d_reward; a.done = d_done; a.actions = nullptr; a.actions_in = d_actions; a.host_invalid = e->d_invalid_flag;
    a.stats = e->d_stats; a.ep_ret = e->d_ep_ret; a.sums = e->d_sums; a.done_bits = e->done_bits ? 1 : 0;
    a.n = e->n; a.k_steps = k_steps; a.env_off = e->cfg.env_id_offset; a.seed = e->seed; a.t = e->t; a.limit = e->limit;
    CU_TRY(dispatch_rollout(e, a));
    e->t += (uint64_t)k_steps;
    e->env_steps += (unsigned long long)e->n * (unsigned long long)k_steps;
    e->last_obs = nullptr;
    return clock_push(e);
}

static cudaError_t scratch_reserve(gymcuda_env* e, int slot, size_t bytes, void** out);

// Host-buffer k-step calls as a three-stage pipeline over chunks of steps: the actions of chunk c + 1 travel to the device
// (in_stream) and the trajectory of chunk c - 1 travels back (out_stream) while chunk c is stepped on the handle's stream.
// PCIe is full duplex, so the call lasts about as long as its larger direction -- the trajectory -- alone.  The state is
// carried between the chunk launches in device memory, so the result is that of one k-step launch (tests T4: k-split).
// Chunk = a multiple of 8 steps (the chunked kernels' alignment) holding ~16 MB of trajectory.
static int pipe_chunk_steps(const gymcuda_env* e, int k_steps) {
    const size_t per_step = (size_t)e->n * ((size_t)e->ki.od * 4 + 4 + 1 + (size_t)e->ki.ad * 4);
    size_t c = ((size_t)16 << 20) / (per_step ? per_step : 1);
    c = (c + 7) / 8 * 8;
    if (c < 8) c = 8;
    // the first chunk ends on a multiple of 8 of the absolute step index, so that every later launch starts aligned
    const size_t head = (size_t)((8u - ((unsigned)e->t & 7u)) & 7u);
    size_t first = head ? head + (c > 8 ? c - 8 : 0) : c;
    if (first > (size_t)k_steps) first = (size_t)k_steps;
    return (int)first;   // (the caller uses c for the following chunks)
}
static cudaError_t pipe_prepare(gymcuda_env* e, size_t events) {
    cudaError_t ce;
    if (!e->in_stream && (ce = cudaStreamCreateWithFlags(&e->in_stream, cudaStreamNonBlocking)) != cudaSuccess) return ce;
    if (!e->out_stream && (ce = cudaStreamCreateWithFlags(&e->out_stream, cudaStreamNonBlocking)) != cudaSuccess) return ce;
    while (e->pipe_events.size() < events) {
        cudaEvent_t ev;
        if ((ce = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return ce;
        e->pipe_events.push_back(ev);
    }
    return cudaSuccess;
}

// supplied = true: gymcuda_step_many (actions in, obs / reward / done out); false: gymcuda_rollout_random (everything out)
static int host_k_steps(gymcuda_env* e, bool supplied, int k_steps, const void* actions_in, float* obs, float* reward, uint8_t* done, void* actions_out) {
    const size_t n = (size_t)e->n, od = (size_t)e->ki.od, ab = (size_t)e->ki.ad * 4;
    const size_t kn = (size_t)k_steps * n;
    void *d_obs = nullptr, *d_reward = nullptr, *d_done = nullptr, *d_act = nullptr;
#define HK_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cudaGetLastError(); return fail(_e == cudaErrorMemoryAllocation ? GYMCUDA_ENOMEM : GYMCUDA_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); } } while (0)
    HK_TRY(scratch_reserve(e, 0, obs ? kn * od * 4 : 0, &d_obs));
    HK_TRY(scratch_reserve(e, 1, reward ? kn * 4 : 0, &d_reward));
    HK_TRY(scratch_reserve(e, 2, done ? kn : 0, &d_done));
    HK_TRY(scratch_reserve(e, 3, (supplied || actions_out) ? kn * ab : 0, &d_act));
    const int c_first = pipe_chunk_steps(e, k_steps);
    int c_rest = (int)((((size_t)16 << 20) / (n * (od * 4 + 4 + 1 + ab)) + 7) / 8 * 8);
    if (c_rest < 8) c_rest = 8;
    const size_t chunks = 1 + (size_t)((k_steps - c_first) + c_rest - 1) / (size_t)c_rest;
    HK_TRY(pipe_prepare(e, 2 * chunks + 1));
    // everything queued on the handle's stream so far (earlier asynchronous steps, a pending H2D) precedes the pipeline
    HK_TRY(cudaEventRecord(e->pipe_events[2 * chunks], e->stream));
    HK_TRY(cudaStreamWaitEvent(e->in_stream, e->pipe_events[2 * chunks], 0));
    HK_TRY(cudaStreamWaitEvent(e->out_stream, e->pipe_events[2 * chunks], 0));
    size_t ci = 0;
    for (int k0 = 0; k0 < k_steps; ++ci) {
        const int kc = k0 == 0 ? c_first : (k_steps - k0 < c_rest ? k_steps - k0 : c_rest);
        const size_t off = (size_t)k0 * n, cnt = (size_t)kc * n;
        uint8_t* const d_act_c = d_act ? static_cast<uint8_t*>(d_act) + off * ab : nullptr;
        if (supplied) {
            HK_TRY(cudaMemcpyAsync(d_act_c, static_cast<const uint8_t*>(actions_in) + off * ab, cnt * ab, cudaMemcpyHostToDevice, e->in_stream));
            HK_TRY(cudaEventRecord(e->pipe_events[2 * ci], e->in_stream));
            HK_TRY(cudaStreamWaitEvent(e->stream, e->pipe_events[2 * ci], 0));
        }
        float* const d_obs_c = d_obs ? static_cast<float*>(d_obs) + off * od : nullptr;
        float* const d_rew_c = d_reward ? static_cast<float*>(d_reward) + off : nullptr;
        uint8_t* const d_done_c = d_done ? static_cast<uint8_t*>(d_done) + off : nullptr;
        const int rc = supplied ? gymcuda_step_many_device(e, kc, d_act_c, d_obs_c, d_rew_c, d_done_c)
                                : gymcuda_rollout_random_device(e, kc, d_obs_c, d_rew_c, d_done_c, d_act_c);
        if (rc) { cudaStreamSynchronize(e->in_stream); cudaStreamSynchronize(e->out_stream); cudaStreamSynchronize(e->stream); return rc; }
        HK_TRY(cudaEventRecord(e->pipe_events[2 * ci + 1], e->stream));
        HK_TRY(cudaStreamWaitEvent(e->out_stream, e->pipe_events[2 * ci + 1], 0));
        if (obs) HK_TRY(cudaMemcpyAsync(obs + off * od, d_obs_c, cnt * od * 4, cudaMemcpyDeviceToHost, e->out_stream));
        if (reward) HK_TRY(cudaMemcpyAsync(reward + off, d_rew_c, cnt * 4, cudaMemcpyDeviceToHost, e->out_stream));
        if (done) HK_TRY(cudaMemcpyAsync(done + off, d_done_c, cnt, cudaMemcpyDeviceToHost, e->out_stream));
        if (!supplied && actions_out) HK_TRY(cudaMemcpyAsync(static_cast<uint8_t*>(actions_out) + off * ab, d_act_c, cnt * ab, cudaMemcpyDeviceToHost, e->out_stream));
        k0 += kc;
    }
    HK_TRY(cudaStreamSynchronize(e->out_stream));
#undef HK_TRY
    e->async_steps = false;   // synchronised below
    return step_finish_host(e, nullptr, nullptr, nullptr);   // synchronise the handle's stream + report rejected actions
}

int gymcuda_step_many(gymcuda_env* e, int k_steps, const void* actions, float* obs, float* reward, uint8_t* done) {
    ENTER(e);
    TRACE("step_many");
    if (k_steps <= 0) return fail(GYMCUDA_EINVAL, "k_steps must be > 0");
    if (!actions) return fail(GYMCUDA_EINVAL, "actions is null");
    if (!e->has_state) return fail(GYMCUDA_ESTATE, "Step() before Reset(): the reference dereferences a null state here (CartPoleEnv.cs:40,141)");
    if (int rc0 = drain_async_invalid(e)) return rc0;
    return host_k_steps(e, true, k_steps, actions, obs, reward, done, nullptr);
}

// ------------------------------------------------------------------------------------------------
// rollout
// ------------------------------------------------------------------------------------------------
int gymcuda_rollout_random_device(gymcuda_env* e, int k_steps, float* d_obs, float* d_reward, uint8_t* d_done, void* d_actions) {
    ENTER(e);
    if (int _c = clock_pull(e)) return _c;
    TRACE("rollout_random_device");
    if (k_steps <= 0) return fail(GYMCUDA_EINVAL, "k_steps must be > 0");
    if (!e->has_state) return fail(GYMCUDA_ESTATE, "rollout before Reset()");
    if (int rc = check_device_buffers(e, d_actions, d_obs, d_reward)) return rc;
    if (is_lunar(e)) {   // k step launches with the policy sampled in the kernel (lunar_kernels.cu)
        const size_t n = (size_t)e->n;
        for (int j = 0; j < k_steps; ++j) {
            int rc = step_launch(e, nullptr, 0, 0, d_obs ? d_obs + (size_t)j * n * e->ki.od : e->d_obs, d_reward ? d_reward + (size_t)j * n : e->d_reward,
                                 d_done ? d_done + (size_t)j * n : e->d_done, false,
                                 d_actions ? reinterpret_cast<uint8_t*>(d_actions) + (size_t)j * n * e->ki.ad * 4 : nullptr);
            if (rc) return rc;
        }
        e->last_obs = nullptr;
        return GYMCUDA_OK;
    }
    RolloutArgs a{};
    a.state = e->d_state; a.aux = e->d_aux; a.prm = e->prm; a.sbd = e->d_sbd; a.ep_t = e->d_ept; a.episode = e->d_episode; a.seeds = e->d_seeds;
    a.obs = d_obs; a.reward = d_reward; a.done = d_done; a.actions = d_actions; a.stats = e->d_stats; a.ep_ret = e->d_ep_ret; a.sums = e->d_sums; a.done_bits = e->done_bits ? 1 : 0;
    a.n = e->n; a.k_steps = k_steps; a.env_off = e->cfg.env_id_offset; a.seed = e->seed; a.t = e->t; a.limit = e->limit;
    CU_TRY(dispatch_rollout(e, a));
    e->t += (uint64_t)k_steps;
    e->env_steps += (unsigned long long)e->n * (unsigned long long)k_steps;
    e->last_obs = nullptr;   // the current observations are not materialised; gymcuda_observe recomputes them
    return clock_push(e);
}

// grow-only device scratch for the host-buffer rollout (a cudaMalloc + cudaFree pair per call costs more than a
// short rollout itself)
static cudaError_t scratch_reserve(gymcuda_env* e, int slot, size_t bytes, void** out) {
    *out = nullptr;
    if (bytes == 0) return cudaSuccess;
    if (e->scratch_cap[slot] < bytes) {
        cudaFree(e->scratch[slot]);
        e->scratch[slot] = nullptr; e->scratch_cap[slot] = 0;
        cudaError_t ce = cudaMalloc(&e->scratch[slot], bytes);
        if (ce != cudaSuccess) return ce;
        e->scratch_cap[slot] = bytes;
    }
    *out = e->scratch[slot];
    return cudaSuccess;
}

int gymcuda_rollout_random(gymcuda_env* e, int k_steps, float* obs, float* reward, uint8_t* done, void* actions) {
    ENTER(e);
    TRACE("rollout_random");
    if (k_steps <= 0) return fail(GYMCUDA_EINVAL, "k_steps must be > 0");
    if (!e->has_state) return fail(GYMCUDA_ESTATE, "rollout before Reset()");
    return host_k_steps(e, false, k_steps, nullptr, obs, reward, done, actions);
}

// ------------------------------------------------------------------------------------------------
// ActionSpace.Sample on device
// ------------------------------------------------------------------------------------------------
int gymcuda_sample_actions_device(gymcuda_env* e, const uint8_t* d_mask, void* d_actions_out) {
    ENTER(e);
    if (int _c = clock_pull(e)) return _c;
    if (!d_actions_out) return fail(GYMCUDA_EINVAL, "d_actions_out is null");
    if (int rc = check_device_buffers(e, d_actions_out, nullptr, nullptr)) return rc;
    if (d_mask && e->ki.actn == 0) return fail(GYMCUDA_EINVAL, "Box.sample cannot be provided a mask.");   // Box.cs:70-73
    SampleArgs a{};
    a.seeds = e->d_seeds; a.mask = d_mask; a.out = d_actions_out; a.n = e->n; a.env_off = e->cfg.env_id_offset;
    a.seed = e->seed; a.t = e->t;
    CU_TRY(dispatch_sample(e, a));
    return GYMCUDA_OK;
}

int gymcuda_sample_actions(gymcuda_env* e, const uint8_t* mask, void* actions_out) {
    ENTER(e);
    if (!actions_out) return fail(GYMCUDA_EINVAL, "actions_out is null");
    if (mask && e->ki.actn == 0) return fail(GYMCUDA_EINVAL, "Box.sample cannot be provided a mask.");
    uint8_t* d_mask = nullptr;
    if (mask) {   // [n][act_n] bytes, kept for the life of the handle (a policy that masks does so every step)
        if (!e->d_sample_mask) CU_TRY(cudaMalloc(&e->d_sample_mask, (size_t)e->n * e->ki.actn));
        d_mask = e->d_sample_mask;
        CU_TRY(cudaMemcpyAsync(d_mask, mask, (size_t)e->n * e->ki.actn, cudaMemcpyHostToDevice, e->stream));
    }
    int rc = gymcuda_sample_actions_device(e, d_mask, e->d_actions);
    if (rc == GYMCUDA_OK) {
        cudaError_t ce = cudaMemcpyAsync(actions_out, e->d_actions, e->act_bytes(), cudaMemcpyDeviceToHost, e->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
        if (ce != cudaSuccess) rc = fail(GYMCUDA_ECUDA, "sample copy failed: %s", cudaGetErrorString(ce));
    }
    return rc;
}

// Box.Sample() of an arbitrary box (box_sample.cuh)
int gymcuda_box_sample_device(int device, void* cuda_stream, uint64_t seed, uint64_t index, const float* d_low, const float* d_high,
                              int dim, int count, int as_int, float* d_out) {
    if (!d_low || !d_high || !d_out) return fail(GYMCUDA_EINVAL, "low / high / out is null");
    if (dim <= 0 || count <= 0) return fail(GYMCUDA_EINVAL, "dim and count must be > 0");
    CU_TRY(cudaSetDevice(device));
    BoxSampleArgs a{d_low, d_high, d_out, dim, (long long)dim * (long long)count, as_int, seed, index};
    const long long blocks = (a.total + 255) / 256;
    if (blocks > 0x7fffffffll) return fail(GYMCUDA_EINVAL, "count * dim too large");
    box_sample_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(a);
    CU_TRY(cudaGetLastError());
    return GYMCUDA_OK;
}

int gymcuda_box_sample(int device, uint64_t seed, uint64_t index, const float* low, const float* high, int dim, int count,
                       int as_int, float* out) {
    if (!low || !high || !out) return fail(GYMCUDA_EINVAL, "low / high / out is null");
    if (dim <= 0 || count <= 0) return fail(GYMCUDA_EINVAL, "dim and count must be > 0");
    CU_TRY(cudaSetDevice(device));
    float* d = nullptr;
    const size_t total = (size_t)dim * (size_t)count;
    CU_TRY(cudaMalloc(&d, (total + 2 * (size_t)dim) * 4));
    float *d_low = d + total, *d_high = d_low + dim;
    cudaError_t ce = cudaMemcpy(d_low, low, (size_t)dim * 4, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(d_high, high, (size_t)dim * 4, cudaMemcpyHostToDevice);
    int rc = GYMCUDA_OK;
    if (ce == cudaSuccess) rc = gymcuda_box_sample_device(device, nullptr, seed, index, d_low, d_high, dim, count, as_int, d);
    if (ce == cudaSuccess && rc == GYMCUDA_OK) ce = cudaMemcpy(out, d, total * 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (ce != cudaSuccess) return fail(GYMCUDA_ECUDA, "box sample failed: %s", cudaGetErrorString(ce));
    return rc;
}

// ------------------------------------------------------------------------------------------------
// done compaction
// ------------------------------------------------------------------------------------------------
// builds d_done_idx from the per-CTA sub-lists of the last step launch (once per launch)
static int done_list_build(gymcuda_env* e) {
    // (under the device-resident clock a graph replay may have stepped since the last build without the host knowing: rebuild.
    // The scan works on a copy, so building twice from the same launch gives the same list.)
    if (!e->done_list_stale && !e->device_clock) return GYMCUDA_OK;
    if (is_lunar(e) && !e->done_list_stale) return GYMCUDA_OK;   // partitioned LunarLander step: the kernel wrote the list itself
    const int nb = (e->n + STEP_BLOCK - 1) / STEP_BLOCK;
    CU_TRY(cudaMemcpyAsync(e->d_blk_off, e->d_blk_cnt, (size_t)(nb + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, e->stream));
    partition_scan_kernel<<<1, 1024, 0, e->stream>>>(e->d_blk_off, nb);
    done_list_scatter_kernel<<<(nb + 7) / 8, 256, 0, e->stream>>>(e->d_blk_off, e->d_tmp_idx, nb, STEP_BLOCK, e->d_done_idx);
    CU_TRY(cudaGetLastError());
    e->done_list_stale = false;
    return GYMCUDA_OK;
}

int gymcuda_done_indices(gymcuda_env* e, int32_t* idx, int32_t* count) {
    ENTER(e);
    if (int _c = clock_pull(e)) return _c;
    if (!count) return fail(GYMCUDA_EINVAL, "count is null");
    if (e->seq == 0) { *count = 0; return GYMCUDA_OK; }
    if (idx) { if (int rc = done_list_build(e)) return rc; }
    int32_t* h = reinterpret_cast<int32_t*>(e->h_small + 2);
    CU_TRY(cudaMemcpyAsync(h, e->d_done_count + ((e->seq - 1) & 1), sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    *count = *h;
    if (idx && *count > 0) {
        CU_TRY(cudaMemcpyAsync(idx, e->d_done_idx, (size_t)*count * 4, cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaStreamSynchronize(e->stream));
    }
    return GYMCUDA_OK;
}

int gymcuda_done_indices_device(gymcuda_env* e, const int32_t** d_idx, const int32_t** d_count) {
    ENTER(e);
    if (int _c = clock_pull(e)) return _c;
    if (d_idx) { if (int rc = done_list_build(e)) return rc; *d_idx = e->d_done_idx; }
    if (d_count) *d_count = e->seq == 0 ? e->d_done_count : e->d_done_count + ((e->seq - 1) & 1);
    return GYMCUDA_OK;
}

// ------------------------------------------------------------------------------------------------
// snapshot / teacher forcing
// ------------------------------------------------------------------------------------------------
int gymcuda_get_state(gymcuda_env* e, float* state, int32_t* aux, uint64_t* t) {
    ENTER(e);
    if (int _c = clock_pull(e)) return _c;
    if (e->auxw > 0) {   // LunarLander: device arrays are field-major [word][env]; the ABI is [env][word]
        const size_t n = (size_t)e->n; const int sd = e->ki.sd, ad = e->ki.aux, aw = e->auxw;
        std::vector<float> fs(n * sd); std::vector<int32_t> fa(n * aw), ept(n), epi(n);
        CU_TRY(cudaMemcpyAsync(fs.data(), e->d_state, n * sd * 4, cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaMemcpyAsync(fa.data(), e->d_aux, n * aw * 4, cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaMemcpyAsync(ept.data(), e->d_ept, n * 4, cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaMemcpyAsync(epi.data(), e->d_episode, n * 4, cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaStreamSynchronize(e->stream));
        for (size_t i = 0; i < n; ++i) {
            if (state) for (int f = 0; f < sd; ++f) state[i * sd + f] = fs[(size_t)f * n + i];
            if (aux) { for (int f = 0; f < aw; ++f) aux[i * ad + f] = fa[(size_t)f * n + i]; aux[i * ad + aw] = ept[i]; aux[i * ad + aw + 1] = epi[i]; }
        }
        if (t) *t = e->t;
        return GYMCUDA_OK;
    }
    if (state) CU_TRY(cudaMemcpyAsync(state, e->d_state, (size_t)e->n * e->ki.sd * 4, cudaMemcpyDeviceToHost, e->stream));
    std::vector<int32_t> sbd, ept, epi;
    if (aux) {
        sbd.resize(e->n); ept.resize(e->n); epi.resize(e->n);
        CU_TRY(cudaMemcpyAsync(sbd.data(), e->d_sbd, (size_t)e->n * 4, cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaMemcpyAsync(ept.data(), e->d_ept, (size_t)e->n * 4, cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(cudaMemcpyAsync(epi.data(), e->d_episode, (size_t)e->n * 4, cudaMemcpyDeviceToHost, e->stream));
    }
    CU_TRY(cudaStreamSynchronize(e->stream));
    if (aux) for (int i = 0; i < e->n; ++i) { aux[3 * i] = sbd[i]; aux[3 * i + 1] = ept[i]; aux[3 * i + 2] = epi[i]; }
    if (t) *t = e->t;
    return GYMCUDA_OK;
}

int gymcuda_set_state(gymcuda_env* e, const float* state, const int32_t* aux, uint64_t t) {
    ENTER(e);
    if (int _c = clock_pull(e)) return _c;
    if (e->auxw > 0) {
        if (!state || !aux) return fail(GYMCUDA_EINVAL, "LunarLander set_state needs both state and aux");
        const size_t n = (size_t)e->n; const int sd = e->ki.sd, ad = e->ki.aux, aw = e->auxw;
        std::vector<float> fs(n * sd); std::vector<int32_t> fa(n * aw), ept(n), epi(n);
        for (size_t i = 0; i < n; ++i) {
            for (int f = 0; f < sd; ++f) fs[(size_t)f * n + i] = state[i * sd + f];
            for (int f = 0; f < aw; ++f) fa[(size_t)f * n + i] = aux[i * ad + f];
            ept[i] = aux[i * ad + aw]; epi[i] = aux[i * ad + aw + 1];
        }
        CU_TRY(cudaMemcpyAsync(e->d_state, fs.data(), n * sd * 4, cudaMemcpyHostToDevice, e->stream));
        CU_TRY(cudaMemcpyAsync(e->d_aux, fa.data(), n * aw * 4, cudaMemcpyHostToDevice, e->stream));
        CU_TRY(cudaMemcpyAsync(e->d_ept, ept.data(), n * 4, cudaMemcpyHostToDevice, e->stream));
        CU_TRY(cudaMemcpyAsync(e->d_episode, epi.data(), n * 4, cudaMemcpyHostToDevice, e->stream));
        CU_TRY(cudaStreamSynchronize(e->stream));
        e->t = t; e->has_state = true; e->last_obs = nullptr;
        return clock_push(e);
    }
    if (state) CU_TRY(cudaMemcpyAsync(e->d_state, state, (size_t)e->n * e->ki.sd * 4, cudaMemcpyHostToDevice, e->stream));
    std::vector<int32_t> sbd, ept, epi;
    if (aux) {
        sbd.resize(e->n); ept.resize(e->n); epi.resize(e->n);
        for (int i = 0; i < e->n; ++i) { sbd[i] = aux[3 * i]; ept[i] = aux[3 * i + 1]; epi[i] = aux[3 * i + 2]; }
        CU_TRY(cudaMemcpyAsync(e->d_sbd, sbd.data(), (size_t)e->n * 4, cudaMemcpyHostToDevice, e->stream));
        CU_TRY(cudaMemcpyAsync(e->d_ept, ept.data(), (size_t)e->n * 4, cudaMemcpyHostToDevice, e->stream));
        CU_TRY(cudaMemcpyAsync(e->d_episode, epi.data(), (size_t)e->n * 4, cudaMemcpyHostToDevice, e->stream));
    }
    CU_TRY(cudaStreamSynchronize(e->stream));
    e->t = t;
    e->has_state = true;
    e->last_obs = nullptr;
    return clock_push(e);
}

static int observe_device(gymcuda_env* e) {
    // a masked reset with an all-zero mask only recomputes observations from the stored state
    ResetArgs a{};
    a.state = e->d_state; a.aux = e->d_aux; a.prm = e->prm; a.sbd = e->d_sbd; a.ep_t = e->d_ept; a.episode = e->d_episode; a.seeds = e->d_seeds;
    CU_TRY(cudaMemsetAsync(e->d_mask, 0, (size_t)e->n, e->stream));
    a.mask = e->d_mask; a.obs = e->d_obs; a.n = e->n; a.env_off = e->cfg.env_id_offset; a.seed = e->seed; a.t = e->t;
    CU_TRY(dispatch_reset(e, a));
    e->last_obs = e->d_obs;
    return GYMCUDA_OK;
}

int gymcuda_observe(gymcuda_env* e, float* obs) {
    ENTER(e);
    if (!obs) return fail(GYMCUDA_EINVAL, "obs is null");
    if (!e->has_state) return fail(GYMCUDA_ESTATE, "observe before Reset()");
    int rc = observe_device(e);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(obs, e->d_obs, e->obs_bytes(), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return GYMCUDA_OK;
}

int gymcuda_get_stats(gymcuda_env* e, gymcuda_stats* out, int reset_counters) {
    ENTER(e);
    if (int _c = clock_pull(e)) return _c;
    if (!out) return fail(GYMCUDA_EINVAL, "out is null");
    CU_TRY(cudaMemcpyAsync(e->h_small, e->d_stats, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    int32_t* pending = reinterpret_cast<int32_t*>(e->h_small + 2);   // (h_small[2] doubles as the done_count scratch of gymcuda_done_indices)
    *pending = 0;
    if (e->stats_pending) CU_TRY(cudaMemcpyAsync(pending, e->d_done_count + ((e->seq - 1) & 1), sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    double* hs = reinterpret_cast<double*>(e->h_small + 4);
    hs[0] = hs[1] = 0.0;
    if (e->d_sums) CU_TRY(cudaMemcpyAsync(hs, e->d_sums, 2 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    out->env_steps = e->env_steps;
    out->episodes = e->h_small[0] + (unsigned long long)*pending;
    out->invalid_actions = e->h_small[1];
    out->return_sum = hs[0];
    out->length_sum = (uint64_t)(hs[1] + 0.5);
    if (reset_counters) {
        CU_TRY(cudaMemsetAsync(e->d_stats, 0, 2 * sizeof(unsigned long long), e->stream));
        if (e->d_sums) CU_TRY(cudaMemsetAsync(e->d_sums, 0, 2 * sizeof(double), e->stream));
        CU_TRY(cudaStreamSynchronize(e->stream));
        e->env_steps = 0;
        e->invalid_seen = 0;
        e->stats_pending = false;   // reported and discarded: the next launch must not add it again
        if (e->device_clock && e->seq > 0) {   // (a captured launch always folds: discard the count itself)
            CU_TRY(cudaMemsetAsync(e->d_done_count + ((e->seq - 1) & 1), 0, sizeof(int32_t), e->stream));
            CU_TRY(cudaStreamSynchronize(e->stream));
        }
    }
    return GYMCUDA_OK;
}

// ------------------------------------------------------------------------------------------------
// observation / reward normalisation (normalize.cuh)
// ------------------------------------------------------------------------------------------------
static int norm_reserve(gymcuda_env* e) {
    if (e->d_norm_acc) return GYMCUDA_OK;
    if (e->ki.od > NORM_MAX_OD) return fail(GYMCUDA_EINVAL, "normalisation supports at most %d observation components", NORM_MAX_OD);
    CU_TRY(cudaMalloc(&e->d_norm_acc, NORM_ACC * sizeof(double)));
    CU_TRY(cudaMalloc(&e->d_norm_ret, (size_t)e->n * 4));
    CU_TRY(cudaMemsetAsync(e->d_norm_acc, 0, NORM_ACC * sizeof(double), e->stream));
    CU_TRY(cudaMemsetAsync(e->d_norm_ret, 0, (size_t)e->n * 4, e->stream));
    return GYMCUDA_OK;
}

int gymcuda_normalize_config(gymcuda_env* e, float gamma, float epsilon, float clip_obs, float clip_reward) {
    ENTER(e);
    if (!(gamma >= 0.0f && gamma <= 1.0f) || !(epsilon > 0.0f) || !(clip_obs > 0.0f) || !(clip_reward > 0.0f))
        return fail(GYMCUDA_EINVAL, "normalize_config: gamma in [0, 1], epsilon > 0, clips > 0");
    e->norm_gamma = gamma; e->norm_eps = epsilon; e->norm_clip_obs = clip_obs; e->norm_clip_reward = clip_reward;
    return GYMCUDA_OK;
}

int gymcuda_normalize_reset(gymcuda_env* e) {
    ENTER(e);
    if (int rc = norm_reserve(e)) return rc;
    CU_TRY(cudaMemsetAsync(e->d_norm_acc, 0, NORM_ACC * sizeof(double), e->stream));
    CU_TRY(cudaMemsetAsync(e->d_norm_ret, 0, (size_t)e->n * 4, e->stream));
    return GYMCUDA_OK;
}

int gymcuda_normalize_device(gymcuda_env* e, float* d_obs, float* d_reward, const uint8_t* d_done, int update) {
    ENTER(e);
    if (!d_obs && !d_reward) return fail(GYMCUDA_EINVAL, "normalize: obs and reward are both null");
    if (int rc = check_device_buffers(e, nullptr, nullptr, d_reward)) return rc;
    if (d_obs && !is_aligned(d_obs, 4)) return fail(GYMCUDA_EINVAL, "device observation buffer must be 4-byte aligned");
    if (int rc = norm_reserve(e)) return rc;
    NormArgs a{};
    a.obs = d_obs; a.reward = d_reward; a.done = d_done; a.ret = e->d_norm_ret; a.acc = e->d_norm_acc; a.n = e->n; a.od = e->ki.od;
    a.gamma = e->norm_gamma; a.eps = e->norm_eps; a.clip_obs = e->norm_clip_obs; a.clip_reward = e->norm_clip_reward;
    const int grid = (e->n + NORM_BLOCK - 1) / NORM_BLOCK;
    if (update) norm_update_kernel<<<grid, NORM_BLOCK, 0, e->stream>>>(a);
    norm_apply_kernel<<<grid, NORM_BLOCK, 0, e->stream>>>(a);
    CU_TRY(cudaGetLastError());
    if (d_obs == e->last_obs) e->last_obs = nullptr;   // the handle's copy now holds normalised values
    return GYMCUDA_OK;
}

int gymcuda_normalize(gymcuda_env* e, float* obs, float* reward, const uint8_t* done, int update) {
    ENTER(e);
    if (!obs && !reward) return fail(GYMCUDA_EINVAL, "normalize: obs and reward are both null");
    const size_t n = (size_t)e->n;
    if (obs) CU_TRY(cudaMemcpyAsync(e->d_obs, obs, e->obs_bytes(), cudaMemcpyHostToDevice, e->stream));
    if (reward) CU_TRY(cudaMemcpyAsync(e->d_reward, reward, n * 4, cudaMemcpyHostToDevice, e->stream));
    if (done) CU_TRY(cudaMemcpyAsync(e->d_done, done, n, cudaMemcpyHostToDevice, e->stream));
    int rc = gymcuda_normalize_device(e, obs ? e->d_obs : nullptr, reward ? e->d_reward : nullptr, done ? e->d_done : nullptr, update);
    if (rc) return rc;
    e->last_obs = nullptr;
    if (obs) CU_TRY(cudaMemcpyAsync(obs, e->d_obs, e->obs_bytes(), cudaMemcpyDeviceToHost, e->stream));
    if (reward) CU_TRY(cudaMemcpyAsync(reward, e->d_reward, n * 4, cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return GYMCUDA_OK;
}

int gymcuda_normalize_get(gymcuda_env* e, double* obs_mean, double* obs_var, double* return_var, double* count) {
    ENTER(e);
    if (int rc = norm_reserve(e)) return rc;
    double acc[NORM_ACC];
    CU_TRY(cudaMemcpyAsync(acc, e->d_norm_acc, sizeof(acc), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    const double c = acc[NORM_VALUES];
    for (int j = 0; j < e->ki.od; ++j) {
        const double mean = c > 0.0 ? acc[j] / c : 0.0;
        double var = c > 0.0 ? acc[NORM_MAX_OD + j] / c - mean * mean : 1.0;
        if (obs_mean) obs_mean[j] = mean;
        if (obs_var) obs_var[j] = var > 0.0 ? var : 0.0;
    }
    if (return_var) {
        const double cr = acc[NORM_VALUES + 1];   // the returns have their own count (a call may carry observations only)
        const double mean = cr > 0.0 ? acc[2 * NORM_MAX_OD] / cr : 0.0;
        const double var = cr > 0.0 ? acc[2 * NORM_MAX_OD + 1] / cr - mean * mean : 1.0;
        *return_var = var > 0.0 ? var : 0.0;
    }
    if (count) *count = c;
    return GYMCUDA_OK;
}

// ------------------------------------------------------------------------------------------------
// streams, pinned memory
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// Env.Render, headless and batched (render.cuh / lunar_render.cuh)
// ------------------------------------------------------------------------------------------------
int gymcuda_render_device(gymcuda_env* e, const int32_t* d_env_ids, int count, int width, int height, uint8_t* d_rgb) {
    ENTER(e);
    TRACE("render_device");
    if (!d_rgb) return fail(GYMCUDA_EINVAL, "d_rgb is null");
    if (count <= 0 || width <= 0 || height <= 0 || width > 4096 || height > 4096) return fail(GYMCUDA_EINVAL, "count, width and height must be positive (at most 4096 x 4096)");
    if (!d_env_ids && count > e->n) return fail(GYMCUDA_EINVAL, "count %d exceeds the %d envs of the handle", count, e->n);
    if (!e->has_state) return fail(GYMCUDA_ESTATE, "Render() before Reset()");
    RenderArgs a{e->d_state, d_env_ids, d_rgb, e->n, count, width, height};
    if (e->cfg.env_kind == GYMCUDA_CARTPOLE) {
        const dim3 grid((unsigned)render_grid_x(width, height), (unsigned)count);
        render_cartpole_kernel<<<grid, RENDER_BLOCK, 0, e->stream>>>(a);
        CU_TRY(cudaGetLastError());
        return GYMCUDA_OK;
    }
#ifdef GYMCUDA_WITH_LUNAR
    if (is_lunar(e)) { CU_TRY(lunar_launch_render(e->stream, a)); return GYMCUDA_OK; }
#endif
    return fail(GYMCUDA_EINVAL, "Render() exists in the reference for CartPoleEnv and LunarLanderEnv only");
}

int gymcuda_render(gymcuda_env* e, const int32_t* env_ids, int count, int width, int height, uint8_t* rgb) {
    ENTER(e);
    if (!rgb) return fail(GYMCUDA_EINVAL, "rgb is null");
    if (count <= 0 || width <= 0 || height <= 0 || width > 4096 || height > 4096) return fail(GYMCUDA_EINVAL, "count, width and height must be positive (at most 4096 x 4096)");
    if (env_ids) for (int k = 0; k < count; ++k) if (env_ids[k] < 0 || env_ids[k] >= e->n) return fail(GYMCUDA_EINVAL, "env id %d out of range [0, %d)", env_ids[k], e->n);
    const size_t bytes = (size_t)count * width * height * 3;
    void *d_rgb = nullptr, *d_ids = nullptr;
    cudaError_t ce = scratch_reserve(e, 0, bytes, &d_rgb);
    if (ce == cudaSuccess && env_ids) ce = scratch_reserve(e, 3, (size_t)count * 4, &d_ids);
    if (ce == cudaSuccess && env_ids) ce = cudaMemcpyAsync(d_ids, env_ids, (size_t)count * 4, cudaMemcpyHostToDevice, e->stream);
    if (ce != cudaSuccess) { cudaGetLastError(); return fail(ce == cudaErrorMemoryAllocation ? GYMCUDA_ENOMEM : GYMCUDA_ECUDA, "render staging failed: %s", cudaGetErrorString(ce)); }
    if (int rc = gymcuda_render_device(e, static_cast<const int32_t*>(d_ids), count, width, height, static_cast<uint8_t*>(d_rgb))) return rc;
    CU_TRY(cudaMemcpyAsync(rgb, d_rgb, bytes, cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return GYMCUDA_OK;
}

int gymcuda_set_device_clock(gymcuda_env* e, int on) {
    ENTER(e);
    if (on) {
        if (e->device_clock) return GYMCUDA_OK;
        if (!e->d_clock) CU_TRY(cudaMalloc(&e->d_clock, 2 * sizeof(unsigned long long)));
        if (!e->stats_pending && e->seq > 0) CU_TRY(cudaMemsetAsync(e->d_done_count + ((e->seq - 1) & 1), 0, sizeof(int32_t), e->stream));
        e->device_clock = true;
        return clock_push(e);
    }
    if (int rc = clock_pull(e)) return rc;   // the host mirrors are exact again
    e->device_clock = false;
    return GYMCUDA_OK;
}

int gymcuda_set_stream(gymcuda_env* e, void* cuda_stream) {
    ENTER(e);
    CU_TRY(cudaStreamSynchronize(e->stream));
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    return GYMCUDA_OK;
}

int gymcuda_sync(gymcuda_env* e) {
    ENTER(e);
    CU_TRY(cudaStreamSynchronize(e->stream));
    e->async_steps = false;
    if (e->h_invalid[1]) { const int who = e->h_invalid[1] - 1; e->h_invalid[1] = 0; return fail(GYMCUDA_ENCCL, "gather wait timed out: rank %d never published its observations", who); }
    if (*e->h_invalid) return step_finish_host(e, nullptr, nullptr, nullptr);   // actions rejected by the asynchronous *_device steps since the last report
    return GYMCUDA_OK;
}

int gymcuda_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(GYMCUDA_EINVAL, "ptr is null");
    CU_TRY(cudaHostAlloc(ptr, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
    return GYMCUDA_OK;
}

int gymcuda_host_free(void* ptr) {
    if (ptr) CU_TRY(cudaFreeHost(ptr));
    return GYMCUDA_OK;
}

int gymcuda_host_register(void* ptr, size_t bytes) {
    if (!ptr || bytes == 0) return fail(GYMCUDA_EINVAL, "gymcuda_host_register: null pointer or empty range");
    CU_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
    return GYMCUDA_OK;
}

int gymcuda_host_unregister(void* ptr) {
    if (!ptr) return GYMCUDA_OK;
    CU_TRY(cudaHostUnregister(ptr));
    return GYMCUDA_OK;
}

// ------------------------------------------------------------------------------------------------
// NCCL all-gather of observations
// ------------------------------------------------------------------------------------------------
int gymcuda_nccl_load(const char* path) { return nccl_load(path); }

int gymcuda_nccl_unique_id(uint8_t id_out[128]) {
    if (!id_out) return fail(GYMCUDA_EINVAL, "id_out is null");
    int rc = nccl_load(nullptr);
    if (rc) return rc;
    NCCL_TRY(g_nccl.GetUniqueId(id_out));
    return GYMCUDA_OK;
}

int gymcuda_comm_init(gymcuda_env* e, const uint8_t id[128], int rank, int world_size) {
    ENTER(e);
    if (!id || world_size <= 0 || rank < 0 || rank >= world_size) return fail(GYMCUDA_EINVAL, "bad rank/world_size");
    int rc = nccl_load(nullptr);
    if (rc) return rc;
    Id128 uid;
    std::memcpy(uid.b, id, 128);
    NCCL_TRY(g_nccl.CommInitRank(&e->comm, world_size, uid, rank));
    e->rank = rank;
    e->world = world_size;
    return GYMCUDA_OK;
}

int gymcuda_allgather_obs(gymcuda_env* e, const float* d_obs, float* d_out) {
    ENTER(e);
    if (!e->comm) return fail(GYMCUDA_ENCCL, "gymcuda_comm_init has not been called on this handle");
    if (!d_out) return fail(GYMCUDA_EINVAL, "d_out is null");
    const float* src = d_obs;
    if (!src) {
        if (!e->last_obs) { int rc = observe_device(e); if (rc) return rc; }
        src = e->last_obs;
    }
    NCCL_TRY(g_nccl.AllGather(src, d_out, (size_t)e->n * e->ki.od, /* ncclFloat32 */ 7, e->comm, e->stream));
    return GYMCUDA_OK;
}

// ------------------------------------------------------------------------------------------------
// fused step + observation gather over peer memory
// ------------------------------------------------------------------------------------------------
int gymcuda_gather_create(gymcuda_env* e, int rank, int world_size, uint8_t handle_out[64]) {
    ENTER(e);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (!handle_out || world_size <= 0 || world_size > MAX_PEERS || rank < 0 || rank >= world_size)
        return fail(GYMCUDA_EINVAL, "bad rank/world_size (at most %d ranks: the GPUs of one box)", MAX_PEERS);
    if (e->g_local) return fail(GYMCUDA_EINVAL, "gather buffer already created");
    const size_t obs_bytes = 2 * (size_t)world_size * e->obs_bytes();
    e->g_flags_off = (obs_bytes + 255) & ~(size_t)255;
    e->g_counter_off = e->g_flags_off + 256;
    CU_TRY(cudaMalloc(&e->g_local, e->g_counter_off + 256));
    CU_TRY(cudaMemset(e->g_local, 0, e->g_counter_off + 256));
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, e->g_local));
    std::memcpy(handle_out, &h, 64);
    e->g_world = world_size; e->g_rank = rank; e->g_seq = 0;
    for (int r = 0; r < MAX_PEERS; ++r) e->g_peer[r] = nullptr;
    e->g_peer[rank] = e->g_local;
    return GYMCUDA_OK;
}

int gymcuda_gather_open(gymcuda_env* e, const uint8_t* handles) {
    ENTER(e);
    if (!e->g_local) return fail(GYMCUDA_EINVAL, "call gymcuda_gather_create first");
    if (!handles) return fail(GYMCUDA_EINVAL, "handles is null");
    for (int r = 0; r < e->g_world; ++r) {
        if (r == e->g_rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * 64, 64);
        CU_TRY(cudaIpcOpenMemHandle(&e->g_peer[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    return GYMCUDA_OK;
}

int gymcuda_step_gather_device(gymcuda_env* e, const void* d_actions, float* d_reward, uint8_t* d_done, const float** d_gathered) {
    ENTER(e);
    if (!d_actions) return fail(GYMCUDA_EINVAL, "d_actions is null");
    if (!e->g_local) return fail(GYMCUDA_EINVAL, "gather buffer not created");
    if (int rc = check_device_buffers(e, d_actions, nullptr, d_reward)) return rc;
    for (int r = 0; r < e->g_world; ++r) if (!e->g_peer[r]) return fail(GYMCUDA_EINVAL, "gymcuda_gather_open has not mapped rank %d", r);
    e->async_steps = true;
    if (is_lunar(e) && e->n >= 4 * STEP_BLOCK) {
        // several kernels per step (contact partition): the step fills this rank's own slot, gather_push_kernel sends it to the peers
        e->g_seq += 1;
        const size_t slot = (size_t)e->n * e->ki.od;
        const size_t at = ((size_t)(e->g_seq & 1u) * e->g_world + e->g_rank) * slot;
        float* own = reinterpret_cast<float*>(e->g_local) + at;
        int rc = step_launch(e, d_actions, 0, 0, own, d_reward ? d_reward : e->d_reward, d_done ? d_done : e->d_done);
        if (rc) return rc;
        PushArgs pa{};
        pa.src = reinterpret_cast<const float4*>(own); pa.world = e->g_world; pa.rank = e->g_rank; pa.gseq = e->g_seq; pa.n4 = slot / 4;
        pa.block_counter = reinterpret_cast<unsigned*>(e->g_local + e->g_counter_off);
        for (int r = 0; r < e->g_world; ++r) {
            pa.dst[r] = reinterpret_cast<float4*>(reinterpret_cast<float*>(e->g_peer[r]) + at);
            pa.flags[r] = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(e->g_peer[r]) + e->g_flags_off);
        }
        const int grid = (int)std::min<size_t>((pa.n4 + 255) / 256, (size_t)e->sm_count * 4);
        gather_push_kernel<<<grid, 256, 0, e->stream>>>(pa);
        CU_TRY(cudaGetLastError());
        const float* base0 = reinterpret_cast<const float*>(e->g_local) + (size_t)(e->g_seq & 1u) * e->g_world * slot;
        e->last_obs = own;
        if (d_gathered) *d_gathered = base0;
        return GYMCUDA_OK;
    }
    int rc = step_launch(e, d_actions, 0, 0, nullptr, d_reward ? d_reward : e->d_reward, d_done ? d_done : e->d_done, true);
    if (rc) return rc;
    const float* base = reinterpret_cast<const float*>(e->g_local) + (size_t)(e->g_seq & 1u) * e->g_world * (size_t)e->n * e->ki.od;
    e->last_obs = base + (size_t)e->g_rank * (size_t)e->n * e->ki.od;
    if (d_gathered) *d_gathered = base;
    return GYMCUDA_OK;
}

int gymcuda_gather_wait(gymcuda_env* e) {
    ENTER(e);
    if (!e->g_local) return fail(GYMCUDA_EINVAL, "gather buffer not created");
    if (e->h_invalid[1]) { const int who = e->h_invalid[1] - 1; e->h_invalid[1] = 0; return fail(GYMCUDA_ENCCL, "an earlier gather wait timed out: rank %d never published its observations", who); }
    gather_wait_kernel<<<1, 32, 0, e->stream>>>(reinterpret_cast<const uint32_t*>(e->g_local + e->g_flags_off), e->g_world, e->g_seq, e->d_invalid_flag + 1);
    CU_TRY(cudaGetLastError());
    return GYMCUDA_OK;
}

}  // extern "C"

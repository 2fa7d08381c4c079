/* libgymcuda -- C ABI of the Blackwell-native vectorised environment engine.
 *
 * This is the drop-in boundary for ONE hot path of SciSharp/Gym.NET: the batched state transition,
 * reward, termination (+ auto-reset and done compaction) of the classic-control environments and
 * LunarLander.  The reference has no FFI for this path -- the seam is its managed API -- so every
 * entry point below names the managed member it stands behind (paths relative to the reference):
 *
 *   Env.Reset / Env.Step / Env.Seed            src/Gym/Envs/Env.cs:20-29
 *   IVecEnv.Reset / Step / Seed / Close        src/Gym/Envs/IVecEnv.cs:8-19, src/Gym/Envs/VecEnv.cs:39-53
 *   VecEnvWrapper.Step (serial per-env loop)   src/Gym/Envs/VecEnvWrapper.cs:22-24
 *   Step {Observation, Reward, Done, Info}     src/Gym/Observations/Step.cs:7-20
 *   CartPoleEnv.Step / Reset / Seed            src/Gym.Environments/Envs/Classic/CartPoleEnv.cs:137-186, 63-67, 196-198
 *   LunarLanderEnv.Step / Reset / Seed         src/Gym.Environments/Envs/Aether/LunarLanderEnv.cs:574-774, 489-572, 900-903
 *   Discrete.Sample / Box.Sample (random policy) src/Gym/Spaces/Discrete.cs:17-28, src/Gym/Spaces/Box.cs:69-90
 *
 * The C# binding a maintainer adds ([DllImport("gymcuda")] + Gym.Environments.Vector.CudaVecEnv) is
 * shown in INTEGRATION.md and shipped under csharp/.
 *
 * Conventions
 *   - plain pointers and sizes only; no exceptions cross the boundary: every call returns
 *     GYMCUDA_OK (0) or a negative gymcuda_status, and gymcuda_last_error() gives the text
 *     (thread-local).  The C# shim maps GYMCUDA_EACTION -> InvalidActionError
 *     (src/Gym/Exceptions/InvalidActionError.cs:7-10), GYMCUDA_EINVAL -> ArgumentException.
 *   - one handle = one GPU = one CUDA stream; calls on a handle must be externally serialised
 *     (reference Env instances are not thread-safe either); distinct handles may be used from
 *     distinct threads or processes (one process per GPU).
 *   - the library owns all device memory; the caller owns host buffers for the duration of a call.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with GYMCUDA_ECUDA.
 *   - layouts: obs [num_envs][obs_dim] float32, reward [num_envs] float32, done [num_envs] uint8 (0/1),
 *     discrete actions int32 [num_envs], box actions float32 [num_envs][act_dim];
 *     rollout trajectories are [k_steps][num_envs][...] in the same element layouts.
 */
#ifndef GYMCUDA_H
#define GYMCUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GYMCUDA_VERSION 111 /* 0.1.11: round-2 additions (step_many, terminal obs, device clock, box sample, render, obs view); purely additive */

typedef enum gymcuda_status {
    GYMCUDA_OK = 0,
    GYMCUDA_EINVAL = -1,   /* bad argument                       -> ArgumentException   */
    GYMCUDA_EACTION = -2,  /* action outside the action space    -> InvalidActionError  */
    GYMCUDA_ECUDA = -3,    /* CUDA runtime / driver failure, or no device                */
    GYMCUDA_ENCCL = -4,    /* NCCL failure or NCCL not loadable                          */
    GYMCUDA_ENOMEM = -5,   /* device or host allocation failed                           */
    GYMCUDA_ESTATE = -6    /* Step before Reset (CartPoleEnv.cs:40,141 would dereference null) */
} gymcuda_status;

typedef enum gymcuda_env_kind {
    GYMCUDA_CARTPOLE = 0,          /* Gym.Environments.Envs.Classic.CartPoleEnv                     */
    GYMCUDA_PENDULUM = 1,          /* north_star env, absent from the reference (README.md:76)      */
    GYMCUDA_MOUNTAINCAR = 2,       /* north_star env, absent from the reference (README.md:75)      */
    GYMCUDA_MOUNTAINCAR_CONT = 3,  /* north_star env, absent from the reference (README.md:74)      */
    GYMCUDA_ACROBOT = 4,           /* north_star env, absent from the reference (README.md:73)      */
    GYMCUDA_LUNARLANDER = 5,       /* Gym.Environments.Envs.Aether.LunarLanderEnv (discrete)         */
    GYMCUDA_LUNARLANDER_CONT = 6   /* LunarLanderEnv(continuous: true)                              */
} gymcuda_env_kind;

enum {
    GYMCUDA_FLAG_AUTO_RESET = 1u,    /* reset in-kernel at `done`; obs returned is the post-reset one */
    GYMCUDA_FLAG_EPISODE_STATS = 2u, /* accumulate return / length of finished episodes on device     */
    GYMCUDA_FLAG_DONE_BITS = 4u      /* done byte: 1 = terminated, 2 = truncated by the time limit only */
};

typedef struct gymcuda_config {
    uint32_t struct_size;   /* = sizeof(gymcuda_config); set by gymcuda_config_default            */
    int32_t env_kind;       /* gymcuda_env_kind                                                    */
    int32_t num_envs;       /* env instances held by THIS handle (this GPU's shard)                */
    int32_t device;         /* CUDA device ordinal                                                 */
    uint64_t seed;          /* Env.Seed; streams are keyed by (seed, global env id)                */
    uint32_t env_id_offset; /* global id of local env 0: rank * num_envs for sharded batches       */
    uint32_t flags;         /* GYMCUDA_FLAG_*                                                      */
    int32_t time_limit;     /* 0 = env default (CartPole: none, like the reference), <0 none, >0 steps */
    float gravity;          /* LunarLanderEnv ctor (LunarLanderEnv.cs:381): [-12, 0], default -10  */
    int32_t enable_wind;    /* LunarLanderEnv ctor                                                  */
    float wind_power;       /* [0, 20], default 15                                                 */
    float turbulence_power; /* [0, 2], default 1.5                                                 */
} gymcuda_config;

typedef struct gymcuda_env gymcuda_env;

typedef struct gymcuda_space_info {
    int32_t obs_dim;       /* ObservationSpace = Box(low, high, float32) of this length           */
    int32_t act_dim;       /* 1 for Discrete                                                      */
    int32_t act_n;         /* Discrete(n); 0 when the action space is a Box                       */
    int32_t state_dim;     /* float32 words per env in get_state/set_state                        */
    int32_t aux_dim;       /* int32 words per env in get_state/set_state                          */
    int32_t time_limit;    /* resolved limit (0 = none)                                           */
    float obs_low[8], obs_high[8];
    float act_low[2], act_high[2];
} gymcuda_space_info;

/* ---- library ------------------------------------------------------------------------------ */
int gymcuda_version(void);
const char* gymcuda_last_error(void);
int gymcuda_device_count(int* count);

/* ---- lifecycle: new XxxEnv(...) / Env.CloseEnvironment / IVecEnv.Close ------------------------ */
int gymcuda_config_default(gymcuda_config* cfg, int env_kind, int num_envs);
int gymcuda_create(const gymcuda_config* cfg, gymcuda_env** out);
int gymcuda_destroy(gymcuda_env* env);
int gymcuda_space(const gymcuda_env* env, gymcuda_space_info* out);
int gymcuda_num_envs(const gymcuda_env* env);

/* ---- Env.Seed(int) / VecEnv.Seed(int[]) ----------------------------------------------------- */
int gymcuda_seed(gymcuda_env* env, uint64_t seed);
int gymcuda_seed_each(gymcuda_env* env, const int32_t* seeds, int n);

/* ---- Env.Reset / IVecEnv.Reset --------------------------------------------------------------- */
/* Host buffers; synchronous.  obs_out may be NULL. */
int gymcuda_reset(gymcuda_env* env, float* obs_out);
int gymcuda_reset_masked(gymcuda_env* env, const uint8_t* mask, float* obs_out);

/* ---- Env.Step / IVecEnv.Step ------------------------------------------------------------------ */
/* Host buffers; H2D(actions) -> kernel -> D2H(obs, reward, done); synchronous.
 * Pageable buffers are staged with cudaMemcpyAsync (one DMA for obs|reward|done when they are adjacent);
 * page-locked buffers (gymcuda_host_alloc, cudaHostAlloc, gymcuda_host_register / cudaHostRegister)
 * are read and written by the kernel directly over PCIe (zero-copy), which is the fast path; the two
 * kinds can be mixed freely among the four buffers.  A page-locked buffer the kernel's vector accesses
 * cannot address in place -- observations not 16-byte aligned, a 2-D Box action array not 8-byte aligned,
 * e.g. a pinned managed float[] whose data starts at 8 mod 16 -- is simply staged like a pageable one.
 * Returns GYMCUDA_EACTION if any action was outside the action space of an env kind that rejects
 * it (all but CartPole, whose reference only Debug.Asserts, CartPoleEnv.cs:139); those envs are
 * left unstepped, the others step normally. */
int gymcuda_step(gymcuda_env* env, const void* actions, float* obs, float* reward, uint8_t* done);
/* Same, all pointers in device memory, asynchronous on the handle's stream; no copies.
 * Alignment (every *_device entry point; cudaMalloc'd and torch-allocated buffers satisfy it): observations
 * 16 bytes, actions 4 bytes (8 for a 2-D Box), rewards 4; anything else is GYMCUDA_EINVAL. */
int gymcuda_step_device(gymcuda_env* env, const void* d_actions, float* d_obs, float* d_reward,
                        uint8_t* d_done);
/* d_obs = GYMCUDA_NO_OBS: the step writes no observation copy at all.  For the envs whose observation IS their state vector
 * (CartPole: CartPoleEnv.cs:183 returns `state`; MountainCar, MountainCarContinuous) a learner on the device reads the
 * current observations in place through gymcuda_obs_view_device -- the step then moves the 41 B per CartPole env step of
 * SURVEY 8d (state read + written, action, reward, done) instead of 57.  For the other envs gymcuda_observe still
 * recomputes the observations from the state on request. */
#define GYMCUDA_NO_OBS ((float*)(uintptr_t)1)
/* *d_obs = device pointer to [num_envs][obs_dim] float32 holding the current observations WITHOUT a copy (the library's
 * state array), valid until gymcuda_destroy and always current on the handle's stream; GYMCUDA_EINVAL for an env kind
 * whose observation is computed from its state (Pendulum, Acrobot, LunarLander). */
int gymcuda_obs_view_device(gymcuda_env* env, const float** d_obs);
/* Broadcast of one action to every env: the shipped IVecEnv.Step(int action) (IVecEnv.cs:14). */
int gymcuda_step_broadcast(gymcuda_env* env, int32_t action, float* obs, float* reward, uint8_t* done);
/* k_steps env steps with CALLER-SUPPLIED actions in ONE launch (the rollout kernel fed from `actions` [k_steps][num_envs]
 * (x act_dim) instead of the in-kernel policy): the state stays in registers between the steps, auto-reset / time limit /
 * statistics as in gymcuda_step, outputs [k_steps][num_envs]... each optional (NULL = not written).  For learners that
 * produce a block of actions at a time (open-loop plans, action repeats, the alternating-action loop of the reference's
 * test, tests/Gym.Tests/Envs/Classic/CartpoleEnvironment.cs:19-26): one launch and no state round trip per step instead
 * of k.  An invalid action leaves that env unstepped for that step (obs = its current observation, reward 0, done 0) and is
 * counted; the host-buffer call then returns GYMCUDA_EACTION like gymcuda_step.  LunarLander: k step launches. */
int gymcuda_step_many_device(gymcuda_env* env, int k_steps, const void* d_actions, float* d_obs, float* d_reward,
                             uint8_t* d_done);
int gymcuda_step_many(gymcuda_env* env, int k_steps, const void* actions, float* obs, float* reward, uint8_t* done);

/* ---- terminal observations under auto-reset (SURVEY 8b "Auto-reset": "optionally the terminal obs in a side buffer") ----
 * With GYMCUDA_FLAG_AUTO_RESET a step that ends an episode returns the POST-reset observation; value bootstrapping at a
 * time-limit truncation needs the observation of the state the episode ended in.  `buffer` [num_envs][obs_dim] float32:
 * every later gymcuda_step* call writes that observation into the rows of the envs whose step returned done (other rows are
 * left as they are).  buffer may be device memory or page-locked host memory (written by the kernel in place; 16-byte
 * aligned), or pageable host memory (staged: copied whole by the host-buffer step calls before they return).  NULL turns
 * the side buffer off.  Not written by the fused rollouts (gymcuda_rollout_random*, gymcuda_step_many*). */
int gymcuda_set_terminal_obs(gymcuda_env* env, float* buffer);

/* ---- fused random-policy rollout (Discrete.Sample / Box.Sample inside the kernel) ------------- */
/* k_steps env steps in ONE launch, state in registers, trajectory streamed out.  Every output
 * pointer is optional (NULL = not written).  *_device: device pointers, asynchronous. */
int gymcuda_rollout_random_device(gymcuda_env* env, int k_steps, float* d_obs, float* d_reward,
                                  uint8_t* d_done, void* d_actions);
int gymcuda_rollout_random(gymcuda_env* env, int k_steps, float* obs, float* reward, uint8_t* done,
                           void* actions);

/* ---- ActionSpace.Sample() on device: Discrete.Sample(mask) / Box.Sample() ---------------------------- */
/* One action per env from the engine's ACTION stream at the current step index -- the draws
 * gymcuda_rollout_random consumes, so `sample; step` loops reproduce a rollout bit for bit.
 * mask: Discrete spaces only, uint8 [num_envs][act_n], entries == 1 are valid (src/Gym/Spaces/Discrete.cs:17-28:
 * uniform over the valid entries, Start when none); a mask on a Box space is GYMCUDA_EINVAL (Box.cs:70-73). */
int gymcuda_sample_actions(gymcuda_env* env, const uint8_t* mask, void* actions_out);
int gymcuda_sample_actions_device(gymcuda_env* env, const uint8_t* d_mask, void* d_actions_out);
/* Box.Sample() of an ARBITRARY box on device (src/Gym/Spaces/Box.cs:69-90), the reference's four-way split per component:
 * low and high finite -> uniform(low, high); only low finite -> low + exponential(1); only high finite -> high +
 * exponential(1) (the reference ADDS the draw to High, Box.cs:83 -- upstream gym subtracts it; the reference's behaviour is
 * kept); neither -> normal(0.5, 1) (Box.cs:81: mean 0.5, the reference's constant).  out [count][dim] float32; sample c,
 * component j is a pure function of (seed, index + c, j): Philox stream 4, one 32-bit word for uniform / exponential
 * (-log1p(-u), u = (w >> 8) * 2^-24), two words for the normal (Box-Muller).  as_int != 0 floors the values (the reference
 * floors for integer dtypes, Box.cs:85-88).  low / high / out: device pointers (*_device, asynchronous on `cuda_stream`) or
 * host pointers (synchronous). */
int gymcuda_box_sample_device(int device, void* cuda_stream, uint64_t seed, uint64_t index, const float* d_low,
                              const float* d_high, int dim, int count, int as_int, float* d_out);
int gymcuda_box_sample(int device, uint64_t seed, uint64_t index, const float* low, const float* high, int dim,
                       int count, int as_int, float* out);

/* ---- done compaction (valid after a step) ------------------------------------------------------ */
/* Indices (local env ids, ascending within a thread block, block order unspecified) of the envs
 * whose last step returned done, and their number.  idx may be NULL to fetch only the count. */
int gymcuda_done_indices(gymcuda_env* env, int32_t* idx, int32_t* count);
int gymcuda_done_indices_device(gymcuda_env* env, const int32_t** d_idx, const int32_t** d_count);

/* ---- snapshot / teacher forcing ----------------------------------------------------------------- */
/* state [num_envs][state_dim] float32, aux [num_envs][aux_dim] int32, t = global step counter. */
int gymcuda_get_state(gymcuda_env* env, float* state, int32_t* aux, uint64_t* t);
int gymcuda_set_state(gymcuda_env* env, const float* state, const int32_t* aux, uint64_t t);
/* Current observations of all envs (no stepping). */
int gymcuda_observe(gymcuda_env* env, float* obs);

/* ---- Env.Render, headless and batched -------------------------------------------------------------
 * What CartPoleEnv.Render (CartPoleEnv.cs:69-135) and LunarLanderEnv.Render (LunarLanderEnv.cs:776-890) draw, rasterised on
 * the device for `count` envs of the batch (env_ids, or envs 0..count-1 when NULL) into rgb [count][height][width][3] uint8:
 * the reference's 600 x 400 canvas sampled at width x height pixel centres (600 x 400 gives the reference's frame; 84 x 84
 * gives the down-sampled picture an image-based agent wants without materialising the large one).  Same primitives, colours
 * and draw order as the reference; coverage is decided at the pixel centre with NO anti-aliasing (ImageSharp's edge blending
 * is not reproduced), LunarLander's exhaust particles are not simulated and not drawn.  The other env kinds have no Render in
 * the reference: GYMCUDA_EINVAL.  *_device: device pointers, asynchronous on the handle's stream. */
int gymcuda_render_device(gymcuda_env* env, const int32_t* d_env_ids, int count, int width, int height, uint8_t* d_rgb);
int gymcuda_render(gymcuda_env* env, const int32_t* env_ids, int count, int width, int height, uint8_t* rgb);

/* ---- episode statistics: what callers keep by hand (examples/.../BasePlaySession.cs:58-69) -- accumulated
 * on device by every step / rollout ------------------------- */
typedef struct gymcuda_stats {
    uint64_t env_steps;      /* env steps executed                */
    uint64_t episodes;       /* episodes finished (done returned) */
    uint64_t invalid_actions;
    double return_sum;       /* GYMCUDA_FLAG_EPISODE_STATS: sum of the returns of finished episodes */
    uint64_t length_sum;     /*                             sum of their lengths                    */
} gymcuda_stats;
int gymcuda_get_stats(gymcuda_env* env, gymcuda_stats* out, int reset_counters);

/* ---- observation / reward normalisation: the wrapper every learner adds around Env.Step (the reference's callers keep
 * their running statistics by hand, examples/.../BasePlaySession.cs:58-69) -- SURVEY 8f rank 3 -------------------------
 * Running mean / variance of every observation component and of the per-env discounted return (ret = ret * gamma + reward,
 * zeroed where `done`), accumulated over all envs and calls on the device (double sums); then, in place,
 *   obs <- clip((obs - mean) / sqrt(var + epsilon), +-clip_obs),   reward <- clip(reward / sqrt(var_ret + epsilon), +-clip_reward).
 * update != 0: the batch is added to the statistics before it is normalised (training); 0: statistics frozen (evaluation).
 * obs / reward may each be NULL (left alone); done may be NULL (no episode ended).  Defaults: gamma 0.99, epsilon 1e-8,
 * clips 10.  *_device: device pointers (the step's own output buffers), asynchronous on the handle's stream. */
int gymcuda_normalize_config(gymcuda_env* env, float gamma, float epsilon, float clip_obs, float clip_reward);
int gymcuda_normalize_device(gymcuda_env* env, float* d_obs, float* d_reward, const uint8_t* d_done, int update);
int gymcuda_normalize(gymcuda_env* env, float* obs, float* reward, const uint8_t* done, int update);
/* obs_mean / obs_var: [obs_dim]; any pointer may be NULL.  count = observations accumulated so far (the returns keep
 * their own count: a call that carries only observations, e.g. the batch Reset returned, does not dilute return_var). */
int gymcuda_normalize_get(gymcuda_env* env, double* obs_mean, double* obs_var, double* return_var, double* count);
int gymcuda_normalize_reset(gymcuda_env* env);

/* ---- streams, pinned memory -------------------------------------------------------------------- */
/* Use the caller's CUDA stream (cudaStream_t) for every launch and copy; NULL restores the
 * handle's own stream (to target the legacy default stream pass cudaStreamLegacy, not 0). */
int gymcuda_set_stream(gymcuda_env* env, void* cuda_stream);
/* CUDA-graph capture of gymcuda_step_device (and gymcuda_step_gather_device): by default the global step index `t` (it keys
 * LunarLander's per-step dispersion draws) and the launch sequence number (it selects the done-list counter) are host
 * state baked into the kernel arguments, so a REPLAYED launch would repeat them.  on != 0 moves both into device memory:
 * the step kernels read them there and a one-thread kernel advances them after every step, so a captured step can be
 * replayed any number of times and each replay is the next step -- bit-identical to the same steps launched one by one.
 * Capture with the handle's stream set to the capturing stream (gymcuda_set_stream); the actions / outputs are the static
 * buffers of the graph.  Every other entry point keeps working in this mode (it synchronises to read the clock back), but
 * only the *_device step calls may be captured.  on == 0 returns to host-side counters (synchronises). */
int gymcuda_set_device_clock(gymcuda_env* env, int on);
int gymcuda_sync(gymcuda_env* env);
int gymcuda_host_alloc(void** ptr, size_t bytes); /* pinned (page-locked) host memory */
int gymcuda_host_free(void* ptr);
/* Page-lock and map memory the CALLER owns (cudaHostRegister) so that gymcuda_step reads / writes it in place:
 * what the C# shim does once with the GCHandle-pinned result arrays it hands back from Step (managed arrays are
 * pageable for CUDA even while pinned for the GC).  The range must stay allocated and pinned until
 * gymcuda_host_unregister.  Each of the four buffers of gymcuda_step is treated on its own: registered ones are
 * accessed directly, the others are staged with a copy. */
int gymcuda_host_register(void* ptr, size_t bytes);
int gymcuda_host_unregister(void* ptr);

/* ---- multi-GPU: optional all-gather of observations over NVLink (one process per GPU) ------- */
/* Rank 0 calls gymcuda_nccl_unique_id and distributes the 128 bytes out of band; every rank then
 * calls gymcuda_comm_init.  nccl_library_path may be NULL (default search: libnccl.so.2). */
int gymcuda_nccl_load(const char* nccl_library_path);
int gymcuda_nccl_unique_id(uint8_t id_out[128]);
int gymcuda_comm_init(gymcuda_env* env, const uint8_t id[128], int rank, int world_size);
/* Gathers every rank's current [num_envs][obs_dim] observations into d_out
 * [world_size][num_envs][obs_dim] (device pointer), asynchronous on the handle's stream.
 * d_obs == NULL gathers the observations of the last step/reset. */
int gymcuda_allgather_obs(gymcuda_env* env, const float* d_obs, float* d_out);

/* ---- fused step + observation all-gather over NVLink peer memory (no NCCL on the step path) ------ */
/* Every rank: gather_create (allocates its double-buffered [world][num_envs][obs_dim] gather buffer and
 * arrival flags, returns a 64-byte cudaIpcMemHandle), exchange the handles out of band, gather_open.
 * gymcuda_step_gather_device then runs the step kernel with the observation stores aimed at slot `rank`
 * of EVERY rank's buffer (peer stores), and *d_gathered receives this rank's [world][num_envs][obs_dim]
 * view of the step; gymcuda_gather_wait enqueues the arrival wait on the handle's stream. */
int gymcuda_gather_create(gymcuda_env* env, int rank, int world_size, uint8_t handle_out[64]);
int gymcuda_gather_open(gymcuda_env* env, const uint8_t* handles /* [world_size][64] */);
int gymcuda_step_gather_device(gymcuda_env* env, const void* d_actions, float* d_reward, uint8_t* d_done,
                               const float** d_gathered);
int gymcuda_gather_wait(gymcuda_env* env);

#ifdef __cplusplus
}
#endif
#endif /* GYMCUDA_H */

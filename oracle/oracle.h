/* TEST INFRASTRUCTURE ONLY.
 *
 * C API of the CPU oracle: per-instance scalar restatements of the reference's Step()/Reset()
 * (src/Gym.Environments/Envs/Classic/CartPoleEnv.cs, src/Gym.Environments/Envs/Aether/LunarLanderEnv.cs)
 * looped serially over instances exactly like the reference's VecEnvWrapper.Step
 * (src/Gym/Envs/VecEnvWrapper.cs:22-24).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product (libgymcuda) never does.
 */
#ifndef GYM_ORACLE_H
#define GYM_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORACLE_CARTPOLE = 0,
    ORACLE_PENDULUM = 1,
    ORACLE_MOUNTAINCAR = 2,
    ORACLE_MOUNTAINCAR_CONT = 3,
    ORACLE_ACROBOT = 4,
    ORACLE_LUNARLANDER = 5,
    ORACLE_LUNARLANDER_CONT = 6
};

/* Arithmetic modes.
 * F64          the reference's arithmetic: double state, double math, libm sin/cos
 *              (C# Math.Sin/Cos), constants rounded through float32 exactly as the C# consts are.
 * F64_F32STORE same arithmetic, state rounded to float32 after every step (the engine stores
 *              float32 state in HBM) -- the mode teacher-forced parity is judged in.
 * F32          the engine's storage-precision arithmetic (detmath v1), bit-identical to the CUDA
 *              kernels; used for bit-for-bit free-running regression at scale. */
enum { ORACLE_MODE_F64 = 0, ORACLE_MODE_F64_F32STORE = 1, ORACLE_MODE_F32 = 2 };

enum { ORACLE_FLAG_AUTO_RESET = 1u, ORACLE_FLAG_DONE_BITS = 4u /* done: 1 terminated, 2 truncated only */ };

typedef struct oracle_env oracle_env;

/* time_limit: 0 = env default (CartPole: none, as the reference; Pendulum 200, MountainCar 200,
 * MountainCarContinuous 999, Acrobot 500), <0 = none, >0 explicit. */
oracle_env* oracle_create(int kind, int num_envs, uint64_t seed, uint32_t env_id_offset,
                          uint32_t flags, int time_limit, int mode);
void oracle_destroy(oracle_env*);
int oracle_dims(int kind, int* state_dim, int* aux_dim, int* obs_dim, int* act_dim, int* act_n);
void oracle_seed(oracle_env*, uint64_t seed);
void oracle_seed_each(oracle_env*, const int32_t* seeds);
void oracle_set_threads(oracle_env*, int threads);
/* LunarLanderEnv ctor arguments (LunarLanderEnv.cs:381); call before the first reset. */
void oracle_set_lunar_params(oracle_env*, float gravity, int use_wind, float wind_power, float turbulence_power);

void oracle_reset(oracle_env*, float* obs);
void oracle_reset_masked(oracle_env*, const uint8_t* mask, float* obs);
/* actions: int32[n] (discrete) or float[n*act_dim] (box).  Returns the number of invalid actions. */
int oracle_step(oracle_env*, const void* actions, float* obs, float* reward, uint8_t* done);
/* K steps with pre-generated actions [K][n][act_dim] and no outputs (CPU-baseline timing loop). */
int oracle_step_many(oracle_env*, int k_steps, const void* actions);
/* K steps of the random policy; any output pointer may be NULL.  Layouts [K][n][dim]. */
void oracle_rollout_random(oracle_env*, int k_steps, float* obs, float* reward, uint8_t* done,
                           void* actions);
/* ActionSpace.Sample() of every instance at the current step index; mask (Discrete only) may be NULL. */
void oracle_sample_actions(oracle_env*, const uint8_t* mask, void* actions_out);
/* state[n][state_dim] as doubles (float32 values are exactly representable), aux[n][aux_dim]. */
void oracle_get_state(oracle_env*, double* state, int32_t* aux, uint64_t* t);
void oracle_set_state(oracle_env*, const double* state, const int32_t* aux, uint64_t t);

/* raw generator access for known-answer tests */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void oracle_draw(uint64_t seed, uint32_t env_id, uint64_t index, uint32_t stream, uint32_t sub,
                 uint32_t out[4]);
void oracle_sincosf(const float* x, float* s, float* c, size_t n);

#ifdef __cplusplus
}
#endif
#endif

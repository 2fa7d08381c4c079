/* A plain C99 caller of the C ABI: what any FFI (P/Invoke, ctypes, cgo ...) binds, with no Python in between.
 *
 *   make -C examples            (links against gym.net_b200/csrc/libgymcuda.so)
 *   examples/c_driver [num_envs] [steps]
 *
 * Runs the reference's caller loop (README.md:32-52 / tests/Gym.Tests/Envs/Classic/CartpoleEnvironment.cs:19-30:
 * alternating actions i % 2, episodes restarted when done) for a CartPole batch through host buffers, then the
 * fused random-policy rollout, and prints env-steps/s.  Without a CUDA device gymcuda_create fails with
 * GYMCUDA_ECUDA and the program says so: the library has no CPU path. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "../include/gymcuda.h"

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != GYMCUDA_OK) {                                                     \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, gymcuda_last_error());     \
            return rc_ == GYMCUDA_ECUDA ? 3 : 1;                                     \
        }                                                                            \
    } while (0)

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 65536;
    const int steps = argc > 2 ? atoi(argv[2]) : 1000;
    gymcuda_config cfg;
    gymcuda_env* env = NULL;
    CHECK(gymcuda_config_default(&cfg, GYMCUDA_CARTPOLE, n));
    cfg.flags = GYMCUDA_FLAG_AUTO_RESET;
    cfg.seed = 0;
    CHECK(gymcuda_create(&cfg, &env));

    /* page-locked host buffers: the step kernel reads / writes them in place over PCIe */
    int32_t* actions; float* obs; float* reward; uint8_t* done;
    void* p;
    CHECK(gymcuda_host_alloc(&p, (size_t)n * 4)); actions = (int32_t*)p;
    CHECK(gymcuda_host_alloc(&p, (size_t)n * 16)); obs = (float*)p;
    CHECK(gymcuda_host_alloc(&p, (size_t)n * 4)); reward = (float*)p;
    CHECK(gymcuda_host_alloc(&p, (size_t)n)); done = (uint8_t*)p;

    CHECK(gymcuda_reset(env, obs));
    unsigned long long episodes = 0;
    double t0 = now_s();
    for (int t = 0; t < steps; ++t) {
        for (int i = 0; i < n; ++i) actions[i] = t % 2;
        CHECK(gymcuda_step(env, actions, obs, reward, done));
        for (int i = 0; i < n; ++i) episodes += done[i];
    }
    double dt = now_s() - t0;
    printf("gymcuda_step      : %d envs x %d steps, %llu episodes, %.3e env-steps/s (host buffers)\n",
           n, steps, episodes, (double)n * steps / dt);

    gymcuda_stats st;
    t0 = now_s();
    CHECK(gymcuda_rollout_random_device(env, 512, NULL, NULL, NULL, NULL));   /* state only: nothing streamed out */
    CHECK(gymcuda_sync(env));
    dt = now_s() - t0;
    CHECK(gymcuda_get_stats(env, &st, 0));
    printf("rollout_random    : %d envs x 512 steps in one launch, %.3e env-steps/s, %llu episodes so far\n",
           n, (double)n * 512 / dt, (unsigned long long)st.episodes);

    gymcuda_host_free(actions); gymcuda_host_free(obs); gymcuda_host_free(reward); gymcuda_host_free(done);
    CHECK(gymcuda_destroy(env));
    return 0;
}

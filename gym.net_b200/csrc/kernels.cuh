// Generic kernels of the engine: one thread = one env instance.
//
//   step_kernel     one env step per launch: coalesced load of the state, transition, reward,
//                   termination, in-kernel auto-reset, coalesced stores, then warp-ballot +
//                   block-scan compaction of the done mask (one atomicAdd per block).
//                   Replaces the serial per-env loop of VecEnvWrapper.Step
//                   (src/Gym/Envs/VecEnvWrapper.cs:22-24) over Env.Step (src/Gym/Envs/Env.cs:21).
//   rollout_kernel  k fused steps of the random policy (Discrete.Sample / Box.Sample,
//                   src/Gym/Spaces/Discrete.cs:27, src/Gym/Spaces/Box.cs:84): state stays in registers,
//                   actions come from the per-env Philox stream, the trajectory is streamed to HBM
//                   with evict-first stores.  This is the caller loop of the reference's tests
//                   (tests/Gym.Tests/Envs/Classic/CartpoleEnvironment.cs:19-30) moved on-device.
//   reset_kernel    Env.Reset for all / masked envs; with an all-zero mask it only recomputes the
//                   observations from the stored state.
//   ctor_kernel     per-env constructor draws (LunarLander's wind phase, LunarLanderEnv.cs:409-410).
//
// HBM layout (structure of arrays, all indexed by local env id):
//   state   classic: Vec[n], float4 (CartPole, Acrobot) or float2 (Pendulum, MountainCar*)
//           LunarLander: float[SD][n] field-major (lunar.cuh)
//   aux     LunarLander only: int32[AUXW][n] field-major (contact flags, limit states, manifold ids)
//   sbd     int32[n]    CartPole steps_beyond_done (touched only when auto-reset is off)
//   ep_t    int32[n]    episode step counter (touched only when a time limit is set)
//   episode int32[n]    number of resets the env has had = index of its next RESET draw
//                       (touched only by lanes that reset)
//   seeds   int32[n]    optional per-env seeds (VecEnv.Seed(int[]))
// No generator state lives in memory: action draws are functions of (seed, env id, t), reset draws of
// (seed, env id, episode ordinal) -- which is what lets the rollout kernel pre-generate the next
// initial state and refill it for the whole warp at once instead of diverging at every `done`.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

#include "env_classic.cuh"
#include "philox.cuh"

namespace gymcuda {

constexpr int MAX_PEERS = 8;   // GPUs of one NVSwitch box

struct StepArgs {
    void* state;
    int32_t* aux;
    int32_t* sbd;
    int32_t* ep_t;
    int32_t* episode;
    const int32_t* seeds;
    const int32_t* perm;               // thread -> env map (LunarLander contact partition), null = identity
    const int32_t* split;              // with perm: device count of the envs in the first class of the partition
    int part;                          // 0: every position; 1: positions [0, *split) (first class); 2: positions [*split, n)
    int sample;                        // 1: the action is ActionSpace.Sample() of step t (the rollout's random policy), `actions` is not read
    void* act_out;                     // sampled actions are also written here (may be null)
    float* terminal_obs;               // [n][OD], may be null: under auto-reset, the observation of the TERMINAL state of the envs whose step returned done (the returned obs is the post-reset one)
    const void* actions;
    float* obs;
    float* reward;
    uint8_t* done;
    int32_t* done_idx;                 // may be null; written directly only by the warp-granular epilogue (LunarLander)
    int32_t* blk_cnt;                  // [CTAs + 1] finished episodes per CTA of this launch   } the CTA-level epilogue: the compact list is
    int32_t* tmp_idx;                  // [CTAs * BLOCK] every CTA's own compact sub-list        } built on demand by done_list_scatter_kernel
    int32_t* done_count;               // [2], indexed by the parity of `seq`
    unsigned long long* stats;         // [0] episodes finished, [1] invalid actions
    float* ep_ret;                     // per-env running episode return (null: statistics off)
    double* sums;                      // [0] sum of finished-episode returns, [1] sum of their lengths
    int done_bits;                     // 1: done byte = 1 terminated, 2 truncated by the time limit only
    int* host_invalid;                 // mapped host flag raised when an action is rejected
    int n;
    uint32_t env_off;
    uint64_t seed;
    uint64_t t;
    int limit;
    int use_bcast;
    int32_t bcast_action;
    uint32_t seq;                      // step-launch sequence number of this handle
    int fold_prev;                     // 1: done_count[(seq + 1) & 1] still holds the episode count of the previous step launch, not yet added to stats[0]
    const unsigned long long* clock;   // device-resident {t, seq} (gymcuda_set_device_clock): read instead of the baked `t` / `seq`, so that a captured launch can be replayed
    EnvParams prm;
    // fused observation gather over NVLink peer memory (world == 0: off)
    int world, rank;
    uint32_t gseq;                     // gather sequence number (parity selects the double buffer)
    float* peer_obs[MAX_PEERS];        // every rank's gather buffer [2][world][n][OD], mapped with cudaIpc
    uint32_t* peer_flags[MAX_PEERS];   // every rank's arrival flags [world]
    unsigned* block_counter;           // last-block detection
};

struct RolloutArgs {
    void* state;
    int32_t* aux;
    int32_t* sbd;
    int32_t* ep_t;
    int32_t* episode;
    const int32_t* seeds;
    const int32_t* perm;   // thread -> env map, null = identity
    float* obs;        // [k][n][OD]   may be null
    float* reward;     // [k][n]       may be null
    uint8_t* done;     // [k][n]       may be null
    void* actions;     // [k][n][AD]   may be null
    const void* actions_in;   // [k][n][AD] caller-supplied actions (gymcuda_step_many*): null = the in-kernel random policy.  Generic variant (ALL_OUT = false) only.
    int* host_invalid;        // with actions_in: mapped host flag raised when an action is rejected
    unsigned long long* stats;
    float* ep_ret;
    double* sums;
    int done_bits;
    int n;
    int k_steps;
    uint32_t env_off;
    uint64_t seed;
    uint64_t t;
    int limit;
    EnvParams prm;
};

struct ResetArgs {
    void* state;
    int32_t* aux;
    int32_t* sbd;
    int32_t* ep_t;
    int32_t* episode;
    const int32_t* seeds;
    const uint8_t* mask;   // may be null = all
    float* obs;            // may be null
    float* ep_ret;         // may be null
    int n;
    uint32_t env_off;
    uint64_t seed;
    uint64_t t;
    EnvParams prm;
};

// ---------------------------------------------------------------- observation stores
template <int OD, bool STREAM>
__device__ __forceinline__ void store_obs(float* base, size_t env_index, const float* o) {
    float* p = base + env_index * OD;
    if (OD == 4) {
        const float4 v = make_float4(o[0], o[1], o[2], o[3]);
        if (STREAM) __stcs(reinterpret_cast<float4*>(p), v); else *reinterpret_cast<float4*>(p) = v;
    } else if (OD == 8) {
        const float4 v0 = make_float4(o[0], o[1], o[2], o[3]), v1 = make_float4(o[4], o[5], o[6], o[7]);
        if (STREAM) { __stcs(reinterpret_cast<float4*>(p), v0); __stcs(reinterpret_cast<float4*>(p) + 1, v1); }
        else { reinterpret_cast<float4*>(p)[0] = v0; reinterpret_cast<float4*>(p)[1] = v1; }
    } else if (OD % 2 == 0) {
#pragma unroll
        for (int k = 0; k < OD / 2; ++k) {
            const float2 v = make_float2(o[2 * k], o[2 * k + 1]);
            if (STREAM) __stcs(reinterpret_cast<float2*>(p) + k, v); else reinterpret_cast<float2*>(p)[k] = v;
        }
    } else {
#pragma unroll
        for (int k = 0; k < OD; ++k) { if (STREAM) __stcs(p + k, o[k]); else p[k] = o[k]; }
    }
}

// ---------------------------------------------------------------- actions
template <class Act> struct ActCast { __device__ static __forceinline__ Act from_int(int32_t a) { return (Act)a; } };
template <> struct ActCast<float2> { __device__ static __forceinline__ float2 from_int(int32_t a) { return make_float2((float)a, 0.0f); } };

template <class E> struct ActIO {
    using Act = typename E::Act;
    __device__ static __forceinline__ Act load(const void* base, int i) { return reinterpret_cast<const Act*>(base)[i]; }
    __device__ static __forceinline__ Act load_at(const void* base, size_t idx) { return __ldcs(reinterpret_cast<const Act*>(base) + idx); }
    __device__ static __forceinline__ Act bcast(int32_t a) { return ActCast<Act>::from_int(a); }
    __device__ static __forceinline__ void store(void* base, size_t idx, Act a) { __stcs(reinterpret_cast<Act*>(base) + idx, a); }
};

// Random policy: Discrete.Sample = randint(0, N) (Discrete.cs:27), Box.Sample = uniform(low, high) (Box.cs:84).
// Draw t of env e comes from block (t >> SHIFT) of the ACTION stream; the block is regenerated only
// when t crosses a block boundary (CartPole: once per 128 steps).
//   init(t)     makes the generator valid for step index t (one Philox evaluation)
//   next(t)     the draw of step t; calls must be consecutive in t after init
//   at<J>(tc)   the same draw for step tc + J of an 8-step chunk (tc % 8 == 0, J = 0..7 in order): block and
//               word boundaries can only fall on J == 0 (or J == 4), so the unrolled chunk carries no per-step tests
template <class E, int ACTN = E::ACTN, int AD = E::AD> struct ActionGen;

template <class E> struct ActionGen<E, 2, 1> {
    Block b;
    uint32_t bits;   // current word, already shifted so that bit 0 is the draw of the next step
    __device__ __forceinline__ void init(uint64_t seed, uint32_t gid, uint64_t t) {
        b = draw(seed, gid, t >> 7, STREAM_ACTION);
        bits = word(b, ((uint32_t)t >> 5) & 3u) >> ((uint32_t)t & 31u);
    }
    __device__ __forceinline__ void refill(uint64_t seed, uint32_t gid, uint64_t t) {   // t % 32 == 0
        if (((uint32_t)t & 127u) == 0) b = draw(seed, gid, t >> 7, STREAM_ACTION);
        bits = word(b, ((uint32_t)t >> 5) & 3u);
    }
    __device__ __forceinline__ int32_t next(uint64_t seed, uint32_t gid, uint64_t t) {
        if (((uint32_t)t & 31u) == 0) refill(seed, gid, t);
        const int32_t a = (int32_t)(bits & 1u);
        bits >>= 1;
        return a;
    }
    template <int J> __device__ __forceinline__ int32_t at(uint64_t seed, uint32_t gid, uint64_t tc) {
        if (J == 0 && ((uint32_t)tc & 31u) == 0) refill(seed, gid, tc);
        const int32_t a = (bits & (1u << J)) ? 1 : 0;   // one predicate-setting LOP3 feeds both the force select and the stored action
        if (J == 7) bits >>= 8;
        return a;
    }
};
template <class E> struct ActionGen<E, 4, 1> {
    Block b;
    uint32_t bits;
    __device__ __forceinline__ void init(uint64_t seed, uint32_t gid, uint64_t t) {
        b = draw(seed, gid, t >> 6, STREAM_ACTION);
        bits = word(b, ((uint32_t)t >> 4) & 3u) >> (2u * ((uint32_t)t & 15u));
    }
    __device__ __forceinline__ void refill(uint64_t seed, uint32_t gid, uint64_t t) {   // t % 16 == 0
        if (((uint32_t)t & 63u) == 0) b = draw(seed, gid, t >> 6, STREAM_ACTION);
        bits = word(b, ((uint32_t)t >> 4) & 3u);
    }
    __device__ __forceinline__ int32_t next(uint64_t seed, uint32_t gid, uint64_t t) {
        if (((uint32_t)t & 15u) == 0) refill(seed, gid, t);
        const int32_t a = (int32_t)(bits & 3u);
        bits >>= 2;
        return a;
    }
    template <int J> __device__ __forceinline__ int32_t at(uint64_t seed, uint32_t gid, uint64_t tc) {
        if (J == 0 && ((uint32_t)tc & 15u) == 0) refill(seed, gid, tc);
        const int32_t a = (int32_t)((bits >> (2 * J)) & 3u);
        if (J == 7) bits >>= 16;
        return a;
    }
};
// one 32-bit word per draw, four draws per block (Discrete(3), Box 1-D)
template <class E> struct WordGen {
    Block b;
    __device__ __forceinline__ void init(uint64_t seed, uint32_t gid, uint64_t t) { b = draw(seed, gid, t >> 2, STREAM_ACTION); }
    __device__ __forceinline__ uint32_t next_word(uint64_t seed, uint32_t gid, uint64_t t) {
        if (((uint32_t)t & 3u) == 0) b = draw(seed, gid, t >> 2, STREAM_ACTION);
        return word(b, (uint32_t)t & 3u);
    }
    template <int J> __device__ __forceinline__ uint32_t word_at(uint64_t seed, uint32_t gid, uint64_t tc) {
        if ((J & 3) == 0) b = draw(seed, gid, (tc + J) >> 2, STREAM_ACTION);
        return (J & 3) == 0 ? b.w0 : ((J & 3) == 1 ? b.w1 : ((J & 3) == 2 ? b.w2 : b.w3));
    }
};
template <class E> struct ActionGen<E, 3, 1> : WordGen<E> {
    __device__ __forceinline__ int32_t next(uint64_t seed, uint32_t gid, uint64_t t) { return (int32_t)__umulhi(this->next_word(seed, gid, t), 3u); }
    template <int J> __device__ __forceinline__ int32_t at(uint64_t seed, uint32_t gid, uint64_t tc) {
        return (int32_t)__umulhi(this->template word_at<J>(seed, gid, tc), 3u);
    }
};
template <class E> __device__ __forceinline__ float box_uniform(uint32_t w) { return uniformf(E::ACT_LOW, E::ACT_HIGH, w); }
template <class E> struct ActionGen<E, 0, 1> : WordGen<E> {
    __device__ __forceinline__ float next(uint64_t seed, uint32_t gid, uint64_t t) { return box_uniform<E>(this->next_word(seed, gid, t)); }
    template <int J> __device__ __forceinline__ float at(uint64_t seed, uint32_t gid, uint64_t tc) {
        return box_uniform<E>(this->template word_at<J>(seed, gid, tc));
    }
};
template <class E> struct ActionGen<E, 0, 2> {
    Block b;
    __device__ __forceinline__ void init(uint64_t seed, uint32_t gid, uint64_t t) { b = draw(seed, gid, t >> 1, STREAM_ACTION); }
    __device__ __forceinline__ float2 next(uint64_t seed, uint32_t gid, uint64_t t) {
        if (((uint32_t)t & 1u) == 0) b = draw(seed, gid, t >> 1, STREAM_ACTION);
        const uint32_t j = 2u * ((uint32_t)t & 1u);
        return make_float2(box_uniform<E>(word(b, j)), box_uniform<E>(word(b, j + 1)));
    }
    template <int J> __device__ __forceinline__ float2 at(uint64_t seed, uint32_t gid, uint64_t tc) { return next(seed, gid, tc + J); }
};

__device__ __forceinline__ uint64_t seed_of(const int32_t* seeds, uint64_t seed, int i) {
    return seeds ? (uint64_t)(uint32_t)seeds[i] : seed;
}

// ---------------------------------------------------------------- step
constexpr int STEP_BLOCK = 128;
#ifndef GYMCUDA_STEP_BLOCK_BIG
#define GYMCUDA_STEP_BLOCK_BIG 128
#endif
constexpr int STEP_BLOCK_BIG = GYMCUDA_STEP_BLOCK_BIG;   // CTA of the batches of STEP_BIG_BATCH envs and more (classic envs)
constexpr int STEP_BIG_BATCH = 1 << 20;

// envs whose step can fold the auto-reset in (E::FUSED_RESET + E::step_ar): LunarLander
template <class E, class = void> struct FusedReset : std::false_type {};
template <class E> struct FusedReset<E, std::enable_if_t<E::FUSED_RESET>> : std::true_type {};

// envs stepped by three lanes each (E::TRIO: the contact class of LunarLander, lunar_core.cuh): a warp holds ten envs in its
// lanes 0..29 (lane = 3 * slot + sub), lanes 30 and 31 idle; the three lanes of an env compute the same values and the first
// one (sub == 0) does the global stores and takes part in the done / invalid votes
template <class E, class = void> struct TrioEnv : std::false_type {};
template <class E> struct TrioEnv<E, std::enable_if_t<E::TRIO>> : std::true_type {};
constexpr int TRIO_ENVS_PER_WARP = 10;

// STEP_BLOCK = the launch shape at every size.  (BLOCK is a parameter because larger CTAs were measured for the batches of a
// million envs and more -- fewer same-address atomics on the done counter: 128 / 256 / 512 / 1024 threads give 160 / 166 / 182 /
// 238 us per 16.7 M-env CartPole launch without finished episodes, so STEP_BLOCK_BIG stays at 128.)
template <class E, bool AUTO_RESET, bool LIMIT, int BLOCK = STEP_BLOCK>
__global__ void __launch_bounds__(BLOCK) step_kernel(const StepArgs p) {
    // step index and launch sequence number: baked arguments, or the device-resident clock a CUDA-graph replay advances
    const uint64_t now_t = p.clock ? (uint64_t)p.clock[0] : p.t;
    const uint32_t now_seq = p.clock ? (uint32_t)p.clock[1] : p.seq;
    using S = typename E::S;
    using Act = typename E::Act;
    int tix = blockIdx.x * BLOCK + threadIdx.x;
    constexpr bool TRIO = TrioEnv<E>::value;
    bool primary = true;               // this lane writes the env's results (TRIO: the first lane of the three)
    if constexpr (TRIO) {
        const unsigned ln = threadIdx.x & 31u;
        primary = ln % 3u == 0u;
        tix = ln < 3u * TRIO_ENVS_PER_WARP ? (tix >> 5) * TRIO_ENVS_PER_WARP + (int)(ln / 3u) : 0x3fffffff;
    }
    int hi = p.n;
    if (p.part == 1) hi = *p.split;
    if (p.part == 2) tix += *p.split;
    if (tix >= hi) tix = p.n;          // not in this launch's class: idle (the compaction below still needs the whole block)
    int i = tix;
    bool done = false;
    bool invalid = false;
    bool trunc_only = false;
    // Auto-reset of the classic envs is DEFERRED to the end of the kernel: a reset is a Philox block + the uniform maps
    // (~100 instructions) that a warp pays in full as soon as ONE of its lanes finished an episode (CartPole under the
    // random policy: 4.5 % of the envs per step, i.e. 77 % of the warps); the done envs of the whole CTA are instead listed
    // in shared memory (the ranks of the done compaction below) and reset side by side by its first lanes -- one warp
    // instead of four pays, and the 16 M-env step drops from issue-bound to DRAM-bound.  Same draws, same results.
    constexpr bool DEFER = AUTO_RESET && !FusedReset<E>::value;
    bool deferred = false;
    float fin_ret = 0.0f;
    int32_t fin_len = 0;
    uint8_t done_byte = 0;
    if (tix < p.n) {
        if (p.perm) i = p.perm[tix];
        S s = E::load(p.state, p.aux, p.n, i, p.prm);
        Act a;
        if (p.sample) {
            ActionGen<E> gen;
            const uint64_t sseed = seed_of(p.seeds, p.seed, i);
            gen.init(sseed, p.env_off + (uint32_t)i, now_t);
            a = gen.next(sseed, p.env_off + (uint32_t)i, now_t);
            if (p.act_out && primary) reinterpret_cast<Act*>(p.act_out)[i] = a;
        } else {
            a = p.use_bcast ? ActIO<E>::bcast(p.bcast_action) : ActIO<E>::load(p.actions, i);
        }
        int32_t sbd = -1;
        if (E::HAS_SBD && !AUTO_RESET) sbd = p.sbd[i];
        int32_t ept = 0;
        if (LIMIT) ept = p.ep_t[i];
        StepOut r{0.0f, 0u};
        invalid = E::REJECT_INVALID && !E::valid(a);
        if (!invalid) {
            const uint64_t seed = seed_of(p.seeds, p.seed, i);
            const uint32_t gid = p.env_off + (uint32_t)i;
            int32_t ep_ord = 0;
            if constexpr (FusedReset<E>::value) {
                if (AUTO_RESET) ep_ord = p.episode[i];
                const bool allow = AUTO_RESET && p.terminal_obs == nullptr;
                r = E::step_ar(s, a, seed, gid, now_t, allow, (uint32_t)ep_ord);
            } else {
                r = E::step(s, a, sbd, seed, gid, now_t);
            }
            if (LIMIT) { ept += 1; if (ept >= p.limit && !r.done) { r.done = 1u; trunc_only = true; } }   // truncation folded into done
            if (p.ep_ret) {   // episode statistics (the caller-side bookkeeping of BasePlaySession.cs:58-69)
                float ret = p.ep_ret[i] + r.reward;
                if (r.done) { fin_ret = ret; fin_len = ept; ret = 0.0f; }
                if (primary) p.ep_ret[i] = ret;
            }
            if (AUTO_RESET && r.done) {
                if constexpr (DEFER) {
                    if (p.terminal_obs) { float to[E::OD]; E::obs(s, to); store_obs<E::OD, false>(p.terminal_obs, (size_t)i, to); }
                    deferred = true;   // new state, episode ordinal and observation: written by the CTA's reset pass below
                } else {
                    const int32_t ep = FusedReset<E>::value ? ep_ord : p.episode[i];
                    if (!r.did_reset) {
                        if (p.terminal_obs && primary) { float to[E::OD]; E::obs(s, to); store_obs<E::OD, false>(p.terminal_obs, (size_t)i, to); }
#ifndef GYMCUDA_PROBE_SKIP_RESET   // (timing probe only, never defined in the product build: what the in-kernel reset of a finished env costs the launch)
                        E::reset(s, seed, gid, (uint32_t)ep, now_t + 1, p.prm);
#endif
                    }
                    if (primary) p.episode[i] = ep + 1;
                }
                sbd = -1;
                ept = 0;
            }
            if (!deferred && primary) E::store(p.state, p.aux, p.n, i, s);
            if (E::HAS_SBD && !AUTO_RESET) p.sbd[i] = sbd;
            if (LIMIT && primary) p.ep_t[i] = ept;
        }
        if (!deferred && primary) {
            float o[E::OD];
            E::obs(s, o);
            if (p.obs) store_obs<E::OD, false>(p.obs, (size_t)i, o);
            // step + all-gather in one kernel: the observation goes straight into slot `rank` of EVERY rank's
            // gather buffer with peer stores over NVLink (the local copy is just the peer == rank case)
            for (int r = 0; r < p.world; ++r)
                store_obs<E::OD, false>(p.peer_obs[r], ((size_t)(p.gseq & 1u) * p.world + p.rank) * (size_t)p.n + (size_t)i, o);
        }
        if (primary) p.reward[i] = r.reward;
        done_byte = (p.done_bits && trunc_only) ? (uint8_t)2 : (uint8_t)r.done;
        done = r.done != 0;
        if constexpr (TRIO) { done = done && primary; invalid = invalid && primary; }   // one vote per env
    }

    // ---- envs whose step is long and uneven (LunarLander: FusedReset): WARP-granular epilogue, no CTA barrier -- a warp
    // whose landers are done retires at once and frees its registers instead of waiting at __syncthreads() for the
    // slowest lander of the CTA (ncu, round 2: 7-10 % of the warp time of both lunar kernels sat at that barrier).
    // One atomicAdd per warp with a finished episode: 2048 warps per 65 536-lander launch, no contention to speak of.
    if constexpr (FusedReset<E>::value) {
        if (p.world == 0) {   // (the fused gather needs the CTA-level "last block" signal: general path below)
            const unsigned lane_w = threadIdx.x & 31;
            const unsigned mw = __ballot_sync(0xffffffffu, done);
            const unsigned miw = __ballot_sync(0xffffffffu, invalid);
            if (tix < p.n && primary) p.done[i] = done_byte;
            if (miw != 0 && lane_w == 0) {
                atomicAdd(&p.stats[1], (unsigned long long)__popc(miw));
                *reinterpret_cast<volatile int*>(p.host_invalid) = 1;
                __threadfence_system();
            }
            if (mw != 0) {
                int base = 0;
                if (lane_w == 0) base = atomicAdd(p.done_count + (now_seq & 1), __popc(mw));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (done && p.done_idx != nullptr) p.done_idx[base + __popc(mw & ((1u << lane_w) - 1u))] = i;
                if (p.sums != nullptr) {
                    double rs = done ? (double)fin_ret : 0.0, ls = done ? (double)fin_len : 0.0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) { rs += __shfl_xor_sync(0xffffffffu, rs, o); ls += __shfl_xor_sync(0xffffffffu, ls, o); }
                    if (lane_w == 0) { atomicAdd(&p.sums[0], rs); atomicAdd(&p.sums[1], ls); }
                }
            }
            if (blockIdx.x == 0 && threadIdx.x == 0 && p.part != 2) {   // fold the previous launch's count, zero the next launch's counter (see below)
                int32_t* prev = p.done_count + ((now_seq + 1) & 1);
                if (p.fold_prev) p.stats[0] += (unsigned long long)*prev;
                *prev = 0;
            }
            return;
        }
    }

    // ---- done compaction: warp ballot + popc prefix -> block scan -> one atomicAdd per block
    __shared__ int warp_cnt[BLOCK / 32];
    __shared__ __align__(16) uint8_t done_tile[BLOCK];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, done);
    const unsigned mi = __ballot_sync(0xffffffffu, invalid);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    // the done bytes of the CTA go out as one 128 B line (32 lanes x 4 B) instead of four 32 B pieces: over PCIe
    // (zero-copy host buffers) every store instruction is a packet, and over HBM it is one full sector group
    const bool packed = !TRIO && p.perm == nullptr && p.part == 0 && (reinterpret_cast<uintptr_t>(p.done) & 3u) == 0;
    if (packed) done_tile[threadIdx.x] = done_byte;
    else if (tix < p.n && primary) p.done[i] = done_byte;
    if (mi != 0 && lane == 0) {
        atomicAdd(&p.stats[1], (unsigned long long)__popc(mi));
        *reinterpret_cast<volatile int*>(p.host_invalid) = 1;
        __threadfence_system();
    }
    __syncthreads();
    int warp_off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; ++w) {
        const int c = warp_cnt[w];
        if (w < (int)warp) warp_off += c;
        total += c;
    }
    if (packed && threadIdx.x < BLOCK / 4) {
        const int first = blockIdx.x * BLOCK + 4 * (int)threadIdx.x;
        if (first + 3 < p.n) *reinterpret_cast<uint32_t*>(p.done + first) = *reinterpret_cast<const uint32_t*>(done_tile + 4 * threadIdx.x);
        else for (int k = first; k < p.n && k < first + 4; ++k) p.done[k] = done_tile[k - blockIdx.x * BLOCK];
    }
    if constexpr (DEFER) {
        if (total > 0) {   // block-uniform
            __shared__ int reset_list[BLOCK];
            if (deferred) reset_list[warp_off + __popc(m & ((1u << lane) - 1u))] = i;
            __syncthreads();
            for (int j = (int)threadIdx.x; j < total; j += BLOCK) {
                const int e = reset_list[j];
                const int32_t ep = p.episode[e];
                S s{};
                E::reset(s, seed_of(p.seeds, p.seed, e), p.env_off + (uint32_t)e, (uint32_t)ep, now_t + 1, p.prm);
                p.episode[e] = ep + 1;
                E::store(p.state, p.aux, p.n, e, s);
                float o[E::OD];
                E::obs(s, o);
                if (p.obs) store_obs<E::OD, false>(p.obs, (size_t)e, o);
                for (int r = 0; r < p.world; ++r)
                    store_obs<E::OD, false>(p.peer_obs[r], ((size_t)(p.gseq & 1u) * p.world + p.rank) * (size_t)p.n + (size_t)e, o);
            }
        }
    }
    // The done LIST is not built here: a compact list needs an exclusive prefix over the CTAs, i.e. one same-address atomic
    // WITH its return value per CTA, and at 131 072 CTAs (16.7 M envs) the queue on that one address alone lasts ~57 us while
    // every CTA waits for its turn (measured: 160 us per launch without finished episodes, 217 us with).  Instead every CTA
    // leaves its count and its own compact sub-list (blk_cnt, tmp_idx) and only adds to the step's counter with a
    // fire-and-forget reduction; gymcuda_done_indices* builds the list when it is asked for (scan + scatter, kernels below).
    int32_t* count = p.done_count + (now_seq & 1);
    if (done && p.tmp_idx != nullptr) p.tmp_idx[(size_t)blockIdx.x * BLOCK + warp_off + __popc(m & ((1u << lane) - 1u))] = i;
    if (threadIdx.x == 0) {
        if (p.blk_cnt != nullptr) p.blk_cnt[blockIdx.x] = total;
        if (total > 0) atomicAdd(count, total);   // result unused: RED
        if (blockIdx.x == 0 && p.part != 2) {
            // the other counter holds the total of the PREVIOUS step launch (complete: same stream): it joins the running
            // number of finished episodes here -- one plain add per launch instead of one more atomic per CTA -- and is
            // zeroed for the next launch.  (The host adds the not-yet-folded count of the latest launch when it reports.)
            int32_t* prev = p.done_count + ((now_seq + 1) & 1);
            if (p.fold_prev) p.stats[0] += (unsigned long long)*prev;
            *prev = 0;
        }
    }
    if (p.sums != nullptr && m != 0) {   // finished episodes of this warp -> one atomic pair
        double rs = done ? (double)fin_ret : 0.0, ls = done ? (double)fin_len : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { rs += __shfl_xor_sync(0xffffffffu, rs, o); ls += __shfl_xor_sync(0xffffffffu, ls, o); }
        if (lane == 0) { atomicAdd(&p.sums[0], rs); atomicAdd(&p.sums[1], ls); }
    }
    // ---- gather signal: once the LAST block's peer stores are fenced, publish gseq in every rank's flag word
    if (p.world > 0) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned prev = atomicAdd(p.block_counter, 1u);
            if (prev == gridDim.x - 1) {
                *p.block_counter = 0u;
                __threadfence_system();
                for (int r = 0; r < p.world; ++r) *reinterpret_cast<volatile uint32_t*>(p.peer_flags[r] + p.rank) = p.gseq;
                __threadfence_system();
            }
        }
    }
}

// advances the device-resident clock after the kernels of a step (device-clock mode only)
static __global__ void clock_tick_kernel(unsigned long long* clock, unsigned long long steps) { clock[0] += steps; clock[1] += 1ull; }

// Observation gather of an env whose step is several kernels (LunarLander: contact partition + two step kernels): the
// step writes this rank's slot of its OWN gather buffer; this kernel then pushes the slot into the same slot of every
// peer's buffer with 16-byte peer stores over NVLink and, from its last block, publishes gseq in every rank's flag word
// -- the tail of step_kernel's fused path as a kernel of its own (2 MiB per peer at 65 536 landers: ~20 us after a
// 500 us step, against ~100 us for ncclAllGather).
struct PushArgs {
    const float4* src;
    float4* dst[MAX_PEERS];
    uint32_t* flags[MAX_PEERS];
    int world, rank;
    uint32_t gseq;
    size_t n4;                 // float4 elements of the slot
    unsigned* block_counter;
};
static __global__ void __launch_bounds__(256) gather_push_kernel(const PushArgs p) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < p.n4; i += (size_t)gridDim.x * 256) {
        const float4 v = p.src[i];
        for (int r = 0; r < p.world; ++r)
            if (r != p.rank) p.dst[r][i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(p.block_counter, 1u);
        if (prev == gridDim.x - 1) {
            *p.block_counter = 0u;
            __threadfence_system();
            for (int r = 0; r < p.world; ++r) *reinterpret_cast<volatile uint32_t*>(p.flags[r] + p.rank) = p.gseq;
            __threadfence_system();
        }
    }
}

// Waits (on the consumer's stream) until every rank's observations of gather step `gseq` have landed in
// this rank's buffer: one thread per peer spins on that peer's arrival flag.  Bounded: after ~2 s without
// the flag (a peer died or never launched its step) it gives up and raises `timeout_flag` (mapped host
// memory) instead of hanging the GPU.
static __global__ void gather_wait_kernel(const uint32_t* flags, int world, uint32_t gseq, int* timeout_flag) {
    const int r = threadIdx.x;
    if (r < world) {
        const volatile uint32_t* f = flags + r;
        const long long start = clock64();
        while ((int32_t)(*f - gseq) < 0) {
            __nanosleep(64);
            if (clock64() - start > 4000000000ll) { *reinterpret_cast<volatile int*>(timeout_flag) = 1 + r; break; }
        }
    }
    __threadfence_system();
}

// ---------------------------------------------------------------- fused random-policy rollout
// Threads per CTA of the rollout kernel.  64 by default (many small CTAs spread a large batch evenly).  When the whole
// batch fits in ONE wave of 16-warp CTAs and is large enough to use most SMs (CartPole's 65 536 envs: 128 CTAs on 148
// SMs), ROLLOUT_BLOCK_WAVE gives every SM that works exactly 4 warps per sub-partition and makes each SM stream one
// contiguous 512-env slice of every trajectory row: measured 147 -> 130 us per launch, the DRAM time of the launch
// (block 32 / 128 / 192 / 256 / 1024: 155 / 149 / 175 / 138 / 231 us).  At 262 144 envs the same CTAs need several
// waves at one CTA per SM and lose (Pendulum 138 -> 154 us), hence the one-wave rule in launch_rollout.
constexpr int ROLLOUT_BLOCK = 64;
constexpr int ROLLOUT_BLOCK_WAVE = 512;
constexpr int ROLLOUT_REFILL = 8;   // steps between warp-wide refills of the pre-generated reset state

template <class E>
__device__ __noinline__ typename E::S reset_cold(uint64_t seed, uint32_t gid, int32_t ep, uint64_t t, EnvParams prm) {
    typename E::S next;   // returned by value: taking the address of the caller's copy would pin it in local memory
    E::reset(next, seed, gid, (uint32_t)ep, t, prm);
    return next;
}

// E::step, or its reduced-range variant E::step<true> for envs that have one (HAS_SMALL)
template <class E, bool SMALL> struct StepSel {
    __device__ static __forceinline__ StepOut go(typename E::S& s, typename E::Act a, int32_t& sbd, uint64_t seed, uint32_t gid, uint64_t t) {
        return E::step(s, a, sbd, seed, gid, t);
    }
};
template <class E> struct StepSel<E, true> {
    __device__ static __forceinline__ StepOut go(typename E::S& s, typename E::Act a, int32_t& sbd, uint64_t seed, uint32_t gid, uint64_t t) {
        return E::template step<true>(s, a, sbd, seed, gid, t);
    }
};

// compile-time unrolled 8-step chunk: f(integral_constant<int, J>) for J = 0..7
template <int J, class F>
__device__ __forceinline__ void unroll8(F& f) {
    f(std::integral_constant<int, J>{});
    if constexpr (J < 7) unroll8<J + 1>(f);
}

// ALL_OUT: every trajectory pointer is non-null, the optional statistics / truncation bits are off and the whole
// trajectory has fewer than 2^32 rows x envs (the benchmark / learner case, checked by the host): the stores are
// unconditional, all four trajectory arrays are addressed from ONE running 32-bit row index (one IMAD.WIDE per
// address instead of 64-bit shift/add chains), and the statistics code is compiled out.
// Envs with ROLLOUT_CHUNK run the bulk of the launch as 8-step chunks aligned to the absolute step index,
// fully unrolled: action-word and reset-refill boundaries fall only on chunk starts, so the per-step loop tests,
// shifts and branches of the generic loop disappear; with HAS_SMALL (CartPole) a warp whose angles are all in
// the polynomial range runs the chunk on step<true> (no range reduction, no branches around sincos).
// SUPPLIED (with ALL_OUT): the actions of every step come from the caller's `actions_in[k][n]` (gymcuda_step_many*) instead
// of the policy stream and are not written back: the same chunked, 32-bit-indexed kernel, the eight action loads of a
// chunk issued together at its start.  An action the env rejects leaves that env unstepped for that row (observation
// unchanged, reward 0, done 0), as in step_kernel; a chunk that holds one runs step by step.
template <class E, bool AUTO_RESET, bool LIMIT, bool ALL_OUT, int BLOCK = ROLLOUT_BLOCK, bool SUPPLIED = false>
__global__ void __launch_bounds__(BLOCK) rollout_kernel(const RolloutArgs p) {
    static_assert(ALL_OUT || !SUPPLIED, "the generic variant reads actions_in through a run-time test");
    using S = typename E::S;
    using Act = typename E::Act;
    const int tix = blockIdx.x * BLOCK + threadIdx.x;
    unsigned episodes = 0;
    double fin_ret = 0.0, fin_len = 0.0;
    // Observations of 3 or 6 floats per env (Pendulum, Acrobot): a warp's 32 rows are one contiguous 384 / 768 B
    // run, but per-lane stores of 4 / 8 B pieces at a 12 / 24 B stride touch every 32 B sector 2-3 times.  A full
    // warp stages its rows in shared memory and writes the run as 16 B vectors: each sector is written once.
    constexpr bool STAGE_OBS = ALL_OUT && (E::OD == 3 || E::OD == 6);
    __shared__ __align__(16) float obs_tile[STAGE_OBS ? BLOCK * E::OD : 4];
    const unsigned lane = threadIdx.x & 31u;
    const bool staged = STAGE_OBS && p.perm == nullptr && (p.n & 3) == 0 && (tix - (int)lane + 32 <= p.n);   // warp-uniform
    if (tix < p.n) {
        const int i = p.perm ? p.perm[tix] : tix;
        S s = E::load(p.state, p.aux, p.n, i, p.prm);
        int32_t sbd = -1;
        if (E::HAS_SBD && !AUTO_RESET) sbd = p.sbd[i];
        int32_t ept = 0;
        if (LIMIT) ept = p.ep_t[i];
        const uint64_t seed = seed_of(p.seeds, p.seed, i);
        const uint32_t gid = p.env_off + (uint32_t)i;
        ActionGen<E> gen;
        bool supplied = SUPPLIED;   // the actions come from the caller (gymcuda_step_many*), not from the policy stream
        if constexpr (!ALL_OUT) supplied = p.actions_in != nullptr;
        if (!supplied) gen.init(seed, gid, p.t);
        const Act* const act_in = reinterpret_cast<const Act*>(p.actions_in);
        unsigned rejected = 0;
        const size_t n = (size_t)p.n;
        // next initial state, pre-generated: consumed at `done`, refilled for all lanes of the warp that
        // need it every REFILL steps (one Philox evaluation per warp per refill instead of per done)
        int32_t ep = 0;
        if (AUTO_RESET) ep = p.episode[i];
        constexpr bool PREGEN = AUTO_RESET && E::PREGEN_RESET;
        S next;
        if (PREGEN) next = s;
        unsigned have = 0;   // (word-sized flags: a bool that lives across the cold calls gets byte-packed with PRMTs)
        constexpr bool STATS = !ALL_OUT;
        float ret = (STATS && p.ep_ret) ? p.ep_ret[i] : 0.0f;
        uint32_t row = (uint32_t)i;   // ALL_OUT: k * n + i, addresses all four trajectory arrays
        float* const obs_base = p.obs; float* const reward_base = p.reward; uint8_t* const done_base = p.done;
        Act* const act_base = reinterpret_cast<Act*>(p.actions);

        // one env step of this lane: transition, time limit, statistics, auto-reset, trajectory stores
        // limit_tag false: the caller has established that no lane can reach the time limit in this step (the
        // episode counter still runs); for an env that never terminates by itself (Pendulum) `done` is then a
        // compile-time 0 and the whole reset path drops out of the chunk
        // rej_tag true (SUPPLIED only): this lane's action was rejected -- nothing but the stores happens for it
        auto body = [&](int k, const Act a, auto small_tag, auto limit_tag, auto rej_tag) {
            constexpr bool SMALL = decltype(small_tag)::value;
            constexpr bool CHECK_LIMIT = LIMIT && decltype(limit_tag)::value;
            constexpr bool MAY_REJECT = decltype(rej_tag)::value;
            const uint64_t t = p.t + (uint64_t)k;
            StepOut r{0.0f, 0u};
            bool trunc_only = false;
            bool stepped = true;
            if constexpr (MAY_REJECT) { stepped = E::valid(a); rejected += stepped ? 0u : 1u; }
            if (stepped) {
                r = StepSel<E, SMALL>::go(s, a, sbd, seed, gid, t);
                if (LIMIT) ept += 1;
                if (CHECK_LIMIT) { if (ept >= p.limit && !r.done) { r.done = 1u; trunc_only = true; } }
            }
            if (STATS) ret += r.reward;
            if (r.done) {
                episodes += 1;
                if (STATS && p.sums) { fin_ret += (double)ret; fin_len += (double)ept; }
                if (STATS) ret = 0.0f;
                if (AUTO_RESET) {
                    if (PREGEN) {
                        if (!have) next = reset_cold<E>(seed, gid, ep, t + 1, p.prm);   // second done before the refill: rare
                        s = next;
                        have = 0;
                    } else {
                        E::reset(s, seed, gid, (uint32_t)ep, t + 1, p.prm);
                    }
                    ep += 1;
                    sbd = -1;
                    ept = 0;
                }
            }
            if (ALL_OUT) {
                float o[E::OD];
                E::obs(s, o);
                if (STAGE_OBS && staged) {
                    float* tile = obs_tile + (threadIdx.x - lane) * E::OD;
#pragma unroll
                    for (int j = 0; j < E::OD; ++j) tile[lane * E::OD + j] = o[j];
                    __syncwarp();
                    float4* dst = reinterpret_cast<float4*>(obs_base + (size_t)(row - lane) * E::OD);
#pragma unroll
                    for (int v = (int)lane; v < 8 * E::OD; v += 32) __stcs(dst + v, reinterpret_cast<const float4*>(tile)[v]);
                    __syncwarp();
                } else {
                    store_obs<E::OD, true>(obs_base, (size_t)row, o);
                }
                __stcs(reward_base + row, r.reward);
                __stcs(done_base + row, (uint8_t)r.done);
                if constexpr (!SUPPLIED) __stcs(act_base + row, a);
                row += (uint32_t)p.n;
            } else {
                const size_t idx = (size_t)k * n + (size_t)i;
                if (p.obs) {
                    float o[E::OD];
                    E::obs(s, o);
                    store_obs<E::OD, true>(p.obs, idx, o);
                }
                if (p.reward) __stcs(p.reward + idx, r.reward);
                if (p.done) __stcs(p.done + idx, (p.done_bits && trunc_only) ? (uint8_t)2 : (uint8_t)r.done);
                if (p.actions) ActIO<E>::store(p.actions, idx, a);
            }
        };

        int k = 0;
        if constexpr (ALL_OUT && E::ROLLOUT_CHUNK) {
            // head: single steps until the absolute step index is a multiple of 8
            int head = (int)((8u - ((uint32_t)p.t & 7u)) & 7u);
            if (head > p.k_steps) head = p.k_steps;
            constexpr bool MAY_REJECT = SUPPLIED && E::REJECT_INVALID;
            using RejTag = std::integral_constant<bool, MAY_REJECT>;
            for (; k < head; ++k) {
                if constexpr (SUPPLIED) body(k, __ldcs(act_in + row), std::false_type{}, std::true_type{}, RejTag{});
                else body(k, gen.next(seed, gid, p.t + (uint64_t)k), std::false_type{}, std::true_type{}, std::false_type{});
            }
            // SUPPLIED: the actions of the NEXT chunk are loaded while this one is stepped (two register sets of eight)
            Act ahead[SUPPLIED ? 8 : 1];
            if constexpr (SUPPLIED) {
                if (k + 8 <= p.k_steps) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) ahead[j] = __ldcs(act_in + (row + (uint32_t)j * (uint32_t)p.n));
                }
            }
#pragma unroll 1
            for (; k + 8 <= p.k_steps; k += 8) {
                const uint64_t tc = p.t + (uint64_t)k;
                if (PREGEN && !have) {
                    E::reset(next, seed, gid, (uint32_t)ep, tc, p.prm);
                    have = 1;
                }
                Act acts[SUPPLIED ? 8 : 1];
                if constexpr (SUPPLIED) {
                    bool ok = true;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        acts[j] = ahead[j];
                        if (MAY_REJECT) ok = ok && E::valid(acts[j]);
                    }
                    if (k + 16 <= p.k_steps) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) ahead[j] = __ldcs(act_in + (row + (uint32_t)(8 + j) * (uint32_t)p.n));
                    }
                    if constexpr (MAY_REJECT) {
                        if (!__all_sync(__activemask(), ok)) {   // rare: a rejected action somewhere in the warp's chunk
                            auto f = [&](auto jc) { constexpr int J = decltype(jc)::value; body(k + J, acts[J], std::false_type{}, std::true_type{}, std::true_type{}); };
                            unroll8<0>(f);
                            continue;
                        }
                    }
                }
                // the action of step tc + J: the caller's, or draw J of the chunk
                auto act_of = [&](auto jc) -> Act {
                    constexpr int J = decltype(jc)::value;
                    if constexpr (SUPPLIED) return acts[J]; else return gen.template at<J>(seed, gid, tc);
                };
                if constexpr (E::HAS_SMALL && AUTO_RESET) {
                    if (__all_sync(__activemask(), E::small_ok(s))) {
                        if constexpr (LIMIT) {
                            // no lane within 8 steps of the time limit (a reset only lowers the counter): the chunk
                            // runs without the per-step limit test
                            if (__all_sync(__activemask(), ept + 8 < p.limit)) {
                                auto f = [&](auto jc) { constexpr int J = decltype(jc)::value; body(k + J, act_of(jc), std::true_type{}, std::false_type{}, std::false_type{}); };
                                unroll8<0>(f);
                                continue;
                            }
                        }
                        auto f = [&](auto jc) { constexpr int J = decltype(jc)::value; body(k + J, act_of(jc), std::true_type{}, std::true_type{}, std::false_type{}); };
                        unroll8<0>(f);
                        continue;
                    }
                }
                auto f = [&](auto jc) { constexpr int J = decltype(jc)::value; body(k + J, act_of(jc), std::false_type{}, std::true_type{}, std::false_type{}); };
                unroll8<0>(f);
            }
        }
        // generic loop: everything for the other variants, the tail (< 8 steps) of a chunked launch
        for (int k0 = k; k < p.k_steps; ++k) {
            if (PREGEN && ((unsigned)(k - k0) & (unsigned)(ROLLOUT_REFILL - 1)) == 0u && !have) {
                E::reset(next, seed, gid, (uint32_t)ep, p.t + (uint64_t)k, p.prm);
                have = 1;
            }
            Act a;
            if constexpr (!ALL_OUT) {
                if (supplied) {
                    a = ActIO<E>::load_at(p.actions_in, (size_t)k * n + (size_t)i);
                    if (E::REJECT_INVALID && !E::valid(a)) {
                        // like step_kernel: the env is not stepped; this row of the trajectory holds its unchanged observation
                        rejected += 1;
                        const size_t idx = (size_t)k * n + (size_t)i;
                        if (p.obs) { float o[E::OD]; E::obs(s, o); store_obs<E::OD, true>(p.obs, idx, o); }
                        if (p.reward) __stcs(p.reward + idx, 0.0f);
                        if (p.done) __stcs(p.done + idx, (uint8_t)0);
                        continue;
                    }
                } else {
                    a = gen.next(seed, gid, p.t + (uint64_t)k);
                }
            } else if constexpr (SUPPLIED) {
                a = __ldcs(act_in + row);
            } else {
                a = gen.next(seed, gid, p.t + (uint64_t)k);
            }
            using RejTag = std::integral_constant<bool, SUPPLIED && E::REJECT_INVALID>;
            if constexpr (ALL_OUT && E::HAS_SMALL && !E::ROLLOUT_CHUNK) {
                // envs too large to unroll (Acrobot): the same warp vote, per step, picks the reduced-range step
                if (__all_sync(__activemask(), E::small_ok(s))) { body(k, a, std::true_type{}, std::true_type{}, RejTag{}); continue; }
            }
            body(k, a, std::false_type{}, std::true_type{}, RejTag{});
        }
        if (AUTO_RESET) p.episode[i] = ep;
        if (STATS && p.ep_ret) p.ep_ret[i] = ret;
        E::store(p.state, p.aux, p.n, i, s);
        if (E::HAS_SBD && !AUTO_RESET) p.sbd[i] = sbd;
        if (LIMIT) p.ep_t[i] = ept;
        if constexpr (!ALL_OUT || SUPPLIED) {
            if (rejected) {
                atomicAdd(&p.stats[1], (unsigned long long)rejected);
                *reinterpret_cast<volatile int*>(p.host_invalid) = 1;
                __threadfence_system();
            }
        }
    }
    // episodes finished: warp reduce, one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) episodes += __shfl_xor_sync(0xffffffffu, episodes, o);
    if ((threadIdx.x & 31) == 0 && episodes) atomicAdd(&p.stats[0], (unsigned long long)episodes);
    if (!ALL_OUT && p.sums != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { fin_ret += __shfl_xor_sync(0xffffffffu, fin_ret, o); fin_len += __shfl_xor_sync(0xffffffffu, fin_len, o); }
        if ((threadIdx.x & 31) == 0 && episodes) { atomicAdd(&p.sums[0], fin_ret); atomicAdd(&p.sums[1], fin_len); }
    }
}

// ---------------------------------------------------------------- reset / observe / ctor
template <class E>
__global__ void __launch_bounds__(128) reset_kernel(const ResetArgs p) {
    using S = typename E::S;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    S s = E::load(p.state, p.aux, p.n, i, p.prm);
    if (p.mask == nullptr || p.mask[i]) {
        const int32_t ep = p.episode[i];
        E::reset(s, seed_of(p.seeds, p.seed, i), p.env_off + (uint32_t)i, (uint32_t)ep, p.t, p.prm);
        E::store(p.state, p.aux, p.n, i, s);
        p.episode[i] = ep + 1;
        p.sbd[i] = -1;     // CartPoleEnv.cs:64
        p.ep_t[i] = 0;
        if (p.ep_ret) p.ep_ret[i] = 0.0f;
    }
    if (p.obs) {
        float o[E::OD];
        E::obs(s, o);
        store_obs<E::OD, false>(p.obs, (size_t)i, o);
    }
}

// ---------------------------------------------------------------- contact partition (LunarLander)
// A lander near the ground (one whose broad phase holds a contact pair) runs the narrow phase and, when it touches, a
// several times longer solver path than one in free flight; a warp runs as long as its slowest lane and a launch as
// long as its slowest warp.  A STABLE partition of the env ids by "a contact pair exists" (warp ballot + block counts +
// scan, the machinery of the done compaction) splits the batch into two classes that are stepped by two different
// kernels, concurrently: the free-flight class by step_kernel<LunarLanderT<C, false>> (no narrow phase, no contact rows,
// no contact slots loaded or stored: fewer registers, more resident warps), the other by <C, true>.  A lander cannot
// change class inside a step -- pairs are created by FindNewContacts at the END of World.Step -- so the class is known
// from the stored state.  Stability keeps each class in ascending env order, so the field-major loads of a warp still
// fall in a handful of neighbouring sectors.  Results are unaffected: which thread steps an env is invisible to it.
constexpr int PART_BLOCK = 256;
constexpr int LUNAR_PAIRS_WORD = 26;   // aux word of lunar::Lander::pairs[0] (lunar.cuh static_asserts it)

__device__ __forceinline__ bool lander_in_contact(const int32_t* aux, int n, int i) {
    return aux[(size_t)LUNAR_PAIRS_WORD * (size_t)n + i] != -1;   // first word of the broad-phase pair list: 0xffffffff = no contact exists
}

static __global__ void __launch_bounds__(PART_BLOCK) partition_count_kernel(const int32_t* aux, int n, int32_t* block_free) {
    const int i = blockIdx.x * PART_BLOCK + threadIdx.x;
    const bool free_flight = i < n && !lander_in_contact(aux, n, i);
    const int c = __syncthreads_count(free_flight);
    if (threadIdx.x == 0) block_free[blockIdx.x] = c;
}

// single block: exclusive scan of the per-block counts, in place; block_free[nb] = total
static __global__ void __launch_bounds__(1024) partition_scan_kernel(int32_t* block_free, int nb) {
    __shared__ int carry;
    __shared__ int warp_sum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int j = base + (int)threadIdx.x;
        const int v = j < nb ? block_free[j] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= (unsigned)o) x += y; }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warp_sum[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= (unsigned)o) w += y; }
            warp_sum[threadIdx.x] = w;
        }
        __syncthreads();
        const int incl = x + ((threadIdx.x >> 5) ? warp_sum[(threadIdx.x >> 5) - 1] : 0);
        if (j < nb) block_free[j] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_free[nb] = carry;
}

static __global__ void __launch_bounds__(PART_BLOCK) partition_scatter_kernel(const int32_t* aux, int n, const int32_t* block_free, int nb, int32_t* perm) {
    __shared__ int warp_free[PART_BLOCK / 32];
    const int i = blockIdx.x * PART_BLOCK + threadIdx.x;
    const bool in = i < n;
    const bool free_flight = in && !lander_in_contact(aux, n, i);
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, free_flight);
    if (lane == 0) warp_free[warp] = __popc(m);
    __syncthreads();
    int before = 0;   // free-flight envs of this block before this warp
    for (int w = 0; w < (int)warp; ++w) before += warp_free[w];
    const int free_rank = before + __popc(m & ((1u << lane) - 1u));
    const int local = (int)threadIdx.x;
    if (in) {
        const int free_base = block_free[blockIdx.x], total_free = block_free[nb];
        if (free_flight) perm[free_base + free_rank] = i;
        else perm[total_free + (blockIdx.x * PART_BLOCK - free_base) + (local - free_rank)] = i;
    }
}

// The done list on demand (gymcuda_done_indices*): blk_off = exclusive scan of the per-CTA counts of the last step launch
// (partition_scan_kernel, in place: blk_off[nb] = total), then one warp per CTA copies its sub-list to its place.
static __global__ void __launch_bounds__(256) done_list_scatter_kernel(const int32_t* blk_off, const int32_t* tmp_idx, int nb, int block, int32_t* done_idx) {
    const int b = blockIdx.x * 8 + (int)(threadIdx.x >> 5);
    if (b >= nb) return;
    const int base = blk_off[b], cnt = blk_off[b + 1] - base;
    for (int k = (int)(threadIdx.x & 31); k < cnt; k += 32) done_idx[base + k] = tmp_idx[(size_t)b * block + k];
}

// ActionSpace.Sample() for every env at step index t: the same draws the rollout kernel consumes, so
// sample -> step loops reproduce gymcuda_rollout_random.  With a mask (Discrete only, Discrete.cs:19-25):
// uniform over the entries equal to 1, `Start` (= 0) when none is.
struct SampleArgs { const int32_t* seeds; const uint8_t* mask; void* out; int n; uint32_t env_off; uint64_t seed; uint64_t t; };

template <class E>
__global__ void __launch_bounds__(128) sample_kernel(const SampleArgs p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const uint64_t seed = seed_of(p.seeds, p.seed, i);
    const uint32_t gid = p.env_off + (uint32_t)i;
    ActionGen<E> gen;
    gen.init(seed, gid, p.t);
    typename E::Act a = gen.next(seed, gid, p.t);
    if (E::ACTN > 0 && p.mask != nullptr) {
        const uint8_t* m = p.mask + (size_t)i * (E::ACTN > 0 ? E::ACTN : 1);
        int valid = 0;
        for (int k = 0; k < E::ACTN; ++k) valid += (m[k] == 1);
        int pick = 0;
        if (valid > 0) {
            // an independent full word of the ACTION stream (sub-block 1) picks the j-th valid entry
            const Block b = draw(seed, gid, p.t, STREAM_ACTION, 1);
            int j = (int)__umulhi(b.w0, (uint32_t)valid);
            for (int k = 0; k < E::ACTN; ++k) if (m[k] == 1) { if (j == 0) { pick = k; break; } --j; }
        }
        a = ActCast<typename E::Act>::from_int(pick);
    }
    reinterpret_cast<typename E::Act*>(p.out)[i] = a;
}

template <class E>
__global__ void __launch_bounds__(128) ctor_kernel(const ResetArgs p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    E::ctor(p.state, p.aux, p.n, i, seed_of(p.seeds, p.seed, i), p.env_off + (uint32_t)i);
}

}  // namespace gymcuda

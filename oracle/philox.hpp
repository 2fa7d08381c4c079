// TEST INFRASTRUCTURE ONLY (see oracle/README.md). Nothing under gym.net_b200/ may include this file.
//
// Host restatement of the engine's counter-based RNG ("gymcuda RNG spec v1").
//
// Why our own stream: the reference draws from NumSharp.Lite 0.1.12's NumPyRandom
// (reference call sites: src/Gym.Environments/Envs/Classic/CartPoleEnv.cs:65,
// src/Gym.Environments/Envs/Aether/LunarLanderEnv.cs:496,507,611-612,
// src/Gym/Spaces/Discrete.cs:27, src/Gym/Spaces/Box.cs:84).  That package is an
// un-vendored NuGet dependency (src/Gym/Gym.csproj:29) and is absent from this
// image, so its stream cannot be reproduced: RNG STREAM PARITY IS UNPINNED.
// The *distributions* (uniform(low,high), randint(0,N)) follow the call sites.
//
// Generator: Philox4x32-10 (Salmon et al., SC'11), pinned by the three Random123
// known-answer vectors in tests/test_oracle_philox.py.
//
// Addressing (one 128-bit block per call):
//     key = (seed_lo, global_env_id)
//     ctr = (index_lo, index_hi, stream | sub << 8, seed_hi)
// stream 0 = RESET    index = episode ordinal of the env (number of resets it has had)
// stream 1 = ACTION   index = block number of the random-policy draw (see action_*)
// stream 2 = DYNAMICS index = t (LunarLander's two per-step dispersion draws)
// stream 3 = CTOR     index = 0 (LunarLander's wind / turbulence phase, drawn in its constructor)
#pragma once
#include <cstdint>

namespace oracle {

enum : uint32_t { STREAM_RESET = 0, STREAM_ACTION = 1, STREAM_DYNAMICS = 2, STREAM_CTOR = 3 };

struct Block { uint32_t w[4]; };

inline Block philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                           uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return Block{{c0, c1, c2, c3}};
}

inline Block draw(uint64_t seed, uint32_t env_id, uint64_t index, uint32_t stream, uint32_t sub = 0) {
    return philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), stream | (sub << 8),
                         (uint32_t)(seed >> 32), (uint32_t)seed, env_id);
}

// [0,1) with 24 bits, exact in fp32.
inline float u01(uint32_t w) { return (float)(w >> 8) * 0x1p-24f; }

// low + (high-low)*u, separate multiply and add in fp32 (never fused).
inline float uniformf(float lo, float hi, uint32_t w) {
    volatile float span = hi - lo;
    volatile float prod = span * u01(w);
    return lo + prod;
}

// Random policy, reference semantics Discrete.Sample = Start + randint(0, N)
// (src/Gym/Spaces/Discrete.cs:27) and Box.Sample bounded = uniform(low, high)
// (src/Gym/Spaces/Box.cs:84).
inline int action_discrete2(uint64_t seed, uint32_t env, uint64_t t) {   // 128 draws per block
    Block b = draw(seed, env, t >> 7, STREAM_ACTION);
    return (int)((b.w[(t >> 5) & 3] >> (t & 31)) & 1u);
}
inline int action_discrete4(uint64_t seed, uint32_t env, uint64_t t) {   // 64 draws per block
    Block b = draw(seed, env, t >> 6, STREAM_ACTION);
    uint32_t i = (uint32_t)(t & 63);
    return (int)((b.w[i >> 4] >> (2 * (i & 15))) & 3u);
}
inline int action_discrete3(uint64_t seed, uint32_t env, uint64_t t) {   // 4 draws per block
    Block b = draw(seed, env, t >> 2, STREAM_ACTION);
    return (int)(((uint64_t)b.w[t & 3] * 3u) >> 32);
}
inline float action_box1(uint64_t seed, uint32_t env, uint64_t t, float lo, float hi) {  // 4 per block
    Block b = draw(seed, env, t >> 2, STREAM_ACTION);
    return uniformf(lo, hi, b.w[t & 3]);
}
inline void action_box2(uint64_t seed, uint32_t env, uint64_t t, float lo, float hi, float out[2]) {
    Block b = draw(seed, env, t >> 1, STREAM_ACTION);
    out[0] = uniformf(lo, hi, b.w[2 * (t & 1)]);
    out[1] = uniformf(lo, hi, b.w[2 * (t & 1) + 1]);
}

}  // namespace oracle

// TEST INFRASTRUCTURE: a stand-in for <cuda_runtime.h> that lets g++ compile the DEVICE headers of the engine
// (gym.net_b200/csrc/{detmath,philox,env_classic,lunar,lunar_core}.cuh) as plain host C++ -- see hostsim.cpp.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define GYMCUDA_HOSTSIM 1
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __constant__ static const
#define __restrict__

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 v; v.x = x; v.y = y; return v; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
static inline float __fmul_rn(float a, float b) { return a * b; }   // compiled with -ffp-contract=off: never fused
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }

// ---- what kernels.cuh needs on top.  Two execution modes (hostsim.cpp):
//   * one thread at a time, a "warp" of ONE lane: votes are per-thread, shuffles see no other lane;
//   * simt::launch (../simt.hpp): the threads of a CTA run as fibers and the primitives below are real rendezvous.
#include <cstddef>
#include "../simt.hpp"
struct hostsim_dim3 { unsigned x, y, z; };
static hostsim_dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
#define __global__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
static inline void __threadfence_system() {}
static inline void __nanosleep(unsigned) {}
static inline long long clock64() { static long long ticks = 0; return ticks += 1000000000ll; }   // every look at the clock is 1e9 cycles later
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline void __syncthreads() { if (simt::active()) simt::wait(simt::WAIT_BLOCK, 0, 0, 0, 0); }
static inline int __syncthreads_count(int p) { return simt::active() ? (int)simt::wait(simt::WAIT_BLOCK, 0, 0, p ? 1 : 0, 0) : (p ? 1 : 0); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { if (simt::active()) simt::wait(simt::WAIT_WARP, simt::OP_SYNCWARP, mask, 0, 0); }
static inline unsigned __activemask() { return simt::active() ? (unsigned)simt::wait(simt::WAIT_AMASK, 0, 0, 0, 0) : 1u; }
static inline unsigned __ballot_sync(unsigned mask, int p) {
    return simt::active() ? (unsigned)simt::wait(simt::WAIT_WARP, simt::OP_BALLOT, mask, p ? 1 : 0, 0) : (p ? 1u : 0u);
}
static inline int __all_sync(unsigned mask, int p) {
    return simt::active() ? (int)simt::wait(simt::WAIT_WARP, simt::OP_ALL, mask, p ? 1 : 0, 0) : (p ? 1 : 0);
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int o) {
    if (!simt::active()) return T(0);   // one-lane warp: lanes that do not exist contribute nothing
    return simt::from_bits<T>(simt::wait(simt::WAIT_WARP, simt::OP_SHFL_XOR, mask, simt::to_bits(v), o));
}
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src) {
    if (!simt::active()) return v;      // one-lane warp: lane 0 is this lane
    return simt::from_bits<T>(simt::wait(simt::WAIT_WARP, simt::OP_SHFL_IDX, mask, simt::to_bits(v), src));
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, int o) {
    if (!simt::active()) return v;
    return simt::from_bits<T>(simt::wait(simt::WAIT_WARP, simt::OP_SHFL_UP, mask, simt::to_bits(v), o));
}
template <class T, class U> static inline T atomicAdd(T* p, U v) { T old = *p; *p = (T)(old + (T)v); return old; }

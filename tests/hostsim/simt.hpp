// TEST INFRASTRUCTURE ONLY.  A small SIMT executor for the host build of the device code (hostsim.cpp): the threads of
// one CTA run as cooperative fibers (ucontext), and the warp / block primitives the kernels use -- __ballot_sync,
// __all_sync, __shfl_xor_sync, __shfl_up_sync, __syncwarp, __activemask, __syncthreads, __syncthreads_count -- are
// rendezvous points between them with the CUDA semantics:
//   * a *_sync(mask, ...) completes when every lane of `mask` that has not exited has arrived with the same primitive
//     and mask; lanes get the values the others contributed (a shuffle from a lane outside the group returns the
//     caller's own value);
//   * __activemask() is the set of lanes that sit at an __activemask() when every other live lane of the warp is
//     blocked somewhere else -- i.e. the lanes that are converged at that instruction;
//   * __syncthreads() completes when every thread of the CTA that has not exited has arrived.
// CTAs run one after another (kernels here never communicate between CTAs except through atomics, which are plain
// operations in a single OS thread), so `static` stands in for __shared__.
// Outside simt::launch() (hostsim's one-thread-at-a-time mode) the primitives keep their one-lane meaning.
#pragma once
#include <sys/mman.h>
#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace simt {

enum State { RUNNING = 0, WAIT_WARP, WAIT_AMASK, WAIT_BLOCK, DONE };
enum Op { OP_BALLOT = 1, OP_ALL, OP_SHFL_XOR, OP_SHFL_UP, OP_SYNCWARP, OP_SHFL_IDX };

struct Fiber {
    ucontext_t ctx;
    int state = RUNNING;
    int op = 0;
    unsigned mask = 0;
    int arg = 0;         // shuffle distance
    uint64_t val = 0;    // contributed value (predicate or raw bits)
    uint64_t res = 0;    // delivered result
};

struct Cta {
    std::vector<Fiber> f;
    ucontext_t sched;
    std::vector<int> runq;
    int cur = -1;
    int nthreads = 0;
    const std::function<void()>* body = nullptr;
};

// Which runnable thread goes next: 0 = oldest first (lanes proceed in ascending order), 1 = newest first (descending),
// 2 = pseudo-random.  A kernel that is correct under CUDA's rules gives the same results under every order; one that
// leans on an ordering between lanes that only a missing __syncwarp / __syncthreads would provide does not.
static int g_policy = 0;
static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
static inline uint64_t next_random() { g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17; return g_rng; }

static Cta* g_cta = nullptr;                    // non-null while a launch is running
static void (*g_set_thread)(unsigned) = nullptr;   // installs threadIdx.x before a fiber resumes

constexpr size_t STACK_BYTES = 256 * 1024;

static inline bool active() { return g_cta != nullptr; }

static void trampoline() {
    Cta* c = g_cta;
    (*c->body)();
    c->f[c->cur].state = DONE;
    // returning continues at uc_link (the scheduler)
}

[[noreturn]] static void deadlock(const Cta& c) {
    std::fprintf(stderr, "hostsim simt: deadlock in a CTA of %d threads\n", c.nthreads);
    for (int i = 0; i < c.nthreads; ++i)
        if (c.f[i].state != DONE) std::fprintf(stderr, "  thread %d: state %d op %d mask %08x\n", i, c.f[i].state, c.f[i].op, c.f[i].mask);
    std::abort();
}

// blocks the calling fiber until the scheduler has resolved its rendezvous
static inline uint64_t wait(int state, int op, unsigned mask, uint64_t val, int arg) {
    Cta* c = g_cta;
    Fiber& me = c->f[c->cur];
    me.state = state; me.op = op; me.mask = mask; me.val = val; me.arg = arg;
    swapcontext(&me.ctx, &c->sched);
    return me.res;
}

static void release(Cta& c, int id, uint64_t res) {
    c.f[id].res = res;
    c.f[id].state = RUNNING;
    c.runq.push_back(id);
}

static void resolve_warp(Cta& c, int w) {
    const int lo = w * 32, hi = lo + 32 < c.nthreads ? lo + 32 : c.nthreads;
    unsigned live = 0;
    for (int i = lo; i < hi; ++i) if (c.f[i].state != DONE) live |= 1u << (i - lo);
    // explicit-mask primitives: a group completes as soon as all of its live lanes have arrived
    for (int i = lo; i < hi; ++i) {
        Fiber& a = c.f[i];
        if (a.state != WAIT_WARP) continue;
        const unsigned group = a.mask & live;
        bool ready = true;
        for (int j = lo; j < hi && ready; ++j) {
            if (!((group >> (j - lo)) & 1u)) continue;
            const Fiber& b = c.f[j];
            // a lane still on its way, or sitting at another primitive of a sub-group (it joins this one later)
            if (b.state != WAIT_WARP || b.op != a.op || b.mask != a.mask) { ready = false; break; }
        }
        if (!ready) continue;
        unsigned ballot = 0;
        for (int j = lo; j < hi; ++j) if (((group >> (j - lo)) & 1u) && c.f[j].val) ballot |= 1u << (j - lo);
        const int op = a.op, arg = a.arg;
        uint64_t vals[32];
        for (int j = lo; j < hi; ++j) vals[j - lo] = c.f[j].val;
        for (int j = lo; j < hi; ++j) {
            if (!((group >> (j - lo)) & 1u)) continue;
            const int lane = j - lo;
            uint64_t r = 0;
            switch (op) {
                case OP_BALLOT: r = ballot; break;
                case OP_ALL: r = ballot == group ? 1u : 0u; break;
                case OP_SHFL_XOR: { const int src = lane ^ arg; r = (src < 32 && ((group >> src) & 1u)) ? vals[src] : vals[lane]; break; }
                case OP_SHFL_UP: { const int src = lane - arg; r = (src >= 0 && ((group >> src) & 1u)) ? vals[src] : vals[lane]; break; }
                case OP_SHFL_IDX: { const int src = arg & 31; r = ((group >> src) & 1u) ? vals[src] : vals[lane]; break; }
                case OP_SYNCWARP: r = 0; break;
            }
            release(c, j, r);
        }
    }
    // __activemask(): the lanes waiting at it once nobody else in the warp can still get there
    unsigned at_amask = 0;
    bool all_blocked = true;
    for (int i = lo; i < hi; ++i) {
        if (c.f[i].state == RUNNING) all_blocked = false;
        if (c.f[i].state == WAIT_AMASK) at_amask |= 1u << (i - lo);
    }
    if (all_blocked && at_amask)
        for (int i = lo; i < hi; ++i) if ((at_amask >> (i - lo)) & 1u) release(c, i, at_amask);
}

static void resolve_block(Cta& c) {
    int waiting = 0, done = 0; uint64_t count = 0;
    for (const Fiber& x : c.f) { if (x.state == WAIT_BLOCK) { waiting += 1; count += x.val ? 1 : 0; } else if (x.state == DONE) done += 1; }
    if (waiting > 0 && waiting + done == c.nthreads)
        for (int i = 0; i < c.nthreads; ++i) if (c.f[i].state == WAIT_BLOCK) release(c, i, count);
}

// runs body() once per thread of one CTA; body reads threadIdx / blockIdx as usual
static void run_cta(int nthreads, const std::function<void()>& body, char* stacks) {
    Cta c;
    c.f.resize(nthreads);
    c.nthreads = nthreads;
    c.body = &body;
    for (int i = 0; i < nthreads; ++i) {
        getcontext(&c.f[i].ctx);
        c.f[i].ctx.uc_stack.ss_sp = stacks + (size_t)i * STACK_BYTES;
        c.f[i].ctx.uc_stack.ss_size = STACK_BYTES;
        c.f[i].ctx.uc_link = &c.sched;
        makecontext(&c.f[i].ctx, trampoline, 0);
        c.runq.push_back(i);
    }
    g_cta = &c;
    for (;;) {
        if (c.runq.empty()) {
            for (int w = 0; w * 32 < nthreads; ++w) resolve_warp(c, w);
            resolve_block(c);
            if (c.runq.empty()) {
                bool all_done = true;
                for (const Fiber& x : c.f) if (x.state != DONE) all_done = false;
                if (all_done) break;
                deadlock(c);
            }
        }
        size_t pick = 0;
        if (g_policy == 1) pick = c.runq.size() - 1;
        else if (g_policy == 2) pick = (size_t)(next_random() % c.runq.size());
        const int id = c.runq[pick];
        c.runq.erase(c.runq.begin() + (long)pick);
        c.cur = id;
        g_set_thread((unsigned)id);
        swapcontext(&c.sched, &c.f[id].ctx);
        // the fiber blocked or finished: its warp (or the CTA) may have become resolvable
        resolve_warp(c, id / 32);
        if (c.f[id].state == WAIT_BLOCK || c.f[id].state == DONE) resolve_block(c);
    }
    g_cta = nullptr;
}

// grid x block launch; set_block installs blockIdx.x, set_thread installs threadIdx.x
static void launch(int grid, int block, void (*set_block)(unsigned, unsigned, unsigned), void (*set_thread)(unsigned),
                   const std::function<void()>& body) {
    char* stacks = (char*)mmap(nullptr, (size_t)block * STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (stacks == (char*)MAP_FAILED) { std::perror("hostsim simt: mmap"); std::abort(); }
    g_set_thread = set_thread;
    for (int b = 0; b < grid; ++b) {
        set_block((unsigned)b, (unsigned)grid, (unsigned)block);
        run_cta(block, body, stacks);
    }
    munmap(stacks, (size_t)block * STACK_BYTES);
}

template <class T> static inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, "payload"); std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> static inline T from_bits(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }

}  // namespace simt

// warp_emu.h -- TEST INFRASTRUCTURE: runs device code written for one warp on the CPU.
//
// The product's warp-cooperative searches (raydar_b200/csrc/rdr_fused.cuh) are device-only: 32 lanes exchange data with
// __shfl_sync / __ballot_sync, meet at __syncwarp and fold winners with atomicMin on shared memory.  Here the 32 lanes of
// a warp are 32 fibers (ucontext) of one OS thread; every *_sync intrinsic is a rendezvous of all 32 (the code under test
// only uses full masks at converged points), between two rendezvous a lane runs alone, so divergent per-lane loops need
// no special treatment and atomics are plain read-modify-writes.  A lane that returns while others still wait, or a
// rendezvous that not all lanes reach, is reported as a deadlock instead of hanging.
//
// Include AFTER the product's host-compilable headers and BEFORE rdr_fused.cuh (with RDR_WARP_EMU defined).
#pragma once

#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

namespace warp_emu {

constexpr int LANES = 32;
constexpr size_t STACK_BYTES = 256u << 10;

struct Warp {
    ucontext_t main_ctx, ctx[LANES];
    ucontext_t *sched = nullptr;        // the context lanes yield to (this warp's main_ctx, or the grid's)
    std::vector<char> stacks;
    bool done[LANES];
    int cur = 0;
    unsigned tid_base = 0;              // threadIdx.x of lane 0 (warps of one emulated CTA: 0, 32, 64, ...)
    int warp_index = 0;
    uint32_t arrived = 0, gen = 0;
    uint32_t slot[2][LANES];
    uint64_t progress = 0;              // bumped by every rendezvous arrival and every lane exit (deadlock detection)
    std::function<void(int)> body;
    bool deadlock = false;
};

inline Warp *&current() { static thread_local Warp *w = nullptr; return w; }
inline unsigned lane() { return (unsigned)current()->cur; }

inline void yield_lane() { Warp *w = current(); swapcontext(&w->ctx[w->cur], w->sched); }

// all 32 lanes meet here
inline void rendezvous()
{
    Warp *w = current();
    const uint32_t g = w->gen;
    ++w->progress;
    if (++w->arrived == (uint32_t)LANES) { w->arrived = 0; ++w->gen; return; }
    while (w->gen == g) yield_lane();
}

inline uint32_t exchange(uint32_t v, uint32_t src)
{
    Warp *w = current();
    const uint32_t b = w->gen & 1u;
    w->slot[b][w->cur] = v;
    rendezvous();
    return w->slot[b][src & 31u];
}

inline uint32_t gather_mask(bool pred)
{
    Warp *w = current();
    const uint32_t b = w->gen & 1u;
    w->slot[b][w->cur] = pred ? 1u : 0u;
    rendezvous();
    uint32_t m = 0u;
    for (int i = 0; i < LANES; ++i) m |= w->slot[b][i] << i;
    return m;
}

inline void lane_entry()
{
    Warp *w = current();
    w->body(w->cur);
    w->done[w->cur] = true;
    ++w->progress;
    swapcontext(&w->ctx[w->cur], w->sched);          // never resumed
}

inline void prepare_lanes(Warp &w)
{
    if (w.stacks.empty()) w.stacks.resize(STACK_BYTES * LANES);
    w.arrived = 0; w.gen = 0; w.progress = 0; w.deadlock = false;
    for (int i = 0; i < LANES; ++i) {
        w.done[i] = false;
        getcontext(&w.ctx[i]);
        w.ctx[i].uc_stack.ss_sp = w.stacks.data() + STACK_BYTES * i;
        w.ctx[i].uc_stack.ss_size = STACK_BYTES;
        w.ctx[i].uc_link = w.sched;
        makecontext(&w.ctx[i], (void (*)())lane_entry, 0);
    }
}

// runs body(lane) for lanes 0..31 as one warp; false = deadlock (a lane exited or stalled while others wait)
inline bool run_warp(Warp &w, std::function<void(int)> body)
{
    w.sched = &w.main_ctx;
    w.body = std::move(body);
    Warp *prev = current();
    current() = &w;
    prepare_lanes(w);
    for (;;) {
        bool all_done = true;
        const uint64_t before = w.progress;
        for (int i = 0; i < LANES; ++i) {
            if (w.done[i]) continue;
            all_done = false;
            w.cur = i;
            swapcontext(&w.main_ctx, &w.ctx[i]);
        }
        if (all_done) break;
        if (w.progress == before) { w.deadlock = true; break; }     // a full pass without any arrival or exit
    }
    current() = prev;
    return !w.deadlock;
}

// Several warps of one emulated CTA (threadIdx.x = 32 * warp + lane), interleaved lane by lane on one OS thread: what a
// kernel whose warps talk through global memory (an atomic work counter, a polled progress word) needs.  `order` permutes
// the round-robin so that tests can try different interleavings.  false = a warp deadlocked or max_passes was reached
// (a protocol that never terminates).
inline bool run_grid(std::vector<Warp> &warps, std::function<void(int, int)> body, const std::vector<int> &order = {},
                     uint64_t max_passes = 50u * 1000u * 1000u)
{
    ucontext_t grid_ctx;
    Warp *prev = current();
    for (size_t k = 0; k < warps.size(); ++k) {
        Warp &w = warps[k];
        w.sched = &grid_ctx; w.warp_index = (int)k; w.tid_base = 32u * (unsigned)k;
        w.body = [&body, k](int lane) { body((int)k, lane); };
        prepare_lanes(w);
    }
    bool ok = true;
    for (uint64_t pass = 0;; ++pass) {
        bool all_done = true, any_progress = false;
        for (size_t kk = 0; kk < warps.size(); ++kk) {
            Warp &w = warps[order.empty() ? kk : (size_t)order[kk % order.size()] % warps.size()];
            const uint64_t before = w.progress;
            for (int i = 0; i < LANES; ++i) {
                if (w.done[i]) continue;
                all_done = false;
                w.cur = i;
                current() = &w;
                swapcontext(&grid_ctx, &w.ctx[i]);
            }
            any_progress |= w.progress != before;
        }
        if (all_done) break;
        if (!any_progress || pass >= max_passes) { ok = false; break; }
    }
    current() = prev;
    return ok;
}

struct ThreadIdx { unsigned x, y, z; ThreadIdx() : x(current()->tid_base + lane()), y(0), z(0) {} };

}  // namespace warp_emu

// ---- the CUDA names the code under test uses ------------------------------------------------------------------------
#define __device__
#define __forceinline__ inline
#define threadIdx (warp_emu::ThreadIdx())

struct float2 { float x, y; };
struct uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }

inline uint32_t __shfl_sync(uint32_t, uint32_t v, uint32_t src) { return warp_emu::exchange(v, src); }
inline float __shfl_sync(uint32_t, float v, uint32_t src)
{
    uint32_t u; __builtin_memcpy(&u, &v, 4);
    u = warp_emu::exchange(u, src);
    float r; __builtin_memcpy(&r, &u, 4);
    return r;
}
inline int __shfl_sync(uint32_t, int v, uint32_t src) { return (int)warp_emu::exchange((uint32_t)v, src); }
inline uint32_t __shfl_up_sync(uint32_t, uint32_t v, uint32_t delta)
{
    const unsigned l = warp_emu::lane();
    const uint32_t got = warp_emu::exchange(v, l >= delta ? l - delta : l);
    return l >= delta ? got : v;
}
inline uint32_t __shfl_xor_sync(uint32_t, uint32_t v, uint32_t lane_mask) { return warp_emu::exchange(v, warp_emu::lane() ^ lane_mask); }
inline uint32_t __ballot_sync(uint32_t, bool pred) { return warp_emu::gather_mask(pred); }
inline bool __any_sync(uint32_t, bool pred) { return warp_emu::gather_mask(pred) != 0u; }
inline void __syncwarp() { warp_emu::rendezvous(); }
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline uint32_t __float_as_uint(float f) { uint32_t u; __builtin_memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; __builtin_memcpy(&f, &u, 4); return f; }
inline float __frcp_rn(float x) { return 1.0f / x; }                         // IEEE division: correctly rounded, as rcp.rn
inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { const uint32_t old = *p; *p = old + v; return old; }
inline void __threadfence() {}
inline void __nanosleep(unsigned) { warp_emu::yield_lane(); }
struct float4 { float x, y, z, w; };
inline float4 __ldcg(const float4 *p) { return *p; }
inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v)
{
    const unsigned long long old = *p;                                       // lanes never run concurrently
    if (v < old) *p = v;
    return old;
}

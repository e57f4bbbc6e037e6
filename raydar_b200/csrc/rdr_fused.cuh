// rdr_fused.cuh -- the fused two-level scan (MODE 5): the warp-cooperative cluster scan of rdr_device.cuh rebuilt
// around what the issue-slot profile of sm_100a rewards (profiles/microbench_pipes_r01l.txt):
//
//   * packed FP32: the box tests run two boxes per instruction (fma.rn.f32x2 -> SASS FFMA2: two FMAs in ONE issue
//     slot; the kernel is issue-bound, not FMA-pipe-bound).  The boxes are stored pair-wise in FFMA2 operand order;
//     the per-ray constants enter as broadcast scalars with |x| / -|x| operand modifiers, so they cost no registers;
//   * A0 (ray x every top-level box) reads the boxes from the kernel parameters with uniform constant-bank loads
//     (LDCU.128 into uniform registers): no shared-memory traffic at all in the uniform stage;
//   * no atomics and no divergent pushes on the survivor path: the member stage leaves an 8-bit mask per (ray, cluster)
//     task, the masks are compacted into the warp's survivor lists with one packed shuffle prefix sum (sphere count in
//     the low half, cube count in the high half), and the list lengths live in (warp-uniform) registers;
//   * one call site for each exact stage (the loop flushes full groups of 32 and, after the last round, the rest), and
//     no sphere pre-test in front of the exact sphere test (under SIMT it never saves the exact test, it only adds to it);
//   * a ray that must skip the cull (origin outside the scene bound, non-finite) zeroes its slab constants: every box
//     test then evaluates to tn = tf = 0 and passes, so no "test everything" flag travels with the tasks.
//
// The stages and the winner rule are those of trace_cluster / trace_cluster_coop: conservative boxes only choose which
// primitives get the exact, reference-ordered test; the winner is the (t, original index) minimum of trace_ray
// (cpu.rs:344-352), folded per ray with a 64-bit atomicMin on an order-preserving key.
// Device-only; must be entered by all 32 lanes of a warp.  Needs lay.fused_ok (<= 32 top entries) and a staged blob.
#pragma once

// RDR_WARP_EMU (tests/hostsim only): this file is compiled by g++ against tests/hostsim/warp_emu.h, which runs the 32
// lanes of a warp as fibers and supplies __shfl_sync / __ballot_sync / __syncwarp / atomicMin / threadIdx; the four
// packed-FP32 helpers and the predicated OR below then have plain C++ bodies (one fmaf per half = the rounding of
// fma.rn.f32x2).  The device build never defines it.
#ifdef RDR_WARP_EMU
#include "rdr_trace.cuh"
#else
#include "rdr_device.cuh"
#endif

namespace rdr {

// RDR_EMU_STATS (CPU warp emulator only): per-warp work counters of the fused scan, bumped by lane 0
#if defined(RDR_WARP_EMU) && defined(RDR_EMU_STATS)
struct FusedEmuStats { unsigned long long traces, tasks, member_rounds, sphere_rounds, cube_rounds, sphere_tests, cube_tests; };
inline FusedEmuStats &fused_emu_stats() { static thread_local FusedEmuStats s{}; return s; }
#define RDR_FSTAT(field, v) do { if (lane == 0u) fused_emu_stats().field += (v); } while (0)
#else
#define RDR_FSTAT(field, v) do { } while (0)
#endif

typedef unsigned long long f32x2;

#ifdef RDR_WARP_EMU
inline f32x2 pk2(float lo, float hi) { return (f32x2)f2u(lo) | ((f32x2)f2u(hi) << 32); }
inline f32x2 bc2(float a) { return pk2(a, a); }
inline void un2(f32x2 v, float &lo, float &hi) { lo = u2f((uint32_t)v); hi = u2f((uint32_t)(v >> 32)); }
inline f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    float al, ah, bl, bh, cl, ch;
    un2(a, al, ah); un2(b, bl, bh); un2(c, cl, ch);
    return pk2(fma(al, bl, cl), fma(ah, bh, ch));
}
#else
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 bc2(float a) { f32x2 r; asm("mov.b64 %0, {%1,%1};" : "=l"(r) : "f"(a)); return r; }
__device__ __forceinline__ void un2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
#endif

// per-warp scratch of the fused scan (shared memory)
constexpr uint32_t FUSED_TASK_CAP = 32u * FUSED_MAX_TOP;      // every ray against every cluster
constexpr uint32_t FUSED_SURV_CAP = 128u + 31u + 256u + 1u;   // direct entries + carried remainder + one round of 32 x 8
constexpr uint32_t FUSED_WARP_BYTES = 32u * 8u + FUSED_TASK_CAP * 2u + 2u * FUSED_SURV_CAP * 2u;

struct FusedWarp {
    unsigned long long *best;     // [32] winner key per lane
    uint16_t *tasks;              // [FUSED_TASK_CAP]  owner lane << 8 | top entry
    uint16_t *surv_s, *surv_c;    // [FUSED_SURV_CAP]  owner lane << 11 | member slot
};

__device__ __forceinline__ FusedWarp fused_warp(unsigned char *base, uint32_t warp)
{
    unsigned char *p = base + (size_t)warp * FUSED_WARP_BYTES;
    FusedWarp w;
    w.best = reinterpret_cast<unsigned long long *>(p);
    w.tasks = reinterpret_cast<uint16_t *>(p + 256u);
    w.surv_s = w.tasks + FUSED_TASK_CAP;
    w.surv_c = w.surv_s + FUSED_SURV_CAP;
    return w;
}

// the fused scan's sections of the staged blob.  The kernel derives the pointers from the `extern __shared__` array
// itself (stage_scene<5> in rdr_kernels.cu), so every access compiles to LDS with a 32-bit address; a pointer that
// might also be global compiles to LD.E with 64-bit address arithmetic.
struct FusedView {
    const f4 *pair_block, *member_geom;       // member_geom / member_idx: the fused clustering's own arrays (slot = C * cluster + member)
    const uint32_t *member_idx;
};

// inclusive prefix sum over the warp.  The shuffle's own "source lane in range" predicate guards the add: SHFL.UP P +
// @P IADD per step instead of SHFL + SEL + IADD with a separate lane compare (11 instead of ~20 instructions).
__device__ __forceinline__ uint32_t warp_scan_incl(uint32_t v, uint32_t lane)
{
#pragma unroll
    for (uint32_t off = 1u; off < 32u; off <<= 1) {
#ifndef RDR_WARP_EMU
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tshfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n\t@p add.u32 %0, %0, t;\n\t}"
                     : "+r"(v) : "r"(off));
        (void)lane;
#else
        const uint32_t u = __shfl_up_sync(0xffffffffu, v, off);
        if (lane >= off) v += u;
#endif
    }
    return v;
}

// Appends the survivors of this lane's task -- bit j of `bits` = member slot first + j * stride of ray `owner`, the low
// `ns` bits are spheres -- to the warp's lists.  n_s / n_c: list lengths, warp-uniform.
__device__ __forceinline__ void fused_append(FusedWarp ws, uint32_t lane, uint32_t owner, uint32_t first, uint32_t stride,
                                             uint32_t bits, uint32_t ns, uint32_t &n_s, uint32_t &n_c)
{
    const uint32_t sb = bits & ((1u << ns) - 1u), cb = bits ^ sb;
    const uint32_t packed = __popc(sb) | (__popc(cb) << 16);
    const uint32_t incl = warp_scan_incl(packed, lane);
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t excl = incl - packed;
    uint32_t ps = n_s + (excl & 0xffffu), pc = n_c + (excl >> 16);
    const uint32_t tag = owner << 11;
    // one loop for both lists (surv_c = surv_s + FUSED_SURV_CAP): the low ns bits go to the sphere list
    pc += FUSED_SURV_CAP;
    while (bits != 0u) {
        const uint32_t j = (uint32_t)__ffs((int)bits) - 1u; bits &= bits - 1u;
        const bool sphere = j < ns;
        ws.surv_s[sphere ? ps : pc] = (uint16_t)(tag | (first + j * stride));
        ps += sphere ? 1u : 0u; pc += sphere ? 0u : 1u;
    }
    n_s += total & 0xffffu; n_c += total >> 16;
    __syncwarp();
}

// exact, reference-ordered test of survivors [base, base + n) of a list (n <= 32), one per lane; the owner's ray
// comes by indexed shuffle; a hit is folded into the owner's winner key
template <bool SPHERE>
__device__ __forceinline__ void fused_exact(const FusedView &V, FusedWarp ws, uint32_t lane, const uint16_t *list, uint32_t base,
                                            uint32_t n, v3 o, v3 d)
{
    const uint32_t FULL = 0xffffffffu;
    const bool has = lane < n;
    const uint32_t e = has ? (uint32_t)list[base + lane] : (lane << 11);
    const uint32_t own = e >> 11, slot = e & 0x7ffu;
    const v3 ro = mk3(__shfl_sync(FULL, o.x, own), __shfl_sync(FULL, o.y, own), __shfl_sync(FULL, o.z, own));
    const v3 rd = mk3(__shfl_sync(FULL, d.x, own), __shfl_sync(FULL, d.y, own), __shfl_sync(FULL, d.z, own));
    if (has) {
        const f4 g = V.member_geom[slot];
        float tt;
        const bool hit = SPHERE ? hit_sphere_exact(ro, rd, mk3(g.x, g.y, g.z), g.w, &tt) : hit_cube_exact(ro, rd, mk3(g.x, g.y, g.z), g.w, &tt);
        if (hit) atomicMin(&ws.best[own], coop_key(tt, (int)V.member_idx[slot]));
    }
    __syncwarp();
}

// two boxes (A, B) against one ray: bit 0 / bit 1 of the result = box A / B may be hit.
//   c*: centres, e*: half-extents (already inflated), r*: the ray's 1/d, n*: -(o/d)
// slab_pair_acc ORs the two bits into `m` at compile-time position SHIFT: one compare and one predicated OR per box
// (FSETP + @!P LOP3) instead of the select / add / shift / or chain the compiler builds from the C form.
__device__ __forceinline__ void or_unless_gt(uint32_t &m, float tn, float tf, uint32_t bit)
{
    // tn > tf is false for a NaN operand: the bit is set, as in `(tn > tf ? 0 : bit)`
#ifdef RDR_WARP_EMU
    if (!(tn > tf)) m |= bit;
#else
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\t@!p or.b32 %0, %0, %3;\n\t}" : "+r"(m) : "f"(tn), "f"(tf), "r"(bit));
#endif
}
template <bool BOUNDED = false>
__device__ __forceinline__ void slab_pair_tn_tf(f32x2 cx, f32x2 cy, f32x2 cz, f32x2 ex, f32x2 ey, f32x2 ez,
                                                float rx, float ry, float rz, float nx, float ny, float nz, float best,
                                                float &tna, float &tfa, float &tnb, float &tfb)
{
    const f32x2 tcx = fma2(cx, bc2(rx), bc2(nx)), tcy = fma2(cy, bc2(ry), bc2(ny)), tcz = fma2(cz, bc2(rz), bc2(nz));
    float nxa, nxb, nya, nyb, nza, nzb, fxa, fxb, fya, fyb, fza, fzb;
    un2(fma2(ex, bc2(-fabsf(rx)), tcx), nxa, nxb); un2(fma2(ey, bc2(-fabsf(ry)), tcy), nya, nyb); un2(fma2(ez, bc2(-fabsf(rz)), tcz), nza, nzb);
    un2(fma2(ex, bc2(fabsf(rx)), tcx), fxa, fxb); un2(fma2(ey, bc2(fabsf(ry)), tcy), fya, fyb); un2(fma2(ez, bc2(fabsf(rz)), tcz), fza, fzb);
    // BOUNDED: a box whose entry distance exceeds the ray's best exact t so far cannot hold the winner (ties kept);
    // min.f32 ignores a NaN operand, so "no hit yet" is passed as NaN
    tna = fmaxf(fmaxf(nxa, nya), fmaxf(nza, 0.0f)); tfa = BOUNDED ? fminf(fminf(fxa, fya), fminf(fza, best)) : fminf(fminf(fxa, fya), fza);
    tnb = fmaxf(fmaxf(nxb, nyb), fmaxf(nzb, 0.0f)); tfb = BOUNDED ? fminf(fminf(fxb, fyb), fminf(fzb, best)) : fminf(fminf(fxb, fyb), fzb);
}
template <bool BOUNDED = false>
__device__ __forceinline__ uint32_t slab_pair(f32x2 cx, f32x2 cy, f32x2 cz, f32x2 ex, f32x2 ey, f32x2 ez,
                                              float rx, float ry, float rz, float nx, float ny, float nz, float best = 0.0f)
{
    float tna, tfa, tnb, tfb;
    slab_pair_tn_tf<BOUNDED>(cx, cy, cz, ex, ey, ez, rx, ry, rz, nx, ny, nz, best, tna, tfa, tnb, tfb);
    return (tna > tfa ? 0u : 1u) | (tnb > tfb ? 0u : 2u);
}
template <uint32_t SHIFT>
__device__ __forceinline__ void slab_pair_acc(uint32_t &m, f32x2 cx, f32x2 cy, f32x2 cz, f32x2 ex, f32x2 ey, f32x2 ez,
                                              float rx, float ry, float rz, float nx, float ny, float nz)
{
    float tna, tfa, tnb, tfb;
    slab_pair_tn_tf<false>(cx, cy, cz, ex, ey, ez, rx, ry, rz, nx, ny, nz, 0.0f, tna, tfa, tnb, tfb);
    or_unless_gt(m, tna, tfa, 1u << SHIFT);
    or_unless_gt(m, tnb, tfb, 2u << SHIFT);
}

// slab constants of one ray for the pair tests (make_ray_bvh, rdr_core.cuh): r = 1/d (clamped to +-1e30 for zero /
// denormal components), n = -(o/d), rho = per-ray inflation of sphere-flagged boxes.  A ray that must skip the cull
// (origin outside the scene bound, non-finite) gets all-zero constants: every box test then passes (tn = tf = 0).
struct SlabRay { float rx, ry, rz, nx, ny, nz, rho; };

__device__ __forceinline__ SlabRay slab_ray_setup(const CullConsts &cc, v3 o, v3 d)
{
    SlabRay R;
    R.rx = __frcp_rn(d.x); R.ry = __frcp_rn(d.y); R.rz = __frcp_rn(d.z);         // == 1.0f / d, correctly rounded
    const float big = 1e30f;
    if (!(fabsf(R.rx) < big)) R.rx = copysignf(big, d.x);
    if (!(fabsf(R.ry) < big)) R.ry = copysignf(big, d.y);
    if (!(fabsf(R.rz) < big)) R.rz = copysignf(big, d.z);
    const float a = fma(d.x, d.x, fma(d.y, d.y, fmul(d.z, d.z)));
    const float oo = fma(o.x, o.x, fma(o.y, o.y, fmul(o.z, o.z)));
    const float s_ray = fma(2.0f, oo, cc.sphere_q_max);
    const float Ms = fmul(1.9073486328125e-06f, s_ray);
    const float omax = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z));
    const bool degenerate = !(omax <= cc.origin_bound) || !(a > 1e-30f) || !(a < 1e30f) || !(s_ray < 1e30f) ||
                            isnan_(d.x) || isnan_(d.y) || isnan_(d.z);
    {
        // rho from sqrt.approx (2 ulp): (q1 - r_min) loses at most 2 ulp(q1) <= 2^-22 (r_min + rho), made up by the 1.0001
        // factor and by widening with r_min too -- two MUFU.SQRT instead of two correctly rounded square roots
        float q1, q2;
#ifdef RDR_WARP_EMU
        q1 = fsqrt(fma(cc.sphere_r_min, cc.sphere_r_min, Ms)); q2 = fsqrt(s_ray);
#else
        asm("sqrt.approx.f32 %0, %1;" : "=f"(q1) : "f"(fma(cc.sphere_r_min, cc.sphere_r_min, Ms)));
        asm("sqrt.approx.f32 %0, %1;" : "=f"(q2) : "f"(s_ray));
#endif
        R.rho = fadd(fsub(q1, cc.sphere_r_min), fmul(1.9073486328125e-06f, q2));
        R.rho = fma(R.rho, 1.0001f, fmul(4.76837158203125e-07f, cc.sphere_r_min));
    }
    R.nx = fneg(fmul(o.x, R.rx)); R.ny = fneg(fmul(o.y, R.ry)); R.nz = fneg(fmul(o.z, R.rz));
    if (degenerate) { R.rx = R.ry = R.rz = 0.0f; R.nx = R.ny = R.nz = 0.0f; R.rho = 0.0f; }
    return R;
}

// the ray against the first n_top (<= 32) top-level boxes of the kernel parameters: bit k = box k may be hit.
// The operands are constant-bank addresses (LDCU.128 into uniform registers).
template <uint32_t SHIFT>
__device__ __forceinline__ void top_pair_test(uint32_t &m, const TopPair &t, f32x2 rho2, const SlabRay &R)
{
    const f32x2 sp = pk2(t.sphere[0], t.sphere[1]);
    const f32x2 ex = fma2(sp, rho2, pk2(t.ex[0], t.ex[1])), ey = fma2(sp, rho2, pk2(t.ey[0], t.ey[1])), ez = fma2(sp, rho2, pk2(t.ez[0], t.ez[1]));
    slab_pair_acc<SHIFT>(m, pk2(t.cx[0], t.cx[1]), pk2(t.cy[0], t.cy[1]), pk2(t.cz[0], t.cz[1]), ex, ey, ez, R.rx, R.ry, R.rz, R.nx, R.ny, R.nz);
}
template <uint32_t K>
__device__ __forceinline__ void top_pairs4(uint32_t &m, const TopParams &T, f32x2 rho2, const SlabRay &R)
{
    uint32_t g = 0u;        // own accumulator per group: keeps the predicated-OR dependency chains short
    top_pair_test<2u * K>(g, T.pair[K], rho2, R); top_pair_test<2u * K + 2u>(g, T.pair[K + 1u], rho2, R);
    top_pair_test<2u * K + 4u>(g, T.pair[K + 2u], rho2, R); top_pair_test<2u * K + 6u>(g, T.pair[K + 3u], rho2, R);
    m |= g;
}

// Fully unrolled, four pairs per step: every operand is a compile-time constant-bank address.  (A warp-uniform loop over
// groups of pairs indexes the constant bank with a vector register -- LDC per operand -- and measured 1-2 % slower.)
// (A warp-uniform loop over groups of pairs indexes the constant bank with a vector register -- LDC per operand instead of
// LDCU.128 -- and measured 2 - 5 % slower even though it takes 2.5 KB out of the hot loop: profiles/variants_r03i.txt.)
__device__ __forceinline__ uint32_t top_scan(const TopParams &T, uint32_t n_top, const SlabRay &R)
{
    uint32_t m = 0u;
    const f32x2 rho2 = bc2(R.rho);
    top_pairs4<0u>(m, T, rho2, R);
    if (n_top > 8u) top_pairs4<4u>(m, T, rho2, R);
    if (n_top > 16u) top_pairs4<8u>(m, T, rho2, R);
    if (n_top > 24u) top_pairs4<12u>(m, T, rho2, R);
    if (n_top < 32u) m &= (1u << n_top) - 1u;
    return m;
}

// CAP8: the clusters have 8 member slots (scenes up to ~250 objects): one member step per round, resolved at compile time
template <bool CAP8>
__device__ __forceinline__ Hit trace_fused(const FusedView &V, const FrameParams &P, FusedWarp ws, bool alive, v3 o, v3 d)
{
    const uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;

    const SlabRay R = slab_ray_setup(P.cull, o, d);
    const float rx = R.rx, ry = R.ry, rz = R.rz, nx = R.nx, ny = R.ny, nz = R.nz, rho = R.rho;

    ws.best[lane] = ~0ull;

    // ---- A0: this lane's ray against every top-level box, two per FFMA2, operands from the constant bank ----
    uint32_t m = top_scan(P.top, P.lay.fused_top, R);
    if (!alive) m = 0u;

    uint32_t n_s = 0u, n_c = 0u;                                                 // survivor list lengths (warp-uniform)
    // single-primitive top entries: their box was the entry -> straight to the survivor lists (member slot = C * entry)
    if (P.lay.fused_direct != 0u) {
        const uint32_t dmask = (1u << P.lay.fused_direct) - 1u;
        // one ballot per direct entry (<= 4, warp-uniform loop): the survivor's position is the popcount of the ballot
        // below the lane -- no prefix scan, no per-lane write loop
        const uint32_t below = (1u << lane) - 1u, cap = CAP8 ? 8u : P.lay.fused_cap;
#pragma unroll 1
        for (uint32_t j = 0; j < P.lay.fused_direct; ++j) {
            const bool hit = ((m >> j) & 1u) != 0u;
            const uint32_t b = __ballot_sync(FULL, hit);
            if (j < P.lay.fused_ns_direct) {
                if (hit) ws.surv_s[n_s + __popc(b & below)] = (uint16_t)((lane << 11) | (cap * j));
                n_s += __popc(b);
            } else {
                if (hit) ws.surv_c[n_c + __popc(b & below)] = (uint16_t)((lane << 11) | (cap * j));
                n_c += __popc(b);
            }
        }
        __syncwarp();
        m &= ~dmask;
    }

    // ---- T: compact the (ray, cluster) pairs of the warp into one task list ----
    const uint32_t cnt = __popc(m);
    const uint32_t pre = warp_scan_incl(cnt, lane);
    const uint32_t total = __shfl_sync(FULL, pre, 31);
    RDR_FSTAT(traces, 1u); RDR_FSTAT(tasks, total);
    {
        uint32_t pos = pre - cnt, mm = m;
        while (mm != 0u) {
            const uint32_t k = (uint32_t)__ffs((int)mm) - 1u; mm &= mm - 1u;
            ws.tasks[pos++] = (uint16_t)((lane << 8) | k);
        }
    }
    __syncwarp();

    // ---- rounds of 32 tasks x groups of 8 members: M member boxes -> 8-bit masks -> survivor lists;  E exact tests on
    //      full groups of 32 survivors, and on whatever is left after the last round ----
    const uint32_t per_round = CAP8 ? 1u : (P.lay.fused_cap >> 3);   // 8 members (4 pairs) per step; C is a multiple of 8
    uint32_t t0 = 0u, g = 0u;
#pragma unroll 1
    for (;;) {
        const bool last = t0 >= total;
        if (!last) {
            RDR_FSTAT(member_rounds, 1u);
            const uint32_t t = t0 + lane;
            const bool has = t < total;
            const uint32_t task = has ? (uint32_t)ws.tasks[t] : (lane << 8);
            const uint32_t owner = task >> 8;
            const float qx = __shfl_sync(FULL, rx, owner), qy = __shfl_sync(FULL, ry, owner), qz = __shfl_sync(FULL, rz, owner);
            const float mx = __shfl_sync(FULL, nx, owner), my = __shfl_sync(FULL, ny, owner), mz = __shfl_sync(FULL, nz, owner);
            const f32x2 rho2 = bc2(__shfl_sync(FULL, rho, owner));
            const f4 *blk = V.pair_block + (CAP8 ? 13u : P.lay.fused_stride) * (task & 0xffu);
            const uint32_t m0 = CAP8 ? 0u : 8u * g;               // first member of this step
            const f4 *pb = blk + 3u * (m0 >> 1);
            uint32_t bits = 0u, desc = 0u;
            if (!CAP8) desc = __float_as_uint(blk[2].z);
#pragma unroll
            for (uint32_t p = 0; p < 4u; ++p) {
                const f4 q0 = pb[3u * p], q1 = pb[3u * p + 1u];
                float2 q2;
                if (CAP8 && p == 0u) { const f4 q = pb[2]; q2.x = q.x; q2.y = q.y; desc = __float_as_uint(q.z); }   // flags + desc in one LDS.128
                else q2 = *reinterpret_cast<const float2 *>(pb + 3u * p + 2u);
                const f32x2 e = fma2(pk2(q2.x, q2.y), rho2, pk2(q1.z, q1.w));
                if (p == 0u) slab_pair_acc<0u>(bits, pk2(q0.x, q0.y), pk2(q0.z, q0.w), pk2(q1.x, q1.y), e, e, e, qx, qy, qz, mx, my, mz);
                else if (p == 1u) slab_pair_acc<2u>(bits, pk2(q0.x, q0.y), pk2(q0.z, q0.w), pk2(q1.x, q1.y), e, e, e, qx, qy, qz, mx, my, mz);
                else if (p == 2u) slab_pair_acc<4u>(bits, pk2(q0.x, q0.y), pk2(q0.z, q0.w), pk2(q1.x, q1.y), e, e, e, qx, qy, qz, mx, my, mz);
                else slab_pair_acc<6u>(bits, pk2(q0.x, q0.y), pk2(q0.z, q0.w), pk2(q1.x, q1.y), e, e, e, qx, qy, qz, mx, my, mz);
            }
            const uint32_t count = desc & 63u, n_sph = (desc >> 6) & 63u;
            const uint32_t left = count > m0 ? count - m0 : 0u;   // members of this cluster in this step
            bits &= left >= 8u ? 0xffu : (1u << left) - 1u;
            if (!has) bits = 0u;
            const uint32_t ns = n_sph > m0 ? (n_sph - m0 < 8u ? n_sph - m0 : 8u) : 0u;
            fused_append(ws, lane, owner, (desc >> 12) + m0, 1u, bits, ns, n_s, n_c);
            if (CAP8 || ++g == per_round) { g = 0u; t0 += 32u; }
        }
#pragma unroll 1
        while (n_s >= 32u || (last && n_s != 0u)) {
            const uint32_t n = n_s < 32u ? n_s : 32u;
            n_s -= n;
            RDR_FSTAT(sphere_rounds, 1u); RDR_FSTAT(sphere_tests, n);
            fused_exact<true>(V, ws, lane, ws.surv_s, n_s, n, o, d);
        }
#pragma unroll 1
        while (n_c >= 32u || (last && n_c != 0u)) {
            const uint32_t n = n_c < 32u ? n_c : 32u;
            n_c -= n;
            RDR_FSTAT(cube_rounds, 1u); RDR_FSTAT(cube_tests, n);
            fused_exact<false>(V, ws, lane, ws.surv_c, n_c, n, o, d);
        }
        if (last) break;
    }
    __syncwarp();
    const unsigned long long key = ws.best[lane];
    Hit h; h.idx = -1; h.t = finf();
    if (alive && key != ~0ull) {
        const uint32_t kt = (uint32_t)(key >> 32);
        h.idx = (int)(((uint32_t)key) >> 1);
        h.t = __uint_as_float((kt == 0u && (key & 1ull)) ? 0x80000000u : kt);
    }
    return h;
}

}  // namespace rdr

// rdr_trace.cuh -- per-lane nearest-hit search and the per-pixel sample loop.  __host__ __device__:
// the kernels in rdr_kernels.cu call these on the device; tests/hostsim compiles the same code for
// the CPU to check the logic without a GPU.
#pragma once

#include "raydar_cuda.h"
#include "rdr_layout.h"

namespace rdr {

// view of a packed scene blob (in shared memory on the device)
struct SceneView {
    const f4 *sphere_cull, *cube_cull, *sphere_geom, *cube_geom, *obj_geom, *material;
    const uint32_t *sphere_idx, *cube_idx;
    uint32_t n_spheres, n_cubes, ns_chunks, nc_chunks;
    const f4 *nodes;        // BVH mode: 16 quads per node
    uint32_t n_objects;
    const f4 *top, *member_box, *member_geom;       // cluster scan
    const uint32_t *member_idx;
    uint32_t n_top, nt_chunks, n_direct;
    const f4 *pair_block, *fused_geom;              // fused scan
    const uint32_t *fused_idx;
};

RDR_HD SceneView scene_view(const unsigned char *base, const SceneLayout &L)
{
    SceneView s;
    s.sphere_cull = reinterpret_cast<const f4 *>(base + L.off_sphere_cull);
    s.cube_cull = reinterpret_cast<const f4 *>(base + L.off_cube_cull);
    s.sphere_geom = reinterpret_cast<const f4 *>(base + L.off_sphere_geom);
    s.cube_geom = reinterpret_cast<const f4 *>(base + L.off_cube_geom);
    s.obj_geom = reinterpret_cast<const f4 *>(base + L.off_obj_geom);
    s.material = reinterpret_cast<const f4 *>(base + L.off_material);
    s.sphere_idx = reinterpret_cast<const uint32_t *>(base + L.off_sphere_idx);
    s.cube_idx = reinterpret_cast<const uint32_t *>(base + L.off_cube_idx);
    s.n_spheres = L.n_spheres; s.n_cubes = L.n_cubes;
    s.ns_chunks = L.ns_pad >> 5; s.nc_chunks = L.nc_pad >> 5;
    s.nodes = reinterpret_cast<const f4 *>(base + L.off_nodes);
    s.n_objects = L.n_objects;
    s.top = reinterpret_cast<const f4 *>(base + L.off_top);
    s.member_box = reinterpret_cast<const f4 *>(base + L.off_member_box);
    s.member_geom = reinterpret_cast<const f4 *>(base + L.off_member_geom);
    s.member_idx = reinterpret_cast<const uint32_t *>(base + L.off_member_idx);
    s.n_top = L.n_top; s.nt_chunks = L.nt_pad >> 5; s.n_direct = L.n_direct;
    s.pair_block = reinterpret_cast<const f4 *>(base + L.off_pair_block);
    s.fused_geom = reinterpret_cast<const f4 *>(base + L.off_fused_geom);
    s.fused_idx = reinterpret_cast<const uint32_t *>(base + L.off_fused_idx);
    return s;
}

RDR_HD Material load_material(const SceneView &S, int idx, bool *is_sphere)
{
    const f4 m0 = S.material[3 * idx + 0], m1 = S.material[3 * idx + 1], m2 = S.material[3 * idx + 2];
    Material m;
    m.albedo = mk3(m0.x, m0.y, m0.z); m.roughness = m0.w;
    m.emission = mk3(m1.x, m1.y, m1.z); m.emission_strength = m1.w;
    m.metallic = m2.x; m.transmission = m2.y; m.ior = m2.z;
    *is_sphere = (f2u(m2.w) == 0u);
    return m;
}

// optional instrumentation (host-simulation only; compiled out on the device)
struct TraceStats { uint64_t traces, sphere_exact, cube_exact, degenerate, nodes_visited, entries_hit; };
#if defined(__CUDA_ARCH__)
#define RDR_STAT(stats, field)
#else
#define RDR_STAT(stats, field) do { if (stats) (stats)->field++; } while (0)
#endif

// ---- brute-force nearest hit over the SoA buffer (trace_ray, cpu.rs:344-352) -------------------------
// Phase A: a uniform scan in which every lane tests the same primitive with the conservative FMA
//   test and records survivors in a per-lane 32-bit mask per chunk (no divergent code).
// Phase B: each lane walks its own survivor list with the exact, reference-ordered test.  Lanes
//   with different survivors execute the same instructions, so a warp spends max-over-lanes
//   (not sum-over-lanes) exact tests.
// masks: per-lane scratch words, masks[chunk * stride] (shared memory on the device).
template <bool USE_CULL>
RDR_HD Hit trace_brute(const SceneView &S, const CullConsts &cc, uint32_t *masks, uint32_t stride, v3 o, v3 d,
                       TraceStats *stats = nullptr)
{
    Hit best; best.idx = -1; best.t = finf();
    RayCull rc;
    bool all = !USE_CULL;
    if (USE_CULL) { rc = make_ray_cull(o, d, cc); all = rc.degenerate; }
    RDR_STAT(stats, traces);
    if (USE_CULL && all) RDR_STAT(stats, degenerate);

    // ---- spheres ----
    for (uint32_t ch = 0; ch < S.ns_chunks; ++ch) {
        uint32_t m = 0u;
        if (USE_CULL) {
            const f4 *p = S.sphere_cull + ch * 32u;
            // groups of 8 with compile-time bit positions: one predicated OR-immediate per primitive
            RDR_NOUNROLL
            for (int g = 0; g < 4; ++g) {
                uint32_t mm = 0u;
                RDR_UNROLL
                for (int j = 0; j < 8; ++j) {
                    const f4 s = p[g * 8 + j];
                    if (sphere_may_hit(o, d, rc, s.x, s.y, s.z, s.w)) mm |= (1u << j);
                }
                m |= mm << (g * 8);
            }
        }
        if (all) m = 0xffffffffu;
        const uint32_t left = S.n_spheres - ch * 32u;
        if (left < 32u) m &= (1u << left) - 1u;
        masks[ch * stride] = m;
    }
    {
        uint32_t ch = 0u, m = S.ns_chunks ? masks[0] : 0u;
        for (;;) {
            while (m == 0u && ++ch < S.ns_chunks) m = masks[ch * stride];
            if (m == 0u) break;
            const int j = ffs32(m); m &= m - 1u;
            const uint32_t li = ch * 32u + (uint32_t)j;
            const f4 g = S.sphere_geom[li];
            float t;
            RDR_STAT(stats, sphere_exact);
            if (hit_sphere_exact(o, d, mk3(g.x, g.y, g.z), g.w, &t)) {
                const int idx = (int)S.sphere_idx[li];
                if (hit_better(t, idx, best.t, best.idx)) { best.idx = idx; best.t = t; }
            }
        }
    }
    // ---- cubes (pruned by the best sphere hit) ----
    {
        const float prune = (best.idx >= 0 && !isnan_(best.t)) ? best.t : finf();
        for (uint32_t ch = 0; ch < S.nc_chunks; ++ch) {
            uint32_t m = 0u;
            if (USE_CULL) {
                const f4 *p = S.cube_cull + ch * 32u;
                RDR_NOUNROLL
                for (int g = 0; g < 4; ++g) {
                    uint32_t mm = 0u;
                    RDR_UNROLL
                    for (int j = 0; j < 8; ++j) {
                        const f4 c = p[g * 8 + j];
                        if (cube_may_hit(rc, c.x, c.y, c.z, c.w, prune)) mm |= (1u << j);
                    }
                    m |= mm << (g * 8);
                }
            }
            if (all) m = 0xffffffffu;
            const uint32_t left = S.n_cubes - ch * 32u;
            if (left < 32u) m &= (1u << left) - 1u;
            masks[ch * stride] = m;
        }
        uint32_t ch = 0u, m = S.nc_chunks ? masks[0] : 0u;
        for (;;) {
            while (m == 0u && ++ch < S.nc_chunks) m = masks[ch * stride];
            if (m == 0u) break;
            const int j = ffs32(m); m &= m - 1u;
            const uint32_t li = ch * 32u + (uint32_t)j;
            const f4 g = S.cube_geom[li];
            float t;
            RDR_STAT(stats, cube_exact);
            if (hit_cube_exact(o, d, mk3(g.x, g.y, g.z), g.w, &t)) {
                const int idx = (int)S.cube_idx[li];
                if (hit_better(t, idx, best.t, best.idx)) { best.idx = idx; best.t = t; }
            }
        }
    }
    return best;
}

// ---- BVH nearest hit (same winner as the scan; see rdr_bvh.h for the conservative boxes) ----------------
// Repeats two lock-step-friendly phases until the lane's stack and candidate queue are empty:
//   T  pop nodes, test their 8 entries with the FMA slab test (pruned by the best exact t so far); child
//      nodes go to the stack (with their entry distance, so that stale ones are skipped after `best`
//      improves), primitives go to the lane's candidate queue;
//   E  exact, reference-ordered tests on the queued primitives -- spheres first, then cubes, so that lanes of
//      a warp run the same code -- updating the (t, original index) winner.
// queue: QCAP words per lane (queue[i * stride]); stack: 64 (node, tn) pairs in local memory.
constexpr int BVH_QCAP = 16;
constexpr int BVH_STACK = 64;

RDR_HD void bvh_exact_prim(const SceneView &S, uint32_t payload, v3 o, v3 d, Hit &best, TraceStats *stats)
{
    const int idx = (int)(payload & 0x3fffffffu);
    const f4 g = S.obj_geom[idx];
    float t;
    bool hit;
    if (payload & 0x40000000u) { RDR_STAT(stats, cube_exact); hit = hit_cube_exact(o, d, mk3(g.x, g.y, g.z), g.w, &t); }
    else { RDR_STAT(stats, sphere_exact); hit = hit_sphere_exact(o, d, mk3(g.x, g.y, g.z), g.w, &t); }
    if (hit && hit_better(t, idx, best.t, best.idx)) { best.idx = idx; best.t = t; }
}

RDR_HD Hit trace_bvh(const SceneView &S, const CullConsts &cc, uint32_t *queue, uint32_t stride, v3 o, v3 d,
                     TraceStats *stats = nullptr)
{
    Hit best; best.idx = -1; best.t = finf();
    RDR_STAT(stats, traces);
    if (S.n_objects == 0u) return best;
    const RayBvh rb = make_ray_bvh(o, d, cc);
    if (rb.rc.degenerate) {                       // origin outside the scene bound / non-finite ray: exact test on everything
        RDR_STAT(stats, degenerate);
        RDR_NOUNROLL
        for (uint32_t i = 0; i < S.n_objects; ++i) {
            const uint32_t kind_bits = f2u(S.material[3 * i + 2].w);
            bvh_exact_prim(S, i | (kind_bits ? 0x40000000u : 0u), o, d, best, stats);
        }
        return best;
    }
    uint32_t stk_node[BVH_STACK];
    float stk_tn[BVH_STACK];
    int sp = 1;
    stk_node[0] = 0u; stk_tn[0] = 0.0f;
    uint32_t nq = 0u;
    float prune = finf();
    for (;;) {
        // ---- phase T ----
        while (sp > 0 && nq <= (uint32_t)(BVH_QCAP - 8)) {
            --sp;
            const uint32_t node = stk_node[sp];
            if (stk_tn[sp] > prune) continue;
            RDR_STAT(stats, nodes_visited);
            const f4 *e = S.nodes + (size_t)node * 16u;
            RDR_UNROLL
            for (int k = 0; k < 8; ++k) {
                const f4 q0 = e[2 * k], q1 = e[2 * k + 1];
                float tn;
                if (bvh_entry_may_hit(rb, q0, q1, prune, &tn)) {
                    RDR_STAT(stats, entries_hit);
                    const uint32_t payload = f2u(q1.z);
                    if (payload & 0x80000000u) { queue[nq * stride] = payload; ++nq; }
                    else { stk_node[sp] = payload; stk_tn[sp] = tn; ++sp; }
                }
            }
        }
        // ---- phase E ----
        RDR_NOUNROLL
        for (uint32_t q = 0; q < nq; ++q) {
            const uint32_t payload = queue[q * stride];
            if (!(payload & 0x40000000u)) bvh_exact_prim(S, payload, o, d, best, stats);
        }
        RDR_NOUNROLL
        for (uint32_t q = 0; q < nq; ++q) {
            const uint32_t payload = queue[q * stride];
            if (payload & 0x40000000u) bvh_exact_prim(S, payload, o, d, best, stats);
        }
        nq = 0u;
        if (best.idx >= 0 && !isnan_(best.t)) prune = best.t;
        if (sp == 0) break;
    }
    return best;
}

// ---- two-level ("cluster") scan -------------------------------------------------------------------------------
// The flat scan tests every primitive for every ray.  Here the primitives are grouped on the host into spatially
// coherent clusters of <= 8 (rdr_bvh.h build_clusters; a large primitive such as the floor stays alone):
//   A0  uniform scan over the cluster boxes -- every lane tests the same box, broadcast LDS.128, no divergence --
//       leaving a per-lane bit mask of the clusters the ray may touch (typically 2-4 of ~24 on benchmark.rscn);
//   A1  each lane walks ITS clusters and slab-tests their members (per-lane shared-memory reads, one quad per
//       member); survivors go to the lane's candidate queue;
//   B   exact, reference-ordered tests on the queue, spheres first (behind the cheap sphere pre-test), then cubes.
// The boxes are conservative w.r.t. the as-written tests exactly as in the BVH (per-ray rho on sphere boxes), and
// the winner is the (t, original index) minimum, so the result equals the flat scan's.
// scratch: per-lane words, [0 .. nt_chunks) cluster masks, [CL_QBASE .. CL_QBASE + BVH_QCAP) the queue.
constexpr uint32_t CL_QBASE = 4u;            // up to 128 clusters (1024 primitives)
constexpr uint32_t CL_SCRATCH = CL_QBASE + (uint32_t)BVH_QCAP;

// member quad: (cx, cy, cz, +-(e)), negative e marks a sphere (its box grows by the per-ray rho)
RDR_HD bool member_may_hit(const RayBvh &rb, f4 m, float best)
{
    const RayCull &rc = rb.rc;
    const uint32_t sphere_mask = (uint32_t)((int32_t)f2u(m.w) >> 31);
    const float e = fadd(fabs_(m.w), u2f(f2u(rb.rho) & sphere_mask));
    const float tcx = fma(m.x, rc.inv.x, fneg(rc.od.x));
    const float tcy = fma(m.y, rc.inv.y, fneg(rc.od.y));
    const float tcz = fma(m.z, rc.inv.z, fneg(rc.od.z));
    const float tn = fmax(fmax(fma(fneg(e), rc.ainv.x, tcx), fma(fneg(e), rc.ainv.y, tcy)), fmax(fma(fneg(e), rc.ainv.z, tcz), 0.0f));
    const float tf = fmin(fmin(fma(e, rc.ainv.x, tcx), fma(e, rc.ainv.y, tcy)), fmin(fma(e, rc.ainv.z, tcz), best));
    return !(tn > tf);
}

// (not inlined: the compiler otherwise duplicates these ~4 KB per call site, and the hot loop must fit the
// instruction cache -- see DESIGN.md 4.5)
RDR_HD_NOINLINE Hit cluster_exact(const SceneView &S, const RayCull rc, uint32_t *scratch, uint32_t stride, uint32_t nq, v3 o, v3 d,
                                  Hit best, TraceStats *stats)
{
    RDR_NOUNROLL
    for (uint32_t q = 0; q < nq; ++q) {                     // spheres
        const uint32_t slot = scratch[(CL_QBASE + q) * stride];
        const uint32_t tag = S.member_idx[slot];
        if (tag & 0x40000000u) continue;
        const f4 g = S.member_geom[slot];
        if (!rc.degenerate && !sphere_may_hit(o, d, rc, g.x, g.y, g.z, fmul(g.w, g.w))) continue;
        RDR_STAT(stats, sphere_exact);
        float t;
        if (hit_sphere_exact(o, d, mk3(g.x, g.y, g.z), g.w, &t)) {
            const int idx = (int)(tag & 0x3fffffffu);
            if (hit_better(t, idx, best.t, best.idx)) { best.idx = idx; best.t = t; }
        }
    }
    RDR_NOUNROLL
    for (uint32_t q = 0; q < nq; ++q) {                     // cubes
        const uint32_t slot = scratch[(CL_QBASE + q) * stride];
        const uint32_t tag = S.member_idx[slot];
        if (!(tag & 0x40000000u)) continue;
        const f4 g = S.member_geom[slot];
        RDR_STAT(stats, cube_exact);
        float t;
        if (hit_cube_exact(o, d, mk3(g.x, g.y, g.z), g.w, &t)) {
            const int idx = (int)(tag & 0x3fffffffu);
            if (hit_better(t, idx, best.t, best.idx)) { best.idx = idx; best.t = t; }
        }
    }
    return best;
}

RDR_HD Hit trace_cluster(const SceneView &S, const CullConsts &cc, uint32_t *scratch, uint32_t stride, v3 o, v3 d,
                         TraceStats *stats = nullptr)
{
    Hit best; best.idx = -1; best.t = finf();
    RDR_STAT(stats, traces);
    if (S.n_top == 0u) return best;
    const RayBvh rb = make_ray_bvh(o, d, cc);
    // origin outside the scene bound / non-finite ray: no culling, every member goes through the exact tests
    const bool all = rb.rc.degenerate;
    if (all) RDR_STAT(stats, degenerate);
    // ---- A0: uniform scan over the cluster boxes ----
    for (uint32_t ch = 0; ch < S.nt_chunks; ++ch) {
        uint32_t m = 0u;
        const f4 *p = S.top + ch * 64u;
        const uint32_t left = S.n_top - ch * 32u;
        const int groups = left >= 32u ? 4 : (int)((left + 7u) >> 3);
        RDR_NOUNROLL
        for (int g = 0; g < groups; ++g) {
            uint32_t mm = 0u;
            RDR_UNROLL4
            for (int j = 0; j < 8; ++j) {
                float tn;
                if (bvh_entry_may_hit(rb, p[2 * (g * 8 + j)], p[2 * (g * 8 + j) + 1], finf(), &tn)) mm |= (1u << j);
            }
            m |= mm << (g * 8);
        }
        if (all) m = left >= 32u ? 0xffffffffu : (1u << left) - 1u;
        scratch[ch * stride] = m;
    }
    // ---- A1: members of the lane's clusters into the queue;  B: exact tests whenever it fills, and at the end ----
    uint32_t nq = 0u;
    float prune = finf();
    uint32_t ch = 0u, m = scratch[0];
    {   // single-primitive top entries (a floor cube, ...) were tested in A0: straight to the queue (at most 8 here,
        // any further ones go through the member loop below like clusters of one)
        uint32_t md = m & (S.n_direct >= 32u ? 0xffffffffu : (1u << S.n_direct) - 1u);
        RDR_NOUNROLL
        while (md != 0u && nq < 8u) {
            const int k = ffs32(md); md &= md - 1u; m &= ~(1u << k);
            scratch[(CL_QBASE + nq) * stride] = f2u(S.top[2 * k + 1].z) >> 4; ++nq;
            RDR_STAT(stats, entries_hit);
        }
    }
    bool more = true;
    while (more) {
        while (more && nq <= (uint32_t)(BVH_QCAP - 8)) {
            while (m == 0u && ++ch < S.nt_chunks) m = scratch[ch * stride];
            if (m == 0u) { more = false; break; }
            const int k = ffs32(m); m &= m - 1u;
            const uint32_t payload = f2u(S.top[2 * (ch * 32u + (uint32_t)k) + 1].z);
            const uint32_t first = payload >> 4, count = payload & 15u;
            RDR_STAT(stats, nodes_visited);
            const f4 *mb = S.member_box + first;
            RDR_UNROLL4
            for (uint32_t j = 0; j < 8u; ++j) {
                // a cluster of one (a large primitive on its own) was already tested as a top entry
                if (j < count && (count == 1u || all || member_may_hit(rb, mb[j], prune))) {
                    scratch[(CL_QBASE + nq) * stride] = first + j; ++nq; RDR_STAT(stats, entries_hit);
                }
            }
        }
        best = cluster_exact(S, rb.rc, scratch, stride, nq, o, d, best, stats);
        nq = 0u;
        if (best.idx >= 0 && !isnan_(best.t)) prune = best.t;
    }
    return best;
}

// nearest-hit dispatch of the kernels: MODE 0 = flat scan with cull, 1 = flat scan exact-everything, 2 = BVH,
// 3 = two-level cluster scan
template <int MODE>
RDR_HD Hit trace_any(const SceneView &S, const CullConsts &cc, uint32_t *scratch, uint32_t stride, v3 o, v3 d,
                     TraceStats *stats = nullptr)
{
    if (MODE == 7) return trace_bvh(S, cc, scratch, stride, o, d, stats);       // per-lane twin of the cooperative traversal
    if (MODE >= 3) return trace_cluster(S, cc, scratch, stride, o, d, stats);   // 4, 5: per-lane twin of the cooperative / fused scan
    if (MODE == 2) return trace_bvh(S, cc, scratch, stride, o, d, stats);
    if (MODE == 1) return trace_brute<false>(S, cc, scratch, stride, o, d, stats);
    return trace_brute<true>(S, cc, scratch, stride, o, d, stats);
}

// order-preserving 64-bit key of a hit for the warp-cooperative searches, which fold a ray's candidates with atomicMin:
// (t, original index) ascending = the first-minimum rule of trace_ray (cpu.rs:344-352).  t is -0, >= +0 or NaN
// (cpu.rs:54-58,90-97); -0 == +0 for the comparison (bit 0 remembers the sign), NaN last.
RDR_HD unsigned long long coop_key(float t, int idx)
{
    const uint32_t bits = f2u(t);
    const uint32_t kt = isnan_(t) ? 0x7fc00000u : (bits & 0x7fffffffu);
    const uint32_t zflag = (bits == 0x80000000u) ? 1u : 0u;
    return ((unsigned long long)kt << 32) | ((unsigned long long)(uint32_t)idx << 1) | zflag;
}

// 0, but opaque to the compiler's uniformity analysis (see LaneState)
RDR_HD uint32_t lane_varying_zero(uint32_t *scratch)
{
    volatile uint32_t *v = scratch;
    *v = 0u;
    return *v;
}

// ---- the sample loop of one pixel ----------------------------------------------------------------------
// Loop structure ("sample refill"): the reference nests samples > pixels > bounces.  A lane here owns a
// pixel and alternates two phases until the pixel's samples are used up:
//   trace       nearest hit of the lane's current ray (the first one is the pixel's primary ray);
//   lane_miss / lane_shade_hit
//               a miss ends the path with the sky term and the lane restarts from the cached primary hit BEFORE the
//               shading code, so continuing paths and restarted ones go through the (expensive, exact-arithmetic)
//               shading code once per iteration, together; a path that uses up max_bounces finishes its sample
//               after shading and needs a second pass (rare).
// The render kernel runs the phases in warp lock-step (one __any_sync per iteration is the reconvergence
// point) and hands a lane whose pixel is finished the next unclaimed pixel, so every lane enters every
// trace with a live ray and the scan -- >90 % of the instructions -- runs converged no matter how path
// lengths differ between pixels.
//
// The camera ray has no jitter (cpu.rs:199-202), so all samples of a pixel share the primary ray and its
// nearest hit: both are computed ONCE PER FRAME by primary_kernel (rdr_kernels.cu) into a per-pixel table
// (FrameParams::primary / primary_idx) and reused by every sample of every launch of the frame (bit-identical
// results).  A lane that claims a pixel loads 20 bytes instead of running the camera set-up (240 instructions of
// exact divisions) and a trace, and goes straight to shading the primary hit.
//
// Per-pixel accumulation order is sample-ascending, as in the reference, so a launch over samples
// [s0, s0+n) on top of an accumulator that already holds [0, s0) is bit-identical to one launch.
// The per-lane state that is touched only when a sample ends or starts -- the pixel's accumulator (4 words), the
// primary ray direction (3) and the cached primary hit (2) -- sits behind a storage policy: registers (ColdRegs: the
// host simulation and the per-lane kernels) or the lane's column of a shared-memory array (ColdShared, rdr_kernels.cu:
// 9 registers fewer in the hot loop of the fused kernel, paid with 9 LDS / 4 STS per finished sample).
enum { COLD_ACC = 0, COLD_CAM_D = 4, COLD_H0_IDX = 7, COLD_H0_T = 8, COLD_WORDS = 9,
       COLD_PARK = 9, COLD_PARK_WORDS = 9 };     // + the path state parked across a trace (ColdShared only)
struct ColdRegs {
    float w[COLD_WORDS];
    RDR_HD float get(int i) const { return w[i]; }
    RDR_HD void set(int i, float v) { w[i] = v; }
};

template <class COLD>
struct LaneStateT {
    COLD cold;              // accumulator, primary ray direction, cached primary hit
    Hit hit;                // hit to shade next
    v3 ro, rd, light, atten;
    uint32_t pixel, s, bounce, lane_zero;
    bool alive;             // the lane owns a pixel with samples left
    bool has_ray;           // ... and holds a ray to trace (otherwise st.hit waits to be shaded)

    RDR_HD f4 acc() const { f4 a; a.x = cold.get(COLD_ACC); a.y = cold.get(COLD_ACC + 1); a.z = cold.get(COLD_ACC + 2); a.w = cold.get(COLD_ACC + 3); return a; }
    RDR_HD void set_acc(f4 a) { cold.set(COLD_ACC, a.x); cold.set(COLD_ACC + 1, a.y); cold.set(COLD_ACC + 2, a.z); cold.set(COLD_ACC + 3, a.w); }
    RDR_HD v3 cam_d() const { return mk3(cold.get(COLD_CAM_D), cold.get(COLD_CAM_D + 1), cold.get(COLD_CAM_D + 2)); }
    RDR_HD void set_cam_d(v3 d) { cold.set(COLD_CAM_D, d.x); cold.set(COLD_CAM_D + 1, d.y); cold.set(COLD_CAM_D + 2, d.z); }
    RDR_HD Hit h0() const { Hit h; h.idx = (int)f2u(cold.get(COLD_H0_IDX)); h.t = cold.get(COLD_H0_T); return h; }
    RDR_HD void set_h0(Hit h) { cold.set(COLD_H0_IDX, u2f((uint32_t)h.idx)); cold.set(COLD_H0_T, h.t); }
    // the sample's light joins the accumulator (render_next_sample, cpu.rs:203-212: sum += sample, alpha += 1)
    RDR_HD void accumulate(v3 l)
    {
        f4 a = acc();
        a.x = fadd(a.x, l.x); a.y = fadd(a.y, l.y); a.z = fadd(a.z, l.z); a.w = fadd(a.w, 1.0f);
        set_acc(a);
    }
};
typedef LaneStateT<ColdRegs> LaneState;

// ptxas 12.9 (sm_100a) promotes a counter that starts from a constant and is stepped by a constant to a
// UNIFORM register even when lanes step it at different times (observed: the sample counter in UR4 with
// UIADD3/UISETP/BRA.U, all lanes of a warp sharing it -> too few samples per pixel).  Starting the per-lane
// counters from a value the compiler must treat as lane-varying (a volatile read-back of the lane's scratch
// word) keeps them in vector registers.  tests/test_gpu_parity.py::test_accumulator_bit_exact guards this.
template <class ST>
RDR_HD void lane_init(ST &st, uint32_t *masks)
{
    st.lane_zero = lane_varying_zero(masks);
    st.alive = false; st.has_ray = false;
    st.pixel = 0u; st.s = st.lane_zero; st.bounce = st.lane_zero;
    f4 z; z.x = z.y = z.z = z.w = 0.0f;
    st.set_acc(z);
    st.ro = st.rd = st.light = st.atten = mk3(0.0f, 0.0f, 0.0f);
    st.set_cam_d(st.ro);
    st.hit.idx = -1; st.hit.t = 0.0f; st.set_h0(st.hit);
}

// the sample's path ended with a miss; defined below
template <class ST> RDR_HD void lane_miss(const FrameParams &P, ST &st);

// take ownership of `pixel`: acc = its current accumulator, cam_d / h0 = its primary ray direction and nearest hit (the
// frame's primary table).  Afterwards either st.alive -- the primary hit waits in st.hit to be shaded (no ray to trace
// yet) -- or the pixel is already finished (no samples, no bounces, or a primary ray that misses: every sample is the
// sky) and st.acc() is final.
template <class ST>
RDR_HD void lane_start_pixel(const FrameParams &P, uint32_t pixel, f4 acc, v3 cam_d, Hit h0, ST &st)
{
    st.pixel = pixel;
    st.s = st.lane_zero; st.bounce = st.lane_zero;
    st.light = mk3(0.0f, 0.0f, 0.0f); st.atten = mk3(1.0f, 1.0f, 1.0f);
    st.ro = mk3(P.cam.pos[0], P.cam.pos[1], P.cam.pos[2]);
    st.has_ray = false;
    if (P.max_bounces == 0u) {                // `for _ in 0..0`: light stays zero, alpha still accumulates
        for (uint32_t s = 0; s < P.sample_count; ++s) {
            acc.x = fadd(acc.x, 0.0f); acc.y = fadd(acc.y, 0.0f); acc.z = fadd(acc.z, 0.0f); acc.w = fadd(acc.w, 1.0f);
        }
        st.set_acc(acc);
        st.alive = false;
        return;
    }
    st.set_acc(acc);
    st.alive = P.sample_count > 0u;
    if (!st.alive) return;
    st.rd = cam_d; st.set_cam_d(cam_d);
    st.hit = h0; st.set_h0(h0);
    if (h0.idx < 0) lane_miss(P, st);         // the primary ray misses: every sample is the sky, the pixel ends here
}

// the traced hit of the lane's current ray arrives
template <class ST>
RDR_HD void lane_accept_hit(ST &st, Hit h)
{
    st.hit = h;
    st.has_ray = false;
}

// the finished sample's light is in the accumulator: start the pixel's next sample from the cached primary hit
// (false: the pixel has no samples left)
template <class ST>
RDR_HD bool lane_next_sample(const FrameParams &P, ST &st)
{
    if (++st.s >= P.sample_count) { st.alive = false; return false; }
    st.bounce = st.lane_zero; st.hit = st.h0();
    st.ro = mk3(P.cam.pos[0], P.cam.pos[1], P.cam.pos[2]); st.rd = st.cam_d();
    st.light = mk3(0.0f, 0.0f, 0.0f); st.atten = mk3(1.0f, 1.0f, 1.0f);
    return true;
}

// The traced ray missed (cpu.rs:334-338): light += sky * attenuation, the sample is complete, and the lane
// restarts from the cached primary hit.  When the primary ray itself misses, every sample of the pixel is the
// sky and the pixel is finished here.  Afterwards either !st.alive (pixel done) or st.hit.idx >= 0 (to shade).
template <class ST>
RDR_HD void lane_miss(const FrameParams &P, ST &st)
{
    for (;;) {
        st.light = add3(st.light, mul3(world_sample(P.world, st.rd), st.atten));
        st.accumulate(st.light);
        if (!lane_next_sample(P, st)) return;
        if (st.hit.idx >= 0) return;
    }
}

// Shades st.hit (idx >= 0): closest_hit, scatter, throughput and emission (cpu.rs:257-333).  Returns true when the
// lane now holds a ray that needs tracing; false when the bounce budget is used up (cpu.rs:256,341: the sample keeps
// its emission, no sky term) -- then the sample is finished and either the pixel is done (!st.alive) or the lane
// has restarted from the primary hit, which needs shading again (rare: call once more).
template <class ST>
RDR_HD bool lane_shade_hit(const FrameParams &P, const SceneView &S, ST &st)
{
    bool is_sphere;
    const Material m = load_material(S, st.hit.idx, &is_sphere);
    const f4 g = S.obj_geom[st.hit.idx];
    const Surface sf = closest_hit(st.ro, st.rd, st.hit.t, is_sphere, mk3(g.x, g.y, g.z), g.w);
    const Scatter sc = scatter(st.rd, sf, m, P.seed_lo, P.seed_hi, st.pixel, P.sample_begin + st.s, st.bounce);
    st.ro = sc.origin; st.rd = sc.dir;
    st.atten = mul3(st.atten, m.albedo);
    st.light = add3(st.light, scale3(m.emission, m.emission_strength));
    ++st.bounce;
    if (st.bounce < P.max_bounces) { st.has_ray = true; return true; }
    st.accumulate(st.light);
    lane_next_sample(P, st);
    return false;
}

// scalar driver of the phases for one pixel (host simulation; the kernel drives them in warp lock-step)
template <int MODE>
RDR_HD f4 render_pixel(const FrameParams &P, const SceneView &S, uint32_t *masks, uint32_t stride, uint32_t pixel, f4 acc,
                       TraceStats *stats = nullptr)
{
    LaneState st;
    lane_init(st, masks);
    // the pixel's primary ray and its nearest hit (the kernel reads them from the frame's primary table)
    const v3 cam_d = camera_ray_dir(P.cam, pixel % P.cam.width, pixel / P.cam.width);
    Hit h0; h0.idx = -1; h0.t = 0.0f;
    if (P.max_bounces != 0u && P.sample_count != 0u)
        h0 = trace_any<MODE>(S, P.cull, masks, stride, mk3(P.cam.pos[0], P.cam.pos[1], P.cam.pos[2]), cam_d, stats);
    lane_start_pixel(P, pixel, acc, cam_d, h0, st);
    while (st.alive) {
        if (st.has_ray) {
            lane_accept_hit(st, trace_any<MODE>(S, P.cull, masks, stride, st.ro, st.rd, stats));
            if (st.hit.idx < 0) lane_miss(P, st);
        }
        while (st.alive && !lane_shade_hit(P, S, st)) {}
    }
    return st.acc();
}

// one path with every bounce recorded (debug / parity); returns the number of steps taken
template <int MODE>
RDR_HD uint32_t trace_path_lane(const FrameParams &P, const SceneView &S, uint32_t *masks, uint32_t stride,
                                uint32_t x, uint32_t y, uint32_t sample, RdrPathStep *steps, uint32_t capacity, float rgba[4])
{
    const uint32_t pixel = y * P.cam.width + x;
    v3 ro = mk3(P.cam.pos[0], P.cam.pos[1], P.cam.pos[2]);
    v3 rd = camera_ray_dir(P.cam, x, y);
    v3 light = mk3(0.0f, 0.0f, 0.0f), atten = mk3(1.0f, 1.0f, 1.0f);
    uint32_t written = 0u;
    for (uint32_t bounce = 0; bounce < P.max_bounces; ++bounce) {
        const Hit hit = trace_any<MODE>(S, P.cull, masks, stride, ro, rd);
        RdrPathStep st;
        memset(&st, 0, sizeof st);
        if (hit.idx >= 0) {
            bool is_sphere;
            const Material m = load_material(S, hit.idx, &is_sphere);
            const f4 g = S.obj_geom[hit.idx];
            const Surface sf = closest_hit(ro, rd, hit.t, is_sphere, mk3(g.x, g.y, g.z), g.w);
            const Scatter sc = scatter(rd, sf, m, P.seed_lo, P.seed_hi, pixel, sample, bounce);
            ro = sc.origin; rd = sc.dir;
            atten = mul3(atten, m.albedo);
            light = add3(light, scale3(m.emission, m.emission_strength));
            st.object = hit.idx; st.lobe = sc.lobe; st.front_face = sf.front ? 1u : 0u; st.t = hit.t;
            st.position[0] = sf.p.x; st.position[1] = sf.p.y; st.position[2] = sf.p.z;
            st.normal[0] = sf.n.x; st.normal[1] = sf.n.y; st.normal[2] = sf.n.z;
        } else {
            light = add3(light, mul3(world_sample(P.world, rd), atten));
            st.object = -1; st.lobe = 0u;
        }
        st.origin[0] = ro.x; st.origin[1] = ro.y; st.origin[2] = ro.z;
        st.direction[0] = rd.x; st.direction[1] = rd.y; st.direction[2] = rd.z;
        st.attenuation[0] = atten.x; st.attenuation[1] = atten.y; st.attenuation[2] = atten.z;
        st.light[0] = light.x; st.light[1] = light.y; st.light[2] = light.z;
        if (written < capacity) steps[written] = st;
        ++written;
        if (hit.idx < 0) break;
    }
    rgba[0] = light.x; rgba[1] = light.y; rgba[2] = light.z; rgba[3] = 1.0f;
    return written < capacity ? written : capacity;
}

}  // namespace rdr

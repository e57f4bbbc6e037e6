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
struct TraceStats { uint64_t traces, sphere_exact, cube_exact, degenerate; };
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
#if defined(__CUDA_ARCH__)
#pragma unroll 8
#endif
            for (int j = 0; j < 32; ++j) {
                const f4 s = p[j];
                if (sphere_may_hit(o, d, rc, s.x, s.y, s.z, s.w)) m |= (1u << j);
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
#if defined(__CUDA_ARCH__)
#pragma unroll 8
#endif
                for (int j = 0; j < 32; ++j) {
                    const f4 c = p[j];
                    if (cube_may_hit(rc, c.x, c.y, c.z, c.w, prune)) m |= (1u << j);
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

// 0, but opaque to the compiler's uniformity analysis (see render_pixel)
RDR_HD uint32_t lane_varying_zero(uint32_t *scratch)
{
    volatile uint32_t *v = scratch;
    *v = 0u;
    return *v;
}

// ---- the sample loop of one pixel ----------------------------------------------------------------------
// Loop structure ("sample refill"): the reference nests samples > pixels > bounces.  A lane here owns a
// pixel and runs ONE loop whose body is "trace the lane's current ray, then shade".  When a path ends
// (miss, or max_bounces traces used) the lane immediately starts its pixel's next sample inside the
// same iteration, so every lane enters every trace with a live ray and the scan -- >90 % of the
// work -- runs with all 32 lanes active regardless of how path lengths differ.
//
// The camera ray has no jitter (cpu.rs:199-202), so all samples of a pixel share the primary ray and
// its nearest hit; it is traced once per launch and reused (bit-identical results).
//
// Per-pixel accumulation order is sample-ascending, as in the reference, so a launch over samples
// [s0, s0+n) on top of an accumulator that already holds [0, s0) is bit-identical to one launch.
template <bool USE_CULL>
RDR_HD f4 render_pixel(const FrameParams &P, const SceneView &S, uint32_t *masks, uint32_t stride, uint32_t pixel, f4 acc,
                       TraceStats *stats = nullptr)
{
    const uint32_t n = P.sample_count;
    if (P.max_bounces == 0u) {            // `for _ in 0..0`: light stays zero, alpha still accumulates
        for (uint32_t s = 0; s < n; ++s) { acc.x = fadd(acc.x, 0.0f); acc.y = fadd(acc.y, 0.0f); acc.z = fadd(acc.z, 0.0f); acc.w = fadd(acc.w, 1.0f); }
        return acc;
    }
    const uint32_t x = pixel % P.cam.width, y = pixel / P.cam.width;
    const v3 cam_o = mk3(P.cam.pos[0], P.cam.pos[1], P.cam.pos[2]);
    const v3 cam_d = camera_ray_dir(P.cam, x, y);
    const Hit h0 = trace_brute<USE_CULL>(S, P.cull, masks, stride, cam_o, cam_d, stats);

    // ptxas 12.9 (sm_100a) promotes a loop counter that starts from a constant and is stepped by a constant
    // to a UNIFORM register even when, as here, lanes step it at different times (observed: `s` in UR4,
    // UIADD3/UISETP/BRA.U, every lane of a warp sharing one sample counter -> too few samples per pixel).
    // Starting the per-lane counters from a value the compiler must treat as lane-varying (a volatile
    // shared-memory read-back of this lane's scratch word) keeps them in vector registers.
    // tests/test_gpu_parity.py::test_accumulator_bit_exact guards this.
    const uint32_t lane_zero = lane_varying_zero(masks);
    uint32_t s = lane_zero, bounce = lane_zero;
    v3 ro = cam_o, rd = cam_d;
    v3 light = mk3(0.0f, 0.0f, 0.0f), atten = mk3(1.0f, 1.0f, 1.0f);
    Hit hit = h0;
    bool alive = n > 0u;
    while (alive) {
        for (;;) {                         // shade; on termination start the next sample and shade its primary hit
            bool terminated;
            if (hit.idx >= 0) {
                bool is_sphere;
                const Material m = load_material(S, hit.idx, &is_sphere);
                const f4 g = S.obj_geom[hit.idx];
                const Surface sf = closest_hit(ro, rd, hit.t, is_sphere, mk3(g.x, g.y, g.z), g.w);
                const Scatter sc = scatter(rd, sf, m, P.seed_lo, P.seed_hi, pixel, P.sample_begin + s, bounce);
                ro = sc.origin; rd = sc.dir;
                atten = mul3(atten, m.albedo);
                light = add3(light, scale3(m.emission, m.emission_strength));
                ++bounce;
                terminated = bounce >= P.max_bounces;
            } else {
                light = add3(light, mul3(world_sample(P.world, rd), atten));
                terminated = true;
            }
            if (!terminated) break;
            acc.x = fadd(acc.x, light.x); acc.y = fadd(acc.y, light.y); acc.z = fadd(acc.z, light.z); acc.w = fadd(acc.w, 1.0f);
            if (++s >= n) { alive = false; break; }
            bounce = lane_zero; ro = cam_o; rd = cam_d; hit = h0;
            light = mk3(0.0f, 0.0f, 0.0f); atten = mk3(1.0f, 1.0f, 1.0f);
        }
        if (!alive) break;
        hit = trace_brute<USE_CULL>(S, P.cull, masks, stride, ro, rd, stats);
    }
    return acc;
}

// one path with every bounce recorded (debug / parity); returns the number of steps taken
template <bool USE_CULL>
RDR_HD uint32_t trace_path_lane(const FrameParams &P, const SceneView &S, uint32_t *masks, uint32_t stride,
                                uint32_t x, uint32_t y, uint32_t sample, RdrPathStep *steps, uint32_t capacity, float rgba[4])
{
    const uint32_t pixel = y * P.cam.width + x;
    v3 ro = mk3(P.cam.pos[0], P.cam.pos[1], P.cam.pos[2]);
    v3 rd = camera_ray_dir(P.cam, x, y);
    v3 light = mk3(0.0f, 0.0f, 0.0f), atten = mk3(1.0f, 1.0f, 1.0f);
    uint32_t written = 0u;
    for (uint32_t bounce = 0; bounce < P.max_bounces; ++bounce) {
        const Hit hit = trace_brute<USE_CULL>(S, P.cull, masks, stride, ro, rd);
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

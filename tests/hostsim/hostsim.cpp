// hostsim.cpp -- TEST INFRASTRUCTURE: compiles the product's __host__ __device__ per-lane code
// (raydar_b200/csrc/rdr_core.cuh, rdr_trace.cuh, rdr_pack.h) for the CPU so that the logic the
// CUDA kernels run -- conservative cull + exact test, winner selection, sample-refill loop,
// scatter -- can be checked against the oracle on a machine without a GPU.
// Built by tests/conftest.py with g++ -O2 -ffp-contract=off -mfma (explicit fma() calls map to one
// hardware FMA exactly as FFMA does on the device; nothing else may be contracted).
// libraydar_cuda.so never contains or calls this file.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "raydar_cuda.h"
#include "rdr_pack.h"
#include "rdr_trace.cuh"

// the device-only fused scan, compiled for the CPU against the warp emulator (32 fibers = one warp)
#define RDR_WARP_EMU 1
#include "warp_emu.h"
#include "rdr_fused.cuh"
#include "rdr_device.cuh"
#include "rdr_bvh2.cuh"
#include "rdr_loop.cuh"

using namespace rdr;

namespace {

struct Packed {
    std::vector<unsigned char> storage;
    unsigned char *blob = nullptr;
    FrameParams P{};
    SceneView S{};
    std::vector<uint32_t> masks;
    int status = RDR_OK;
    explicit Packed(const RdrSceneFlat *sc, bool use_bvh = false)
    {
        std::vector<unsigned char> tmp;
        std::string err;
        status = pack_scene_blob(sc, use_bvh, tmp, P, err);
        if (status != RDR_OK) return;
        storage.resize(tmp.size() + 16);
        blob = storage.data();
        while (reinterpret_cast<uintptr_t>(blob) % 16) ++blob;
        memcpy(blob, tmp.data(), tmp.size());
        P.blob = blob;
        S = scene_view(blob, P.lay);
        masks.resize(std::max<uint32_t>(std::max(P.lay.ns_pad, P.lay.nc_pad) / 32 + 1, CL_SCRATCH));
    }
};

// entry k of a pair-packed node (rdr_bvh.h: node2_quad) in the two-quad entry form of bvh_entry_may_hit
void node2_entry(const float *p, int k, f4 &q0, f4 &q1, uint32_t &payload)
{
    const int pair = k >> 1, h = k & 1;
    auto quad = [&](int i) { return p + 4 * node2_quad((uint32_t)i, (uint32_t)pair); };
    q0.x = quad(0)[0 + h]; q0.y = quad(0)[2 + h]; q0.z = quad(1)[0 + h]; q0.w = quad(1)[2 + h];
    q1.x = quad(2)[0 + h]; q1.y = quad(2)[2 + h]; q1.z = 0.0f;
    memcpy(&payload, &quad(3)[2 + h], 4);
    q1.w = (payload != 0xffffffffu && (payload & BVH2_RHO_BIT)) ? 1.0f : 0.0f;
    if (payload != 0xffffffffu) payload &= ~BVH2_RHO_BIT;
}

// rank of entry k in the node's near-to-far order for direction octant oct
uint32_t node2_rank(const float *p, int k, int oct)
{
    uint32_t r; memcpy(&r, p + 4 * node2_quad(3u, (uint32_t)(k >> 1)) + (k & 1), 4);
    return (r >> (3 * oct)) & 7u;
}

// Scalar walk of the pair-packed hierarchy of the warp-cooperative traversal (rdr_bvh2.cuh runs on the device only):
// checks what the host can check -- the builder's boxes, masks and payloads, and the root in the kernel parameters --
// with the same conservative entry test and the same exact tests / winner rule.
Hit trace_bvh2_scalar(const Packed &pk, v3 o, v3 d, TraceStats *st)
{
    const SceneView &S = pk.S;
    Hit best; best.idx = -1; best.t = finf();
    const RayBvh rb = make_ray_bvh(o, d, pk.P.cull);
    const bool all = rb.rc.degenerate;
    if (st) { st->traces++; if (all) st->degenerate++; }
    auto prune = [&]() { return (best.idx >= 0 && !isnan_(best.t)) ? best.t : finf(); };
    std::vector<uint32_t> stack;
    auto visit = [&](f4 q0, f4 q1, uint32_t payload, bool is_prim, bool is_cube) {
        float tn;
        if (!all && !bvh_entry_may_hit(rb, q0, q1, prune(), &tn)) return;
        if (is_prim) {
            // the header masks must agree with the payload bits
            if (!(payload & 0x80000000u) || ((payload & 0x40000000u) != 0u) != is_cube) { best.idx = -2; return; }
            bvh_exact_prim(S, payload, o, d, best, st);
        } else {
            if (payload & 0x80000000u) { best.idx = -2; return; }
            stack.push_back(payload);
        }
    };
    const TopParams &T = pk.P.top;
    for (uint32_t k = 0; k < pk.P.lay.bvh2_root; ++k) {
        const TopPair &tp = T.pair[k >> 1]; const int h = (int)(k & 1u);
        f4 q0, q1; q0.x = tp.cx[h]; q0.y = tp.cy[h]; q0.z = tp.cz[h]; q0.w = tp.ex[h];
        q1.x = tp.ey[h]; q1.y = tp.ez[h]; q1.z = 0.0f; q1.w = tp.sphere[h];
        visit(q0, q1, T.payload[k], (T.prim_mask >> k) & 1u, (T.cube_mask >> k) & 1u);
    }
    const float *nodes = reinterpret_cast<const float *>(pk.blob + pk.P.lay.off_nodes2);
    while (!stack.empty()) {
        const uint32_t node = stack.back(); stack.pop_back();
        if (st) st->nodes_visited++;
        if (node >= pk.P.lay.n_nodes2) { best.idx = -2; return best; }
        const float *p = nodes + 64 * (size_t)node;
        for (int k = 0; k < 8; ++k) {
            f4 q0, q1; uint32_t payload;
            node2_entry(p, k, q0, q1, payload);
            if (payload == 0xffffffffu) continue;
            visit(q0, q1, payload, (payload & 0x80000000u) != 0u, (payload & 0x40000000u) != 0u);
        }
    }
    return best;
}

Hit trace_mode(int use_cull, const Packed &pk, uint32_t *scratch, v3 o, v3 d, TraceStats *st)
{
    if (pk.P.lay.mode == 1u && use_cull == 4) return trace_bvh2_scalar(pk, o, d, st);
    if (pk.P.lay.mode == 1u) return trace_any<2>(pk.S, pk.P.cull, scratch, 1, o, d, st);
    if (use_cull == 3) return trace_any<3>(pk.S, pk.P.cull, scratch, 1, o, d, st);
    return use_cull ? trace_any<0>(pk.S, pk.P.cull, scratch, 1, o, d, st) : trace_any<1>(pk.S, pk.P.cull, scratch, 1, o, d, st);
}

void merge(TraceStats *dst, const TraceStats &src)
{
    if (!dst) return;
    dst->traces += src.traces; dst->sphere_exact += src.sphere_exact; dst->cube_exact += src.cube_exact; dst->degenerate += src.degenerate;
    dst->nodes_visited += src.nodes_visited; dst->entries_hit += src.entries_hit;
}

}  // namespace

extern "C" {

int hs_first_hit(const RdrSceneFlat *sc, int use_cull, int32_t *ids, float *ts, TraceStats *stats)
{
    Packed pk(sc, use_cull == 2 || use_cull == 4);
    if (pk.status != RDR_OK) return pk.status;
    const int64_t n = (int64_t)sc->width * sc->height;
    TraceStats total{};
#pragma omp parallel
    {
        std::vector<uint32_t> masks(pk.masks.size());
        TraceStats local{};
#pragma omp for schedule(dynamic, 4096)
        for (int64_t p = 0; p < n; ++p) {
            const v3 o = mk3(pk.P.cam.pos[0], pk.P.cam.pos[1], pk.P.cam.pos[2]);
            const v3 d = camera_ray_dir(pk.P.cam, (uint32_t)(p % sc->width), (uint32_t)(p / sc->width));
            const Hit h = trace_mode(use_cull, pk, masks.data(), o, d, &local);
            ids[p] = h.idx;
            ts[p] = h.idx >= 0 ? h.t : 0.0f;
        }
#pragma omp critical
        merge(&total, local);
    }
    if (stats) *stats = total;
    return RDR_OK;
}

int hs_trace(const RdrSceneFlat *sc, int use_cull, uint32_t n, const float *rays, int32_t *ids, float *ts, TraceStats *stats)
{
    Packed pk(sc, use_cull == 2 || use_cull == 4);
    if (pk.status != RDR_OK) return pk.status;
    TraceStats total{};
#pragma omp parallel
    {
        std::vector<uint32_t> masks(pk.masks.size());
        TraceStats local{};
#pragma omp for schedule(static)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            const v3 o = mk3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
            const v3 d = mk3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
            const Hit h = trace_mode(use_cull, pk, masks.data(), o, d, &local);
            ids[i] = h.idx;
            ts[i] = h.idx >= 0 ? h.t : 0.0f;
        }
#pragma omp critical
        merge(&total, local);
    }
    if (stats) *stats = total;
    return RDR_OK;
}

int hs_render(const RdrSceneFlat *sc, int use_cull, uint64_t seed, uint32_t sample_begin, uint32_t n_samples,
              uint32_t max_bounces, float *accum, TraceStats *stats)
{
    Packed pk(sc, use_cull == 2 || use_cull == 4);
    if (pk.status != RDR_OK) return pk.status;
    pk.P.seed_lo = (uint32_t)seed; pk.P.seed_hi = (uint32_t)(seed >> 32);
    pk.P.max_bounces = max_bounces; pk.P.sample_begin = sample_begin; pk.P.sample_count = n_samples;
    const int64_t n = (int64_t)sc->width * sc->height;
    TraceStats total{};
#pragma omp parallel
    {
        std::vector<uint32_t> masks(pk.masks.size());
        TraceStats local{};
#pragma omp for schedule(dynamic, 256)
        for (int64_t p = 0; p < n; ++p) {
            f4 acc; acc.x = accum[4 * p]; acc.y = accum[4 * p + 1]; acc.z = accum[4 * p + 2]; acc.w = accum[4 * p + 3];
            acc = pk.P.lay.mode == 1u ? render_pixel<2>(pk.P, pk.S, masks.data(), 1, (uint32_t)p, acc, &local)
                  : use_cull == 3  ? render_pixel<3>(pk.P, pk.S, masks.data(), 1, (uint32_t)p, acc, &local)
                  : use_cull       ? render_pixel<0>(pk.P, pk.S, masks.data(), 1, (uint32_t)p, acc, &local)
                                   : render_pixel<1>(pk.P, pk.S, masks.data(), 1, (uint32_t)p, acc, &local);
            accum[4 * p] = acc.x; accum[4 * p + 1] = acc.y; accum[4 * p + 2] = acc.z; accum[4 * p + 3] = acc.w;
        }
#pragma omp critical
        merge(&total, local);
    }
    if (stats) *stats = total;
    return RDR_OK;
}

// the render kernel's pixel hand-out for one row-stripe shard (rdr_set_row_stripes): k = 0 .. owned_pixels - 1 in the
// order the atomic counter deals them, mapped by the product's stripe_pixel().  Returns owned_pixels.
uint32_t hs_stripe_pixels(uint32_t width, uint32_t height, uint32_t rows, uint32_t index, uint32_t count, uint32_t *pixels)
{
    const uint32_t n = stripe_owned_pixels(width, height, rows, index, count);
    if (pixels) for (uint32_t k = 0; k < n; ++k) pixels[k] = stripe_pixel(width, rows, index, count, k);
    return n;
}

// hs_render restricted to one row-stripe shard: only the owned pixels of `accum` are touched
int hs_render_stripes(const RdrSceneFlat *sc, uint64_t seed, uint32_t sample_begin, uint32_t n_samples, uint32_t max_bounces,
                      uint32_t rows, uint32_t index, uint32_t count, float *accum)
{
    Packed pk(sc, false);
    if (pk.status != RDR_OK) return pk.status;
    pk.P.seed_lo = (uint32_t)seed; pk.P.seed_hi = (uint32_t)(seed >> 32);
    pk.P.max_bounces = max_bounces; pk.P.sample_begin = sample_begin; pk.P.sample_count = n_samples;
    const int64_t n = (int64_t)stripe_owned_pixels(sc->width, sc->height, rows, index, count);
#pragma omp parallel
    {
        std::vector<uint32_t> masks(pk.masks.size());
#pragma omp for schedule(dynamic, 256)
        for (int64_t k = 0; k < n; ++k) {
            const uint32_t p = stripe_pixel(sc->width, rows, index, count, (uint32_t)k);
            f4 acc; acc.x = accum[4 * p]; acc.y = accum[4 * p + 1]; acc.z = accum[4 * p + 2]; acc.w = accum[4 * p + 3];
            acc = render_pixel<0>(pk.P, pk.S, masks.data(), 1, p, acc, nullptr);
            accum[4 * p] = acc.x; accum[4 * p + 1] = acc.y; accum[4 * p + 2] = acc.z; accum[4 * p + 3] = acc.w;
        }
    }
    return RDR_OK;
}

int hs_trace_path(const RdrSceneFlat *sc, int use_cull, uint64_t seed, uint32_t x, uint32_t y, uint32_t sample,
                  uint32_t max_bounces, RdrPathStep *steps, uint32_t capacity, uint32_t *n_steps, float rgba[4])
{
    Packed pk(sc, use_cull == 2 || use_cull == 4);
    if (pk.status != RDR_OK) return pk.status;
    pk.P.seed_lo = (uint32_t)seed; pk.P.seed_hi = (uint32_t)(seed >> 32);
    pk.P.max_bounces = max_bounces;
    *n_steps = pk.P.lay.mode == 1u ? trace_path_lane<2>(pk.P, pk.S, pk.masks.data(), 1, x, y, sample, steps, capacity, rgba)
               : use_cull == 3    ? trace_path_lane<3>(pk.P, pk.S, pk.masks.data(), 1, x, y, sample, steps, capacity, rgba)
               : use_cull         ? trace_path_lane<0>(pk.P, pk.S, pk.masks.data(), 1, x, y, sample, steps, capacity, rgba)
                                  : trace_path_lane<1>(pk.P, pk.S, pk.masks.data(), 1, x, y, sample, steps, capacity, rgba);
    return RDR_OK;
}

void hs_resolve(const float *accum, uint64_t n_pixels, uint32_t divisor, uint8_t *rgba8)
{
    for (uint64_t i = 0; i < n_pixels * 4; ++i) rgba8[i] = (uint8_t)quantise(accum[i], (float)divisor);
}

// raw conservative tests, for the adversarial margin checks: may[i] = 1 if the cull keeps primitive i for ray i
void hs_sphere_cull_batch(uint32_t n, const float *rays, const float *spheres, float q_max, float origin_bound, int32_t *may, int32_t *degenerate)
{
    CullConsts cc{q_max, origin_bound, 0.0f};
    for (uint32_t i = 0; i < n; ++i) {
        const v3 o = mk3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), d = mk3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
        const RayCull rc = make_ray_cull(o, d, cc);
        degenerate[i] = rc.degenerate;
        const float *s = spheres + 4 * (size_t)i;
        may[i] = rc.degenerate || sphere_may_hit(o, d, rc, s[0], s[1], s[2], s[3] * s[3]);
    }
}

// cubes: prims (cx,cy,cz,side); hp = |side|/2 + pad as packed; best[i] = pruning bound (+inf for none)
void hs_cube_cull_batch(uint32_t n, const float *rays, const float *cubes, float pad, float origin_bound, const float *best,
                        int32_t *may, int32_t *degenerate)
{
    CullConsts cc{0.0f, origin_bound, 0.0f};
    for (uint32_t i = 0; i < n; ++i) {
        const v3 o = mk3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), d = mk3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
        const RayCull rc = make_ray_cull(o, d, cc);
        degenerate[i] = rc.degenerate;
        const float *c = cubes + 4 * (size_t)i;
        const float hp = (c[3] < 0 ? -c[3] : c[3]) * 0.5f + pad;
        may[i] = rc.degenerate || cube_may_hit(rc, c[0], c[1], c[2], hp, best[i]);
    }
}

void hs_exact_batch(int sphere, uint32_t n, const float *rays, const float *prims, float *t_out, int32_t *hit)
{
    for (uint32_t i = 0; i < n; ++i) {
        const v3 o = mk3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), d = mk3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
        const float *p = prims + 4 * (size_t)i;
        float t = 0.0f;
        const bool h = sphere ? hit_sphere_exact(o, d, mk3(p[0], p[1], p[2]), p[3], &t) : hit_cube_exact(o, d, mk3(p[0], p[1], p[2]), p[3], &t);
        hit[i] = h; t_out[i] = h ? t : 0.0f;
    }
}

void hs_rng_block(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t block, uint32_t out[4])
{
    const u4 r = rng_block((uint32_t)seed, (uint32_t)(seed >> 32), pixel, sample, bounce, block);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// pad / bounds the packer derives for a scene (so tests can feed the raw cull entry points consistently)
int hs_scene_consts(const RdrSceneFlat *sc, float *q_max, float *origin_bound, float *cube_pad)
{
    Packed pk(sc);
    if (pk.status != RDR_OK) return pk.status;
    *q_max = pk.P.cull.sphere_q_max; *origin_bound = pk.P.cull.origin_bound;
    *cube_pad = 0.0f;
    if (pk.P.lay.n_cubes && pk.P.lay.mode == 0u) {
        const f4 c = pk.S.cube_cull[0], g = pk.S.cube_geom[0];
        *cube_pad = c.w - (g.w < 0 ? -g.w : g.w) * 0.5f;
    }
    return RDR_OK;
}

// which layout a search request gets (pack_scene_for_accel, the decision rdr_new_frame takes): *mode 0 = scan lists, 1 = hierarchy
int hs_pack_mode(const RdrSceneFlat *sc, int accel, uint32_t *mode, uint32_t *n_top)
{
    std::vector<unsigned char> blob; FrameParams P{}; std::string err;
    const int st = pack_scene_for_accel(sc, accel, 1024u, blob, P, err);
    if (st != RDR_OK) return st;
    *mode = P.lay.mode; *n_top = P.lay.n_top;
    return RDR_OK;
}

// hierarchy shape of a scene (BVH mode): nodes, depth is checked by the builder
int hs_bvh_info(const RdrSceneFlat *sc, uint32_t *n_nodes, uint32_t *mode, uint32_t *blob_bytes)
{
    Packed pk(sc, true);
    if (pk.status != RDR_OK) return pk.status;
    *n_nodes = pk.P.lay.n_nodes; *mode = pk.P.lay.mode; *blob_bytes = pk.P.lay.blob_bytes;
    return RDR_OK;
}

// Simulation of the STACK DISCIPLINE of the warp-cooperative traversal (rdr_bvh2.cuh) for groups of 32 primary rays:
// counts node visits, exact tests and rounds for a given policy, so that the policy can be tuned without a GPU.
//   flush_at: survivors are exact-tested when a list holds >= flush_at entries (device: 32) or the stack is empty
//   out[0] = node visits per ray, out[1] = exact tests per ray, out[2] = rounds per warp, out[3] = mean tasks per round
int hs_bvh2_warp_sim(const RdrSceneFlat *sc, uint32_t flush_at, uint32_t order_children, double *out)
{
    Packed pk(sc, true);
    if (pk.status != RDR_OK) return pk.status;
    const SceneView &S = pk.S;
    const int64_t n = (int64_t)sc->width * sc->height;
    const float *nodes = reinterpret_cast<const float *>(pk.blob + pk.P.lay.off_nodes2);
    const TopParams &T = pk.P.top;
    double visits = 0, exacts = 0, rounds = 0, tasks_popped = 0, warps = 0;
    for (int64_t p0 = 0; p0 < n; p0 += 32) {
        const int nl = (int)std::min<int64_t>(32, n - p0);
        v3 o[32], d[32]; RayBvh rb[32]; Hit best[32];
        for (int l = 0; l < nl; ++l) {
            o[l] = mk3(pk.P.cam.pos[0], pk.P.cam.pos[1], pk.P.cam.pos[2]);
            d[l] = camera_ray_dir(pk.P.cam, (uint32_t)((p0 + l) % sc->width), (uint32_t)((p0 + l) / sc->width));
            rb[l] = make_ray_bvh(o[l], d[l], pk.P.cull); best[l].idx = -1; best[l].t = finf();
        }
        struct Task { uint32_t owner, node; };
        struct Surv { uint32_t owner, payload; };
        std::vector<Task> stack; std::vector<Surv> surv;
        auto prune = [&](int l) { return (best[l].idx >= 0 && !isnan_(best[l].t)) ? best[l].t : finf(); };
        auto flush = [&](bool all) {
            while (surv.size() >= flush_at || (all && !surv.empty())) {
                const size_t k = std::min<size_t>(32, surv.size());
                for (size_t i = surv.size() - k; i < surv.size(); ++i) { bvh_exact_prim(S, surv[i].payload, o[surv[i].owner], d[surv[i].owner], best[surv[i].owner], nullptr); exacts += 1; }
                surv.resize(surv.size() - k);
            }
        };
        struct Child { float tn; uint32_t payload; bool prim; };
        auto emit = [&](int l, std::vector<Child> &ch) {
            if (order_children) std::stable_sort(ch.begin(), ch.end(), [](const Child &a, const Child &b) { return a.tn > b.tn; });   // far first: near on top
            for (const Child &c : ch) { if (c.prim) surv.push_back({(uint32_t)l, c.payload}); else stack.push_back({(uint32_t)l, c.payload}); }
        };
        for (uint32_t c0 = 0; c0 < pk.P.lay.bvh2_root; c0 += 8) {
            for (int l = 0; l < nl; ++l) {
                std::vector<Child> ch;
                for (uint32_t k = c0; k < std::min(c0 + 8, pk.P.lay.bvh2_root); ++k) {
                    const TopPair &tp = T.pair[k >> 1]; const int h = (int)(k & 1u);
                    f4 q0, q1; q0.x = tp.cx[h]; q0.y = tp.cy[h]; q0.z = tp.cz[h]; q0.w = tp.ex[h];
                    q1.x = tp.ey[h]; q1.y = tp.ez[h]; q1.z = 0.0f; q1.w = tp.sphere[h];
                    float tn;
                    if (bvh_entry_may_hit(rb[l], q0, q1, finf(), &tn)) ch.push_back({tn, T.payload[k], ((T.prim_mask >> k) & 1u) != 0u});
                }
                emit(l, ch);
            }
            flush(false);
        }
        for (;;) {
            const bool last = stack.empty();
            if (!last) {
                size_t pop = (1024 > stack.size() ? 1024 - stack.size() : 0) / 7; pop = std::max<size_t>(1, std::min<size_t>(32, pop)); pop = std::min(pop, stack.size());
                std::vector<Task> popped(stack.end() - pop, stack.end());
                stack.resize(stack.size() - pop);
                rounds += 1; tasks_popped += (double)pop;
                for (size_t i = 0; i < pop; ++i) {                  // lane i takes stack[top - i]
                    const Task t = popped[pop - 1 - i];
                    visits += 1;
                    const float *pn = nodes + 64 * (size_t)t.node;
                    std::vector<Child> ch;
                    for (int k = 0; k < 8; ++k) {
                        f4 q0, q1; uint32_t payload;
                        node2_entry(pn, k, q0, q1, payload);
                        if (payload == 0xffffffffu) continue;
                        float tn;
                        if (bvh_entry_may_hit(rb[t.owner], q0, q1, prune((int)t.owner), &tn)) ch.push_back({tn, payload, (payload & 0x80000000u) != 0u});
                    }
                    emit((int)t.owner, ch);
                }
            }
            flush(last);
            if (last) break;
        }
        warps += 1;
    }
    out[0] = visits / (double)n; out[1] = exacts / (double)n; out[2] = rounds / warps; out[3] = tasks_popped / std::max(1.0, rounds);
    return RDR_OK;
}

// Alternative discipline for comparison: ONE STACK PER RAY (near child on top), each round serves up to `per_round`
// rays (round-robin over the rays that still have tasks), one task each; survivors are exact-tested when >= flush_at
// have accumulated.  out as in hs_bvh2_warp_sim.
int hs_bvh2_perray_sim(const RdrSceneFlat *sc, uint32_t flush_at, uint32_t per_round, double *out)
{
    Packed pk(sc, true);
    if (pk.status != RDR_OK) return pk.status;
    const SceneView &S = pk.S;
    const int64_t n = (int64_t)sc->width * sc->height;
    const float *nodes = reinterpret_cast<const float *>(pk.blob + pk.P.lay.off_nodes2);
    const TopParams &T = pk.P.top;
    double visits = 0, exacts = 0, rounds = 0, tasks_popped = 0, warps = 0, max_depth = 0;
    for (int64_t p0 = 0; p0 < n; p0 += 32) {
        const int nl = (int)std::min<int64_t>(32, n - p0);
        v3 o[32], d[32]; RayBvh rb[32]; Hit best[32];
        struct Child { float tn; uint32_t payload; bool prim; };
        std::vector<Child> stack[32];
        struct Surv { uint32_t owner, payload; };
        std::vector<Surv> surv;
        auto prune = [&](int l) { return (best[l].idx >= 0 && !isnan_(best[l].t)) ? best[l].t : finf(); };
        auto flush = [&](bool all) {
            while (surv.size() >= (flush_at & 0xffffu) || (all && !surv.empty())) {
                const size_t k = std::min<size_t>(32, surv.size());
                for (size_t i = surv.size() - k; i < surv.size(); ++i) { bvh_exact_prim(S, surv[i].payload, o[surv[i].owner], d[surv[i].owner], best[surv[i].owner], nullptr); exacts += 1; }
                surv.resize(surv.size() - k);
            }
        };
        bool sort_children = true;
        auto emit = [&](int l, std::vector<Child> &ch) {
            if (sort_children) std::stable_sort(ch.begin(), ch.end(), [](const Child &a, const Child &b) { return a.tn > b.tn; });   // far first
            for (const Child &c : ch) { if (c.prim) surv.push_back({(uint32_t)l, c.payload}); else stack[l].push_back(c); }
            max_depth = std::max(max_depth, (double)stack[l].size());
        };
        sort_children = (flush_at & 0x20000u) == 0u;                 // bit 17 of flush_at: root hits pushed in entry order
        for (int l = 0; l < nl; ++l) {
            o[l] = mk3(pk.P.cam.pos[0], pk.P.cam.pos[1], pk.P.cam.pos[2]);
            d[l] = camera_ray_dir(pk.P.cam, (uint32_t)((p0 + l) % sc->width), (uint32_t)((p0 + l) / sc->width));
            rb[l] = make_ray_bvh(o[l], d[l], pk.P.cull); best[l].idx = -1; best[l].t = finf();
            std::vector<Child> ch;
            for (uint32_t k = 0; k < pk.P.lay.bvh2_root; ++k) {
                const TopPair &tp = T.pair[k >> 1]; const int h = (int)(k & 1u);
                f4 q0, q1; q0.x = tp.cx[h]; q0.y = tp.cy[h]; q0.z = tp.cz[h]; q0.w = tp.ex[h];
                q1.x = tp.ey[h]; q1.y = tp.ez[h]; q1.z = 0.0f; q1.w = tp.sphere[h];
                float tn;
                if (bvh_entry_may_hit(rb[l], q0, q1, finf(), &tn)) {
                    if (flush_at & 0x80000u) {                      // bit 19: order the root by the builder's octant ranks, not by tn
                        const int oct = (d[l].x < 0.0f ? 1 : 0) | (d[l].y < 0.0f ? 2 : 0) | (d[l].z < 0.0f ? 4 : 0);
                        tn = (float)((T.rank8[k] >> (5 * oct)) & 31u);
                    }
                    ch.push_back({tn, T.payload[k], ((T.prim_mask >> k) & 1u) != 0u});
                }
            }
            emit(l, ch);
        }
        sort_children = (flush_at & 0x40000u) == 0u;                 // bit 18: node children pushed in entry order too
        flush(false);
        int next = 0;
        for (;;) {
            int served = 0;
            for (int step = 0; step < nl && served < (int)per_round; ++step) {
                const int l = (next + step) % nl;
                if (stack[l].empty()) continue;
                const Child t = stack[l].back(); stack[l].pop_back();
                ++served;
                if ((flush_at & 0x10000u) == 0u && t.tn > prune(l)) continue;   // stale entry (skipped when the stack keeps tn; bit 16 of flush_at: no tn kept)
                visits += 1;
                const float *pn = nodes + 64 * (size_t)t.payload;
                std::vector<Child> ch;
                for (int k = 0; k < 8; ++k) {
                    f4 q0, q1; uint32_t payload;
                    node2_entry(pn, k, q0, q1, payload);
                    if (payload == 0xffffffffu) continue;
                    float tn;
                    if (bvh_entry_may_hit(rb[l], q0, q1, prune(l), &tn)) {
                        if (flush_at & 0x100000u) {                 // bit 20: order the children by the builder's octant ranks
                            const int oct = (d[l].x < 0.0f ? 1 : 0) | (d[l].y < 0.0f ? 2 : 0) | (d[l].z < 0.0f ? 4 : 0);
                            tn = (float)node2_rank(pn, k, oct);
                        }
                        ch.push_back({tn, payload, (payload & 0x80000000u) != 0u});
                    }
                }
                emit(l, ch);
                if (served == (int)per_round) next = (l + 1) % nl;
            }
            const bool last = served == 0;
            if (!last) { rounds += 1; tasks_popped += served; }
            flush(last);
            if (last) break;
        }
        warps += 1;
    }
    out[0] = visits / (double)n; out[1] = exacts / (double)n; out[2] = rounds / warps; out[3] = tasks_popped / std::max(1.0, rounds);
    out[4] = max_depth;
    return RDR_OK;
}

// shape of the pair-packed hierarchy of the cooperative traversal: nodes, root entries
int hs_bvh2_info(const RdrSceneFlat *sc, uint32_t *n_nodes2, uint32_t *n_root, uint32_t *ok)
{
    Packed pk(sc, true);
    if (pk.status != RDR_OK) return pk.status;
    *n_nodes2 = pk.P.lay.n_nodes2; *n_root = pk.P.lay.bvh2_root; *ok = pk.P.lay.bvh2_ok;
    return RDR_OK;
}

int hs_cluster_info(const RdrSceneFlat *sc, uint32_t *n_top, uint32_t *blob_bytes)
{
    Packed pk(sc, false);
    if (pk.status != RDR_OK) return pk.status;
    *n_top = pk.P.lay.n_top; *blob_bytes = pk.P.lay.blob_bytes;
    return RDR_OK;
}

// The PRODUCT's warp-cooperative fused scan (trace_fused, rdr_fused.cuh: the search behind RDR_ACCEL_AUTO up to ~1000
// objects) on the CPU: rays are taken 32 at a time as the lanes of one emulated warp (warp_emu.h); lanes past n are dead
// (alive = false), as idle lanes are in the kernel.  Returns RDR_ERR_UNSUPPORTED when the scene has no fused layout and
// RDR_ERR_INVALID when the emulated warp deadlocks (a lane left a rendezvous the others still wait at).
int hs_trace_fused(const RdrSceneFlat *sc, uint32_t n, const float *rays, int32_t *ids, float *ts)
{
    Packed pk(sc, false);
    if (pk.status != RDR_OK) return pk.status;
    if (!pk.P.lay.fused_ok) return RDR_ERR_UNSUPPORTED;
    FusedView V;
    V.pair_block = pk.S.pair_block; V.member_geom = pk.S.fused_geom; V.member_idx = pk.S.fused_idx;
    const bool cap8 = pk.P.lay.fused_cap == 8u;
    const int64_t n_warps = ((int64_t)n + 31) / 32;
    int bad = 0;
#pragma omp parallel
    {
        warp_emu::Warp *W = new warp_emu::Warp();
        std::vector<unsigned long long> scratch(FUSED_WARP_BYTES / 8u + 1u);
#pragma omp for schedule(dynamic, 1)
        for (int64_t w = 0; w < n_warps; ++w) {
            const FusedWarp ws = fused_warp(reinterpret_cast<unsigned char *>(scratch.data()), 0u);
            Hit out[32];
            const bool ok = warp_emu::run_warp(*W, [&](int lane) {
                const int64_t i = w * 32 + lane;
                const bool alive = i < (int64_t)n;
                const v3 o = alive ? mk3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]) : mk3(0.0f, 0.0f, 0.0f);
                const v3 d = alive ? mk3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]) : mk3(0.0f, 0.0f, 1.0f);
                out[lane] = cap8 ? trace_fused<true>(V, pk.P, ws, alive, o, d) : trace_fused<false>(V, pk.P, ws, alive, o, d);
            });
            if (!ok) {
#pragma omp atomic write
                bad = 1;
                continue;
            }
            for (int lane = 0; lane < 32; ++lane) {
                const int64_t i = w * 32 + lane;
                if (i >= (int64_t)n) break;
                ids[i] = out[lane].idx;
                ts[i] = out[lane].idx >= 0 ? out[lane].t : 0.0f;
            }
        }
        delete W;
    }
    return bad ? RDR_ERR_INVALID : RDR_OK;
}

// The warp-cooperative cluster scan (trace_cluster_coop, rdr_device.cuh: what RDR_ACCEL_AUTO runs when a scan-packed scene
// has more than 32 top-level entries for the fused scan) under the warp emulator.
int hs_trace_coop(const RdrSceneFlat *sc, uint32_t n, const float *rays, int32_t *ids, float *ts)
{
    Packed pk(sc, false);
    if (pk.status != RDR_OK) return pk.status;
    const int64_t n_warps = ((int64_t)n + 31) / 32;
    int bad = 0;
#pragma omp parallel
    {
        warp_emu::Warp *W = new warp_emu::Warp();
        std::vector<unsigned long long> scratch(COOP_WARP_BYTES / 8u + 1u);
#pragma omp for schedule(dynamic, 1)
        for (int64_t w = 0; w < n_warps; ++w) {
            const CoopWarpScratch ws = coop_scratch(reinterpret_cast<unsigned char *>(scratch.data()), 0u);
            Hit out[32];
            const bool ok = warp_emu::run_warp(*W, [&](int lane) {
                const int64_t i = w * 32 + lane;
                const bool alive = i < (int64_t)n;
                const v3 o = alive ? mk3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]) : mk3(0.0f, 0.0f, 0.0f);
                const v3 d = alive ? mk3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]) : mk3(0.0f, 0.0f, 1.0f);
                out[lane] = trace_cluster_coop(pk.S, pk.P.cull, ws, alive, o, d);
            });
            if (!ok) {
#pragma omp atomic write
                bad = 1;
                continue;
            }
            for (int lane = 0; lane < 32; ++lane) {
                const int64_t i = w * 32 + lane;
                if (i >= (int64_t)n) break;
                ids[i] = out[lane].idx;
                ts[i] = out[lane].idx >= 0 ? out[lane].t : 0.0f;
            }
        }
        delete W;
    }
    return bad ? RDR_ERR_INVALID : RDR_OK;
}

// The warp-cooperative hierarchy (trace_bvh2, rdr_bvh2.cuh: RDR_ACCEL_AUTO above 1024 objects, BASELINE config 4) under
// the warp emulator, on a scene packed as a hierarchy.
int hs_trace_bvh2(const RdrSceneFlat *sc, uint32_t n, const float *rays, int32_t *ids, float *ts)
{
    Packed pk(sc, true);
    if (pk.status != RDR_OK) return pk.status;
    if (pk.P.lay.mode != 1u || !pk.P.lay.bvh2_ok) return RDR_ERR_UNSUPPORTED;
    const int64_t n_warps = ((int64_t)n + 31) / 32;
    int bad = 0;
#pragma omp parallel
    {
        warp_emu::Warp *W = new warp_emu::Warp();
        std::vector<unsigned long long> scratch(BVH2_WARP_BYTES / 8u + 1u);
#pragma omp for schedule(dynamic, 1)
        for (int64_t w = 0; w < n_warps; ++w) {
            const Bvh2Warp ws = bvh2_warp(reinterpret_cast<unsigned char *>(scratch.data()), 0u);
            Hit out[32];
            const bool ok = warp_emu::run_warp(*W, [&](int lane) {
                const int64_t i = w * 32 + lane;
                const bool alive = i < (int64_t)n;
                const v3 o = alive ? mk3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]) : mk3(0.0f, 0.0f, 0.0f);
                const v3 d = alive ? mk3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]) : mk3(0.0f, 0.0f, 1.0f);
                out[lane] = trace_bvh2(pk.S, pk.P, ws, alive, o, d);
            });
            if (!ok) {
#pragma omp atomic write
                bad = 1;
                continue;
            }
            for (int lane = 0; lane < 32; ++lane) {
                const int64_t i = w * 32 + lane;
                if (i >= (int64_t)n) break;
                ids[i] = out[lane].idx;
                ts[i] = out[lane].idx >= 0 ? out[lane].t : 0.0f;
            }
        }
        delete W;
    }
    return bad ? RDR_ERR_INVALID : RDR_OK;
}

// The render kernel's sample loop (rdr_loop_body.inc, the very text render_kernel<5 / 6, ...> compiles) around the fused
// scan, for one emulated CTA of EMU_WARPS warps that share the atomic pixel counter: cold = 1 keeps the cold / parked lane
// state in "shared-memory" columns (LaneStateT<ColdShared>), 0 in registers.  order: round-robin permutation of the warps
// (n_order entries, may be NULL).  chunk_samples (RDR_CHUNKED builds only, else ignored): samples per hand-out item.
}  // extern "C"

constexpr int EMU_WARPS = 4, EMU_BLOCK = 32 * EMU_WARPS;

template <bool COLD>
static bool emu_render(const Packed &pk, const FrameParams &P, const std::vector<int> &order)
{
    const SceneView &S = pk.S;
    FusedView V;
    V.pair_block = S.pair_block; V.member_geom = S.fused_geom; V.member_idx = S.fused_idx;
    const bool cap8 = P.lay.fused_cap == 8u;
    std::vector<unsigned long long> scratch((size_t)EMU_WARPS * FUSED_WARP_BYTES / 8u + 1u);
    std::vector<float> cold((size_t)EMU_BLOCK * (COLD_WORDS + COLD_PARK_WORDS), 0.0f);
    std::vector<uint32_t> lane_words(EMU_BLOCK, 0u);
    std::vector<warp_emu::Warp> warps(EMU_WARPS);
    return warp_emu::run_grid(warps, [&](int warp, int lane) {
        const FusedWarp ws = fused_warp(reinterpret_cast<unsigned char *>(scratch.data()), (uint32_t)warp);
        (void)lane;
#define RDR_LOOP_STATE typename lane_state_of<EMU_BLOCK, COLD>::type
#define RDR_LOOP_COLD_BASE (reinterpret_cast<unsigned char *>(cold.data()))
#define RDR_LOOP_LANE_SCRATCH (&lane_words[threadIdx.x])
#define RDR_LOOP_TRACE(alive, o, d) (cap8 ? trace_fused<true>(V, P, ws, alive, o, d) : trace_fused<false>(V, P, ws, alive, o, d))
#include "rdr_loop_body.inc"
#undef RDR_LOOP_STATE
#undef RDR_LOOP_COLD_BASE
#undef RDR_LOOP_LANE_SCRATCH
#undef RDR_LOOP_TRACE
    }, order);
}

extern "C" {

int hs_render_fused_emu(const RdrSceneFlat *sc, uint64_t seed, uint32_t sample_begin, uint32_t n_samples, uint32_t max_bounces,
                        int cold, uint32_t stripe_rows, uint32_t stripe_index, uint32_t stripe_count,
                        const int32_t *order, uint32_t n_order, uint32_t chunk_samples, uint32_t prior_samples, float *accum)
{
    Packed pk(sc, false);
    if (pk.status != RDR_OK) return pk.status;
    if (!pk.P.lay.fused_ok) return RDR_ERR_UNSUPPORTED;
    FrameParams P = pk.P;
    uint32_t counter = 0u;
    P.seed_lo = (uint32_t)seed; P.seed_hi = (uint32_t)(seed >> 32);
    P.max_bounces = max_bounces; P.sample_begin = sample_begin; P.sample_count = n_samples;
    P.accum = reinterpret_cast<f4 *>(accum);
    P.pixel_counter = &counter;
    P.stripe_rows = stripe_rows; P.stripe_index = stripe_index; P.stripe_count = stripe_count;
    P.owned_pixels = stripe_owned_pixels(P.cam.width, P.cam.height, stripe_rows, stripe_index, stripe_count);
    (void)chunk_samples; (void)prior_samples;
    if (n_samples == 0u || P.owned_pixels == 0u) return RDR_OK;              // launch_render skips empty launches
    // the frame's primary table (primary_kernel on the device): camera ray + nearest hit per pixel, by the per-lane scan
    const size_t n_pixels = (size_t)P.cam.width * P.cam.height;
    std::vector<f4> primary(n_pixels);
    std::vector<int32_t> primary_idx(n_pixels);
    {
        std::vector<uint32_t> masks(pk.masks.size());
        for (size_t p = 0; p < n_pixels; ++p) {
            const v3 d = camera_ray_dir(P.cam, (uint32_t)(p % P.cam.width), (uint32_t)(p / P.cam.width));
            const Hit h = trace_any<3>(pk.S, P.cull, masks.data(), 1, mk3(P.cam.pos[0], P.cam.pos[1], P.cam.pos[2]), d);
            primary[p].x = d.x; primary[p].y = d.y; primary[p].z = d.z; primary[p].w = h.idx >= 0 ? h.t : 0.0f;
            primary_idx[p] = h.idx;
        }
    }
    P.primary = primary.data(); P.primary_idx = primary_idx.data();
    std::vector<int> ord(order, order + (order ? n_order : 0u));
    const bool ok = cold ? emu_render<true>(pk, P, ord) : emu_render<false>(pk, P, ord);
    return ok ? RDR_OK : RDR_ERR_INVALID;
}

// fused clustering: cluster[i] = top-level entry that holds original object i (the fused scan's own clustering)
int hs_fused_clusters(const RdrSceneFlat *sc, int32_t *cluster)
{
    Packed pk(sc, false);
    if (pk.status != RDR_OK) return pk.status;
    const SceneLayout &L = pk.P.lay;
    if (!L.fused_ok) return RDR_ERR_UNSUPPORTED;
    for (uint32_t i = 0; i < L.n_objects; ++i) cluster[i] = -1;
    for (uint32_t k = 0; k < L.fused_top; ++k) {
        const f4 *blk = pk.S.pair_block + (size_t)L.fused_stride * k;
        const uint32_t desc = f2u(blk[2].z), count = desc & 63u, first = desc >> 12;
        for (uint32_t j = 0; j < count; ++j) cluster[pk.S.fused_idx[first + j]] = (int32_t)k;
    }
    return RDR_OK;
}

// work counters of the emulated fused scan since the last call (RDR_EMU_STATS builds; zeros otherwise), summed over the
// calling thread's warps only: run hs_trace_fused with OMP_NUM_THREADS=1 to get totals
void hs_fused_emu_stats(unsigned long long out[7], int reset)
{
#if defined(RDR_EMU_STATS)
    FusedEmuStats &s = fused_emu_stats();
    out[0] = s.traces; out[1] = s.tasks; out[2] = s.member_rounds; out[3] = s.sphere_rounds; out[4] = s.cube_rounds;
    out[5] = s.sphere_tests; out[6] = s.cube_tests;
    if (reset) s = FusedEmuStats{};
#else
    for (int i = 0; i < 7; ++i) out[i] = 0ull;
    (void)reset;
#endif
}

// the fused scan's layout figures: [fused_ok, fused_top, fused_cap, fused_direct, fused_ns_direct, fused_stage_bytes, blob_bytes]
int hs_fused_info(const RdrSceneFlat *sc, uint32_t out[7])
{
    Packed pk(sc, false);
    if (pk.status != RDR_OK) return pk.status;
    const SceneLayout &L = pk.P.lay;
    out[0] = L.fused_ok; out[1] = L.fused_top; out[2] = L.fused_cap; out[3] = L.fused_direct; out[4] = L.fused_ns_direct;
    out[5] = L.fused_stage_bytes; out[6] = L.blob_bytes;
    return RDR_OK;
}

}  // extern "C"

// rdr_pack.h -- host-side packing of a borrowed RdrSceneFlat into the blob layout of rdr_layout.h,
// plus the scene-wide constants of the conservative cull tests (rdr_core.cuh).  Pure host C++.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "raydar_cuda.h"
#include "rdr_bvh.h"
#include "rdr_layout.h"

#ifndef RDR_CLUSTER_REFINE
#define RDR_CLUSTER_REFINE 1
#endif

namespace rdr {

inline uint32_t round_up_u32(uint32_t v, uint32_t m) { return (v + m - 1u) / m * m; }

// returns RDR_OK or an RdrStatus with `err` set
// use_bvh: pack the 8-wide hierarchy instead of the scan lists
inline int pack_scene_blob(const RdrSceneFlat *sc, bool use_bvh, std::vector<unsigned char> &blob, FrameParams &P, std::string &err)
{
    if (!sc) { err = "scene is NULL"; return RDR_ERR_INVALID; }
    if (sc->n_objects && (!sc->kind || !sc->geom || !sc->material)) { err = "scene arrays are NULL"; return RDR_ERR_INVALID; }
    if (sc->world_kind == RDR_WORLD_TRANSPARENT) {
        err = "World::Transparent is not supported (todo!() in the reference, scene/world.rs:32)";
        return RDR_ERR_UNSUPPORTED;
    }
    if (sc->world_kind > RDR_WORLD_TRANSPARENT) { err = "unknown world kind"; return RDR_ERR_INVALID; }
    if ((uint64_t)sc->width * sc->height > 0x7fffffffull) { err = "resolution too large"; return RDR_ERR_INVALID; }

    const uint32_t n = sc->n_objects;
    std::vector<uint32_t> spheres, cubes;
    for (uint32_t i = 0; i < n; ++i) {
        if (sc->kind[i] == RDR_SPHERE) spheres.push_back(i);
        else if (sc->kind[i] == RDR_CUBE) cubes.push_back(i);
        else { err = "object " + std::to_string(i) + " has unknown kind " + std::to_string(sc->kind[i]); return RDR_ERR_INVALID; }
    }

    // scene bounds for the cull margins (rdr_core.cuh): |c|_inf + size over objects, camera position
    float obj_bound = 0.0f, q_max = 0.0f, r_min = INFINITY;
    for (uint32_t i = 0; i < n; ++i) {
        const float *g = sc->geom + 4 * (size_t)i;
        const float cm = std::max(std::fabs(g[0]), std::max(std::fabs(g[1]), std::fabs(g[2])));
        obj_bound = std::max(obj_bound, cm + std::fabs(g[3]));
        if (sc->kind[i] == RDR_SPHERE) {
            q_max = std::max(q_max, 2.0f * (g[0] * g[0] + g[1] * g[1] + g[2] * g[2]) + g[3] * g[3]);
            r_min = std::min(r_min, std::fabs(g[3]));
        }
    }
    if (!(r_min < INFINITY)) r_min = 0.0f;
    const float cam_bound = std::max(std::fabs(sc->cam_pos[0]), std::max(std::fabs(sc->cam_pos[1]), std::fabs(sc->cam_pos[2])));
    const float origin_bound = 1.001f * std::max(obj_bound, cam_bound) + 1e-3f;
    const float cube_pad = 3.814697265625e-06f * (origin_bound + obj_bound);   // 2^-18 * B

    SceneLayout L{};
    L.n_objects = n; L.n_spheres = (uint32_t)spheres.size(); L.n_cubes = (uint32_t)cubes.size();
    auto quad_at = [&](uint32_t base, uint32_t i) { return reinterpret_cast<float *>(blob.data() + base) + 4 * (size_t)i; };
    auto fill_objects = [&]() {
        for (uint32_t i = 0; i < n; ++i) {
            const float *g = sc->geom + 4 * (size_t)i;
            const float *m = sc->material + RDR_MAT_STRIDE * (size_t)i;
            float *og = quad_at(L.off_obj_geom, i);
            og[0] = g[0]; og[1] = g[1]; og[2] = g[2]; og[3] = g[3];
            float *mm = reinterpret_cast<float *>(blob.data() + L.off_material) + 12 * (size_t)i;
            mm[0] = m[0]; mm[1] = m[1]; mm[2] = m[2]; mm[3] = m[3];          // albedo, roughness
            mm[4] = m[5]; mm[5] = m[6]; mm[6] = m[7]; mm[7] = m[8];          // emission colour, strength
            mm[8] = m[4]; mm[9] = m[9]; mm[10] = m[10];                      // metallic, transmission, ior
            const uint32_t kind_bits = sc->kind[i];
            memcpy(&mm[11], &kind_bits, 4);
        }
    };

    bool bvh_done = false;
    if (use_bvh) {
        if (n > BVH_INDEX_MASK) { err = "too many objects"; return RDR_ERR_INVALID; }
        std::vector<BvhBuildPrim> prims(n);
        for (uint32_t i = 0; i < n; ++i) {
            const float *g = sc->geom + 4 * (size_t)i;
            BvhBuildPrim &p = prims[i];
            p.c[0] = g[0]; p.c[1] = g[1]; p.c[2] = g[2]; p.index = i; p.cube = sc->kind[i] == RDR_CUBE;
            p.e = p.cube ? std::fabs(g[3]) * 0.5f + cube_pad : std::fabs(g[3]) + 2.0f * cube_pad;
        }
        BvhBuilder builder;
        Bvh2Builder builder2;
        const bool ok2 = builder2.build(prims);
        if (builder.build(std::move(prims))) {
            L.mode = 1u;
            L.n_nodes = builder.n_nodes();
            L.bvh2_ok = (ok2 && builder2.n_nodes() <= 65535u) ? 1u : 0u;   // 16-bit node indices on the traversal stacks (rdr_bvh2.cuh)
            L.n_nodes2 = ok2 ? builder2.n_nodes() : 0u;
            L.bvh2_root = ok2 ? builder2.root.n : 0u;
            uint64_t off = 0;
            L.off_nodes = (uint32_t)off; off += 256ull * L.n_nodes;
            L.off_obj_geom = (uint32_t)off; off += 16ull * n;
            L.off_material = (uint32_t)off; off += 48ull * n;
            off = (off + 127ull) & ~127ull;                      // a node of the cooperative hierarchy = exactly two 128-byte lines
            L.off_nodes2 = (uint32_t)off; off += 256ull * L.n_nodes2;
            if (off > 0xfffffff0ull) { err = "scene too large"; return RDR_ERR_INVALID; }
            L.blob_bytes = std::max(16u, round_up_u32((uint32_t)off, 16u));
            blob.assign(L.blob_bytes, 0);
            memcpy(blob.data() + L.off_nodes, builder.nodes.data(), 256ull * L.n_nodes);
            if (L.n_nodes2) memcpy(blob.data() + L.off_nodes2, builder2.nodes.data(), 256ull * L.n_nodes2);
            if (ok2) {
                const Bvh2Root &R = builder2.root;
                for (uint32_t k = 0; k < FUSED_MAX_TOP; ++k) {
                    TopPair &tp = P.top.pair[k >> 1];
                    const int h = (int)(k & 1u);
                    tp.cx[h] = R.cx[k]; tp.cy[h] = R.cy[k]; tp.cz[h] = R.cz[k];
                    tp.ex[h] = R.ex[k]; tp.ey[h] = R.ey[k]; tp.ez[h] = R.ez[k];
                    tp.sphere[h] = R.sphere[k];
                    P.top.payload[k] = R.payload[k];
                    P.top.rank8[k] = R.rank8[k];
                    for (int oct = 0; oct < 8; ++oct) P.top.node_by_rank[oct][k] = R.node_by_rank[oct][k];
                    if (k < R.n && (R.payload[k] & BVH_PRIM_BIT)) {
                        P.top.prim_mask |= 1u << k;
                        if (R.payload[k] & BVH_CUBE_BIT) P.top.cube_mask |= 1u << k;
                    }
                }
            }
            fill_objects();
            bvh_done = true;
        }
        // a tree deeper than the traversal stack allows (pathological input) falls back to the scan lists
    }
    if (!bvh_done) {
        L.mode = 0u;
        L.ns_pad = round_up_u32(L.n_spheres, 32u); L.nc_pad = round_up_u32(L.n_cubes, 32u);
        // two-level cluster scan (rdr_layout.h)
        std::vector<BvhBuildPrim> cprims(n);
        for (uint32_t i = 0; i < n; ++i) {
            const float *g = sc->geom + 4 * (size_t)i;
            BvhBuildPrim &p = cprims[i];
            p.c[0] = g[0]; p.c[1] = g[1]; p.c[2] = g[2]; p.index = i; p.cube = sc->kind[i] == RDR_CUBE;
            p.e = p.cube ? std::fabs(g[3]) * 0.5f + cube_pad : std::fabs(g[3]) + 2.0f * cube_pad;
        }
        ClusterSet cs = build_clusters(cprims, 8u);
        // single-primitive entries first: the scan pushes them straight to the exact-test queue
        // ... spheres before cubes among them, and spheres before cubes inside every cluster (the fused scan splits
        // a cluster's survivor bits into a sphere and a cube part with one mask)
        for (auto &members : cs.clusters)
            std::stable_sort(members.begin(), members.end(), [&](uint32_t x, uint32_t y) { return !cprims[x].cube && cprims[y].cube; });
        std::stable_sort(cs.clusters.begin(), cs.clusters.end(),
                         [&](const std::vector<uint32_t> &x, const std::vector<uint32_t> &y) {
                             const int kx = x.size() == 1 ? (cprims[x[0]].cube ? 1 : 0) : 2, ky = y.size() == 1 ? (cprims[y[0]].cube ? 1 : 0) : 2;
                             return kx < ky;
                         });
        L.n_direct = 0;
        while (L.n_direct < cs.clusters.size() && L.n_direct < 4u && cs.clusters[L.n_direct].size() == 1) ++L.n_direct;
        L.n_top = (uint32_t)cs.clusters.size(); L.nt_pad = round_up_u32(L.n_top, 32u);
        if (L.nt_pad > 128u) { err = "too many objects for the shared-memory scan (limit 128 clusters): use RDR_ACCEL_BVH or AUTO"; return RDR_ERR_UNSUPPORTED; }
        L.n_members = 9u * L.n_top;           // 8 members + 1 pad quad per cluster: 144-byte stride spreads the banks
        // the fused scan's own clustering: the smallest cluster size (8, 16, 24, 32) that leaves <= 32 top entries
        ClusterSet fs;
        for (uint32_t cap = 8u; cap <= 32u; cap += 8u) {
            fs = build_clusters(cprims, cap);
            L.fused_cap = cap;
            if (fs.clusters.size() <= FUSED_MAX_TOP) break;
        }
        // Surface-area local search on the fused clustering (rdr_bvh.h: refine_clusters): same kernels, same winners --
        // the boxes only choose what gets tested -- but a quarter fewer (ray, cluster) tasks on benchmark.rscn.  It runs
        // for 8-member clusters (scenes up to ~250 objects: ~5 ms at 183 objects; 70 - 200 ms for the 24 / 32-member
        // clusterings of 500 - 900 objects, which is too slow for a new_frame: compile with -DRDR_CLUSTER_REFINE=2 for those;
        // 0 switches it off).  Measured on B200: 6,333 against 5,907 Msamples/s on benchmark.rscn (profiles/variants_r03a.txt).  The result only depends on the objects, so it is kept for the next
        // frame of the same geometry (a camera move, a re-render, the other devices of a multi-GPU handle).
        {
            const int mode = RDR_CLUSTER_REFINE;
            if (fs.clusters.size() <= FUSED_MAX_TOP && (mode >= 2 || (mode == 1 && L.fused_cap == 8u))) {
                struct RefineCache { std::vector<uint32_t> kind; std::vector<float> geom; uint32_t cap = 0; ClusterSet cs; };
                static thread_local RefineCache cache;
                const bool same = cache.cap == L.fused_cap && cache.kind.size() == n && cache.geom.size() == 4u * (size_t)n &&
                                  memcmp(cache.kind.data(), sc->kind, sizeof(uint32_t) * n) == 0 &&
                                  memcmp(cache.geom.data(), sc->geom, sizeof(float) * 4u * n) == 0;
                if (same) fs = cache.cs;
                else {
                    refine_clusters(fs, cprims, L.fused_cap);
                    cache.kind.assign(sc->kind, sc->kind + n); cache.geom.assign(sc->geom, sc->geom + 4u * (size_t)n);
                    cache.cap = L.fused_cap; cache.cs = fs;
                }
            }
        }
        L.fused_ok = (n > 0u && fs.clusters.size() <= FUSED_MAX_TOP) ? 1u : 0u;
        if (L.fused_ok) {
            for (auto &members : fs.clusters)
                std::stable_sort(members.begin(), members.end(), [&](uint32_t x, uint32_t y) { return !cprims[x].cube && cprims[y].cube; });
            std::stable_sort(fs.clusters.begin(), fs.clusters.end(),
                             [&](const std::vector<uint32_t> &x, const std::vector<uint32_t> &y) {
                                 const int kx = x.size() == 1 ? (cprims[x[0]].cube ? 1 : 0) : 2, ky = y.size() == 1 ? (cprims[y[0]].cube ? 1 : 0) : 2;
                                 return kx < ky;
                             });
            L.fused_top = (uint32_t)fs.clusters.size();
            while (L.fused_direct < L.fused_top && L.fused_direct < 4u && fs.clusters[L.fused_direct].size() == 1) {
                if (!cprims[fs.clusters[L.fused_direct][0]].cube) ++L.fused_ns_direct;
                ++L.fused_direct;
            }
            L.fused_stride = 3u * (L.fused_cap / 2u);
            if ((L.fused_stride & 1u) == 0u) ++L.fused_stride;
        }
        // Section order: what the shading code and the fused scan read comes first, so that the fused kernels stage
        // only that prefix (fused_stage_bytes) into shared memory; the flat-scan lists and the cluster scan's sections
        // follow (staged by the other kernels, which copy the whole blob).
        uint32_t off = 0;
        L.off_obj_geom = off;    off += 16u * n;
        L.off_material = off;    off += 48u * n;
        if (L.fused_ok) {
            L.off_pair_block = off; off += 16u * L.fused_stride * L.fused_top;
            L.off_fused_geom = off; off += 16u * L.fused_cap * L.fused_top;
            L.off_fused_idx = off;  off += 4u * L.fused_cap * L.fused_top;
        }
        off = round_up_u32(off, 16u);
        L.fused_stage_bytes = std::max(16u, off);
        L.off_sphere_cull = off; off += 16u * L.ns_pad;
        L.off_cube_cull = off;   off += 16u * L.nc_pad;
        L.off_sphere_geom = off; off += 16u * L.ns_pad;
        L.off_cube_geom = off;   off += 16u * L.nc_pad;
        L.off_sphere_idx = off;  off += 4u * L.ns_pad;
        L.off_cube_idx = off;    off += 4u * L.nc_pad;
        off = round_up_u32(off, 16u);
        L.off_top = off;         off += 32u * L.nt_pad;
        L.off_member_box = off;  off += 16u * L.n_members;
        L.off_member_geom = off; off += 16u * L.n_members;
        L.off_member_idx = off;  off += 4u * L.n_members;
        L.blob_bytes = std::max(16u, round_up_u32(off, 16u));

        blob.assign(L.blob_bytes, 0);
        for (uint32_t j = 0; j < L.n_spheres; ++j) {
            const float *g = sc->geom + 4 * (size_t)spheres[j];
            float *c = quad_at(L.off_sphere_cull, j), *e = quad_at(L.off_sphere_geom, j);
            c[0] = g[0]; c[1] = g[1]; c[2] = g[2]; c[3] = g[3] * g[3];
            e[0] = g[0]; e[1] = g[1]; e[2] = g[2]; e[3] = g[3];
            reinterpret_cast<uint32_t *>(blob.data() + L.off_sphere_idx)[j] = spheres[j];
        }
        for (uint32_t j = 0; j < L.n_cubes; ++j) {
            const float *g = sc->geom + 4 * (size_t)cubes[j];
            float *c = quad_at(L.off_cube_cull, j), *e = quad_at(L.off_cube_geom, j);
            c[0] = g[0]; c[1] = g[1]; c[2] = g[2]; c[3] = std::fabs(g[3]) * 0.5f + cube_pad;
            e[0] = g[0]; e[1] = g[1]; e[2] = g[2]; e[3] = g[3];
            reinterpret_cast<uint32_t *>(blob.data() + L.off_cube_idx)[j] = cubes[j];
        }
        for (uint32_t k = 0; k < L.nt_pad; ++k) {                      // unused top entries: e = -1 (always rejected)
            float *t = reinterpret_cast<float *>(blob.data() + L.off_top) + 8 * (size_t)k;
            t[3] = t[4] = t[5] = -1.0f;
        }
        for (uint32_t k = 0; k < FUSED_MAX_TOP / 2u; ++k)                // unused top pairs: e = -1
            for (int h = 0; h < 2; ++h) P.top.pair[k].ex[h] = P.top.pair[k].ey[h] = P.top.pair[k].ez[h] = -1.0f;
        for (uint32_t k = 0; k < L.n_top; ++k) {
            const std::vector<uint32_t> &members = cs.clusters[k];
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
            bool any_sphere = false;
            for (uint32_t j = 0; j < members.size(); ++j) {
                const BvhBuildPrim &p = cprims[members[j]];
                const float *g = sc->geom + 4 * (size_t)p.index;
                const uint32_t slot = 9u * k + j;
                float *mb = quad_at(L.off_member_box, slot), *mg = quad_at(L.off_member_geom, slot);
                mb[0] = p.c[0]; mb[1] = p.c[1]; mb[2] = p.c[2]; mb[3] = p.cube ? p.e : -p.e;
                mg[0] = g[0]; mg[1] = g[1]; mg[2] = g[2]; mg[3] = g[3];
                reinterpret_cast<uint32_t *>(blob.data() + L.off_member_idx)[slot] = p.index | (p.cube ? BVH_CUBE_BIT : 0u);
                for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p.c[a] - p.e); hi[a] = std::max(hi[a], p.c[a] + p.e); }
                any_sphere |= !p.cube;
            }
            float *t = reinterpret_cast<float *>(blob.data() + L.off_top) + 8 * (size_t)k;
            float e3[3];
            for (int a = 0; a < 3; ++a) {
                t[a] = 0.5f * (lo[a] + hi[a]);
                const float h = std::max(hi[a] - t[a], t[a] - lo[a]);
                e3[a] = std::nextafter(h * (1.0f + 1e-6f), INFINITY);
            }
            t[3] = e3[0]; t[4] = e3[1]; t[5] = e3[2];
            const uint32_t payload = ((9u * k) << 4) | (uint32_t)members.size();
            memcpy(&t[6], &payload, 4);
            t[7] = any_sphere ? 1.0f : 0.0f;
        }
        for (uint32_t k = 0; L.fused_ok && k < L.fused_top; ++k) {
            const std::vector<uint32_t> &members = fs.clusters[k];
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
            bool any_sphere = false;
            uint32_t n_sph = 0;
            float *blk = quad_at(L.off_pair_block, L.fused_stride * k);
            for (uint32_t j = 0; j < L.fused_cap; ++j) {
                float *q = blk + 12u * (j >> 1);
                const uint32_t hh = j & 1u;
                if (j >= members.size()) { q[6 + hh] = -1.0f; continue; }      // unused slot (also masked out by the member count)
                const BvhBuildPrim &p = cprims[members[j]];
                const float *g = sc->geom + 4 * (size_t)p.index;
                q[0 + hh] = p.c[0]; q[2 + hh] = p.c[1]; q[4 + hh] = p.c[2]; q[6 + hh] = p.e;
                q[8 + hh] = p.cube ? 0.0f : 1.0f;
                float *fg = quad_at(L.off_fused_geom, L.fused_cap * k + j);
                fg[0] = g[0]; fg[1] = g[1]; fg[2] = g[2]; fg[3] = g[3];
                reinterpret_cast<uint32_t *>(blob.data() + L.off_fused_idx)[L.fused_cap * k + j] = p.index;
                for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p.c[a] - p.e); hi[a] = std::max(hi[a], p.c[a] + p.e); }
                any_sphere |= !p.cube;
                if (!p.cube) ++n_sph;
            }
            const uint32_t desc = ((L.fused_cap * k) << 12) | (n_sph << 6) | (uint32_t)members.size();
            memcpy(&blk[10], &desc, 4);
            TopPair &tp = P.top.pair[k >> 1];
            const int h = (int)(k & 1u);
            float c3[3], e3[3];
            for (int a = 0; a < 3; ++a) {
                c3[a] = 0.5f * (lo[a] + hi[a]);
                const float hx = std::max(hi[a] - c3[a], c3[a] - lo[a]);
                e3[a] = std::nextafter(hx * (1.0f + 1e-6f), INFINITY);
            }
            tp.cx[h] = c3[0]; tp.cy[h] = c3[1]; tp.cz[h] = c3[2];
            tp.ex[h] = e3[0]; tp.ey[h] = e3[1]; tp.ez[h] = e3[2];
            tp.sphere[h] = any_sphere ? 1.0f : 0.0f;
        }
        fill_objects();
    }

    memcpy(P.cam.inv_proj, sc->inv_proj, sizeof P.cam.inv_proj);
    memcpy(P.cam.inv_view, sc->inv_view, sizeof P.cam.inv_view);
    memcpy(P.cam.pos, sc->cam_pos, sizeof P.cam.pos);
    P.cam.width = sc->width; P.cam.height = sc->height;
    P.world.kind = sc->world_kind;
    memcpy(P.world.a, sc->world_a, sizeof P.world.a);
    memcpy(P.world.b, sc->world_b, sizeof P.world.b);
    P.cull.sphere_q_max = q_max;
    P.cull.origin_bound = origin_bound;
    P.cull.sphere_r_min = r_min;
    P.lay = L;
    return RDR_OK;
}

// The layout for a requested search (RDR_ACCEL_*): the hierarchy for BVH / BVH_COOP and, in AUTO, above auto_threshold
// objects; the scan lists otherwise.  AUTO also takes the hierarchy when the scan's clustering does not fit its masks
// (more than 128 top-level entries: many primitives far larger than the median each stay alone) instead of failing.
inline int pack_scene_for_accel(const RdrSceneFlat *sc, int accel, uint32_t auto_threshold, std::vector<unsigned char> &blob,
                                FrameParams &P, std::string &err)
{
    const bool use_bvh = accel == RDR_ACCEL_BVH || accel == RDR_ACCEL_BVH_COOP ||
                         (accel == RDR_ACCEL_AUTO && sc && sc->n_objects > auto_threshold);
    int st = pack_scene_blob(sc, use_bvh, blob, P, err);
    if (st == RDR_ERR_UNSUPPORTED && !use_bvh && accel == RDR_ACCEL_AUTO && sc && sc->world_kind != RDR_WORLD_TRANSPARENT) {
        err.clear(); blob.clear(); P = FrameParams{};
        st = pack_scene_blob(sc, true, blob, P, err);
    }
    return st;
}

}  // namespace rdr

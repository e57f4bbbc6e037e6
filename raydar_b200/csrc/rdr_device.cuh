// rdr_device.cuh -- device-only helpers: TMA bulk staging of the scene blob into shared memory.
#pragma once

#include "rdr_trace.cuh"

namespace rdr {

#ifndef RDR_WARP_EMU        // the TMA staging is PTX; the CPU warp emulator (tests/hostsim) only needs the scan below
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// global blob -> shared memory with one cp.async.bulk (SASS UBLKCP) completing on an mbarrier;
// thread 0 issues, every thread waits on phase 0.  bytes is a multiple of 16, both sides 16-byte aligned.
__device__ __forceinline__ void stage_blob(unsigned char *dst, const unsigned char *src, uint32_t bytes, uint64_t *bar)
{
    const uint32_t bar_a = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(bar_a) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar_a), "r"(0u) : "memory");
    }
}

// dynamic shared memory: [blob][mbarrier (16 B)][scratch] (see scratch_base in rdr_kernels.cu)
__device__ __forceinline__ uint64_t *bar_ptr(unsigned char *smem, const SceneLayout &L)
{
    return reinterpret_cast<uint64_t *>(smem + L.blob_bytes);
}
__device__ __forceinline__ uint32_t *mask_base(unsigned char *smem, const SceneLayout &L)
{
    return reinterpret_cast<uint32_t *>(smem + L.blob_bytes + 16u);
}

#endif  // RDR_WARP_EMU

// ---- warp-cooperative cluster scan (MODE 4) ---------------------------------------------------------------------
// Same two-level structure and the same tests as trace_cluster (rdr_trace.cuh), but the per-lane, divergent
// stages are regrouped across the warp so that all 32 lanes stay busy whatever the rays do:
//   A0   every lane scans the cluster boxes for ITS ray (uniform loop, broadcast loads) -> per-lane cluster mask;
//   T    the (ray, cluster) pairs of all lanes are compacted into one task list (shuffle prefix sum over the
//        per-lane counts); lane L then takes tasks L, L+32, ... -- usually some OTHER lane's ray, fetched with
//        indexed shuffles -- and slab-tests the cluster's members; survivors go to a warp-wide list (shared
//        atomics);
//   E    the survivors are again dealt out evenly and get the exact, reference-ordered test (sphere pass, then
//        cube pass); a hit is folded into the owning ray's winner with a 64-bit atomicMin on the key
//        (t, original index), which is exactly the first-minimum rule of trace_ray (cpu.rs:344-352);
//   each lane finally reads back the winner of its own ray.
// In the per-lane version a warp waits for its busiest lane (measured: 12 of 32 lanes active in the member
// stage, 3-12 in the exact stage); here the work of a warp is spread evenly.
// Must be called by all 32 lanes of the warp (alive = false for lanes without a ray).
struct CoopWarpScratch {
    unsigned long long *best;     // [32] winner key per lane
    uint32_t *surv_s, *surv_c;    // [COOP_SURV_CAP] each: sphere / cube survivors, (owner lane << 20) | member slot
    uint16_t *tasks;              // [COOP_TASK_CAP] (owner lane << 8) | cluster index within the chunk
    uint32_t *count;              // [2] sphere, cube survivors
};
constexpr uint32_t COOP_TASK_CAP = 1024u;                 // 32 lanes x 32 clusters of one chunk
constexpr uint32_t COOP_SURV_CAP = 384u;                  // flushed when a list could overflow in the next round (+256)
constexpr uint32_t COOP_WARP_BYTES = 32u * 8u + 2u * COOP_SURV_CAP * 4u + COOP_TASK_CAP * 2u + 16u;

__device__ __forceinline__ CoopWarpScratch coop_scratch(unsigned char *base, uint32_t warp)
{
    unsigned char *p = base + (size_t)warp * COOP_WARP_BYTES;
    CoopWarpScratch w;
    w.best = reinterpret_cast<unsigned long long *>(p);
    w.surv_s = reinterpret_cast<uint32_t *>(p + 32u * 8u);
    w.surv_c = w.surv_s + COOP_SURV_CAP;
    w.tasks = reinterpret_cast<uint16_t *>(p + 32u * 8u + 2u * COOP_SURV_CAP * 4u);
    w.count = reinterpret_cast<uint32_t *>(p + 32u * 8u + 2u * COOP_SURV_CAP * 4u + COOP_TASK_CAP * 2u);
    return w;
}

// E: exact tests on the warp's survivor lists, dealt out evenly; a hit is folded into the owner's winner key
__device__ __forceinline__ void coop_exact(const SceneView &S, const CullConsts &cc, CoopWarpScratch ws, v3 o, v3 d, bool all)
{
    const uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_s = ws.count[0], n_c = ws.count[1];
    __syncwarp();
    if (lane == 0u) { ws.count[0] = 0u; ws.count[1] = 0u; }
#pragma unroll 1
    for (uint32_t s0 = 0u; s0 < n_s; s0 += 32u) {                     // spheres
        const uint32_t si = s0 + lane;
        const bool hs = si < n_s;
        const uint32_t e = hs ? ws.surv_s[si] : (lane << 20);
        const uint32_t own = e >> 20, slot = e & 0xfffffu;
        const v3 ro = mk3(__shfl_sync(FULL, o.x, own), __shfl_sync(FULL, o.y, own), __shfl_sync(FULL, o.z, own));
        const v3 rd = mk3(__shfl_sync(FULL, d.x, own), __shfl_sync(FULL, d.y, own), __shfl_sync(FULL, d.z, own));
        const bool own_all = __shfl_sync(FULL, (int)all, own) != 0;
        if (hs) {
            const f4 g = S.member_geom[slot];
            bool go = true;
            if (!own_all) {                                           // cheap conservative sphere test before the exact one
                RayCull rc2;
                sphere_margins(ro, rd, cc, rc2);
                go = rc2.degenerate || sphere_may_hit(ro, rd, rc2, g.x, g.y, g.z, fmul(g.w, g.w));
            }
            float tt;
            if (go && hit_sphere_exact(ro, rd, mk3(g.x, g.y, g.z), g.w, &tt))
                atomicMin(&ws.best[own], coop_key(tt, (int)(S.member_idx[slot] & 0x3fffffffu)));
        }
    }
#pragma unroll 1
    for (uint32_t s0 = 0u; s0 < n_c; s0 += 32u) {                     // cubes
        const uint32_t si = s0 + lane;
        const bool hs = si < n_c;
        const uint32_t e = hs ? ws.surv_c[si] : (lane << 20);
        const uint32_t own = e >> 20, slot = e & 0xfffffu;
        const v3 ro = mk3(__shfl_sync(FULL, o.x, own), __shfl_sync(FULL, o.y, own), __shfl_sync(FULL, o.z, own));
        const v3 rd = mk3(__shfl_sync(FULL, d.x, own), __shfl_sync(FULL, d.y, own), __shfl_sync(FULL, d.z, own));
        if (hs) {
            const f4 g = S.member_geom[slot];
            float tt;
            if (hit_cube_exact(ro, rd, mk3(g.x, g.y, g.z), g.w, &tt))
                atomicMin(&ws.best[own], coop_key(tt, (int)(S.member_idx[slot] & 0x3fffffffu)));
        }
    }
    __syncwarp();
}

__device__ __forceinline__ void coop_push(const SceneView &S, CoopWarpScratch ws, uint32_t owner, uint32_t slot)
{
    const bool cube = (S.member_idx[slot] & 0x40000000u) != 0u;
    const uint32_t pos = atomicAdd(&ws.count[cube ? 1 : 0], 1u);
    (cube ? ws.surv_c : ws.surv_s)[pos] = (owner << 20) | slot;
}

__device__ __forceinline__ Hit trace_cluster_coop(const SceneView &S, const CullConsts &cc, CoopWarpScratch ws, bool alive, v3 o, v3 d)
{
    const uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const RayBvh rb = make_ray_bvh(o, d, cc);
    // origin outside the scene bound / non-finite ray: no culling for this ray (every box test passes)
    const bool all = rb.rc.degenerate;
    const float rho = rb.rho;
    ws.best[lane] = ~0ull;
    if (lane == 0u) { ws.count[0] = 0u; ws.count[1] = 0u; }
    __syncwarp();

    for (uint32_t ch = 0; ch < S.nt_chunks; ++ch) {
        // ---- A0: this lane's ray against the cluster boxes of the chunk ----
        uint32_t m = 0u;
        const uint32_t left = S.n_top - ch * 32u;
        {
            const f4 *p = S.top + ch * 64u;
            const int groups = left >= 32u ? 4 : (int)((left + 7u) >> 3);
#pragma unroll 1
            for (int g = 0; g < groups; ++g) {
                uint32_t mm = 0u;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float tn;
                    if (bvh_entry_may_hit(rb, p[2 * (g * 8 + j)], p[2 * (g * 8 + j) + 1], finf(), &tn)) mm |= (1u << j);
                }
                m |= mm << (g * 8);
            }
            if (all) m = FULL;
            if (left < 32u) m &= (1u << left) - 1u;
            if (!alive) m = 0u;
        }
        // single-primitive top entries (at most 4, first chunk): their box was the entry -> straight to the survivors
        if (ch == 0u && S.n_direct != 0u) {
            uint32_t md = m & ((1u << S.n_direct) - 1u);
            m &= ~((1u << S.n_direct) - 1u);
            while (md != 0u) {
                const uint32_t k = (uint32_t)__ffs((int)md) - 1u; md &= md - 1u;
                coop_push(S, ws, lane, __float_as_uint(S.top[2 * k + 1].z) >> 4);
            }
        }
        // ---- T: compact the (ray, cluster) pairs of the warp into one task list ----
        const uint32_t cnt = __popc(m);
        uint32_t pre = cnt;
#pragma unroll
        for (uint32_t off = 1u; off < 32u; off <<= 1) {
            const uint32_t v = __shfl_up_sync(FULL, pre, off);
            if (lane >= off) pre += v;
        }
        const uint32_t total = __shfl_sync(FULL, pre, 31);
        {
            uint32_t pos = pre - cnt, mm = m;
            while (mm != 0u) {
                const uint32_t k = (uint32_t)__ffs((int)mm) - 1u; mm &= mm - 1u;
                ws.tasks[pos++] = (uint16_t)((lane << 8) | k);
            }
        }
        __syncwarp();
        // ---- rounds of 32 tasks: slab tests of the cluster members; survivors accumulate in the warp's lists ----
#pragma unroll 1
        for (uint32_t t0 = 0u; t0 < total; t0 += 32u) {
            // a round adds at most 256 survivors: run the exact stage first if a list could overflow
            if (ws.count[0] > COOP_SURV_CAP - 256u || ws.count[1] > COOP_SURV_CAP - 256u) coop_exact(S, cc, ws, o, d, all);
            const uint32_t t = t0 + lane;
            const bool has = t < total;
            const uint32_t task = has ? ws.tasks[t] : (lane << 8);
            const uint32_t owner = task >> 8;
            // the owner's ray set-up, fetched by indexed shuffle (every lane takes part)
            RayBvh r2;
            r2.rc.inv = mk3(__shfl_sync(FULL, rb.rc.inv.x, owner), __shfl_sync(FULL, rb.rc.inv.y, owner), __shfl_sync(FULL, rb.rc.inv.z, owner));
            r2.rc.od = mk3(__shfl_sync(FULL, rb.rc.od.x, owner), __shfl_sync(FULL, rb.rc.od.y, owner), __shfl_sync(FULL, rb.rc.od.z, owner));
            r2.rc.ainv = mk3(fabs_(r2.rc.inv.x), fabs_(r2.rc.inv.y), fabs_(r2.rc.inv.z));
            r2.rho = __shfl_sync(FULL, rho, owner);
            const bool owner_all = __shfl_sync(FULL, (int)all, owner) != 0;
            if (has) {
                const uint32_t payload = __float_as_uint(S.top[2 * (ch * 32u + (task & 0xffu)) + 1].z);
                const uint32_t first = payload >> 4, count = payload & 15u;
                const f4 *mb = S.member_box + first;
#pragma unroll
                for (uint32_t j = 0; j < 8u; ++j) {
                    // a cluster of one beyond the first 4 singles was already tested as a top entry
                    if (j < count && (count == 1u || owner_all || member_may_hit(r2, mb[j], finf()))) coop_push(S, ws, owner, first + j);
                }
            }
            __syncwarp();
        }
    }
    coop_exact(S, cc, ws, o, d, all);
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

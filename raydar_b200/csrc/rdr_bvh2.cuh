// rdr_bvh2.cuh -- warp-cooperative traversal of the pair-packed hierarchy (MODE 7) for scenes beyond the fused scan
// (BASELINE config 4: 100k objects, 4 MB of nodes read in place from global memory / L2).
//
// The per-lane traversal (trace_bvh, rdr_trace.cuh) keeps a 64-entry stack per lane in local memory and walks one ray
// per lane: lanes finish at different times (12 of 32 active on config 4) and every node costs 8 scalar box tests.
// Here the rays keep ONE STACK EACH in shared memory, but the WARP serves them together:
//   A0   every lane tests ITS ray against the <= 32 root entries (constant bank, FFMA2 pairs: top_scan of rdr_fused.cuh);
//        root primitives go to the survivor lists; the root's child nodes the ray touches are kept as a bit mask in
//        FRONT-TO-BACK order for the ray's direction octant (the builder's split tree gives 8 fixed permutations);
//   S    each round the first 8 rays that still have work contribute their next node: the top of the ray's stack, else
//        the nearest untouched child of the root;
//   N    FOUR LANES PER TASK: each lane loads one pair of the node's entries with two 256-bit loads (LDG.256, sm_100:
//        the 4 lanes of a task read one whole 128-byte line per instruction) and tests it with FFMA2, bounded by the
//        ray's best exact t so far; the ray's slab constants come by indexed shuffle;
//   P    child nodes go onto the owner's stack far first (every node carries its entries' near-to-far ranks per octant,
//        like the root), primitives to the warp's sphere / cube survivor lists; every position is a popcount of a
//        ballot or of the group's rank mask (no shuffle scan, no atomics);
//   E    full groups of 32 survivors (and, when no ray has work left, the rest) get the exact, reference-ordered test, one
//        per lane, folded into the owner's winner with a 64-bit atomicMin on the (t, original index) key.
// Per-ray depth-first order is what makes the hierarchy pay: a ray descends its nearest subtree first, finds its hit and
// prunes the rest.  Measured with the host simulation of the disciplines (tests/hostsim, config 4, primary rays): one
// shared LIFO stack for the warp 77 node visits and 22 exact tests per ray; per-ray stacks with the root in front-to-
// back order 24 and 6.7 (19 and 4.4 with the children of every node ordered too).  The shared-stack version was built
// first (91 -> 144 Msamples/s after the 4-lanes-per-node change) and replaced.
// A ray's stack holds at most 8 + 7 x 5 entries (node levels 2..8, BVH_MAX_DEPTH).
// The winner is decided by the exact tests and the (t, index) rule only, exactly as in every other search.
// Device-only; must be entered by all 32 lanes of a warp.  Needs lay.bvh2_ok.
#pragma once

#include "rdr_fused.cuh"

namespace rdr {

// two consecutive quads (32 bytes, 32-byte aligned) with one 256-bit load
__device__ __forceinline__ void ld_quads2(const f4 *p, f4 &a, f4 &b)
{
#ifdef RDR_WARP_EMU
    a = p[0]; b = p[1];
#else
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#endif
}

constexpr uint32_t BVH2_RAY_STACK = 56u;                      // per ray: 8 + 7 x 5 pending siblings (node levels 2..8) and slack
constexpr uint32_t BVH2_SURV_CAP = 31u + 256u + 1u;           // carried remainder + one root chunk of 32 x 8 primitives
// stack entries are 16-bit node indices (the packer enables this traversal for hierarchies of <= 65535 nodes, i.e. up to
// ~400k objects): 6.3 KB of shared memory per warp instead of 9.9 KB, which is what lets 24 warps and their cold columns
// share an SM (768-thread CTA; 640 threads, 20 warps and no cold columns before)
constexpr uint32_t BVH2_MAX_NODES = 65535u;
constexpr uint32_t BVH2_WARP_BYTES = 32u * 8u + 2u * 32u * BVH2_RAY_STACK + 4u * 32u + 8u * 8u + 2u * 4u * BVH2_SURV_CAP;

struct Bvh2Warp {
    unsigned long long *best;     // [32] winner key per ray (lane)
    uint16_t *stack;              // [32][BVH2_RAY_STACK] pending child nodes per ray, top = next
    uint32_t *depth;              // [32] entries on each ray's stack
    uint2 *slot;                  // [8]  this round's tasks: (owner lane, node)
    uint32_t *surv_s, *surv_c;    // [BVH2_SURV_CAP] owner lane << 27 | object; surv_c = surv_s + BVH2_SURV_CAP
};

__device__ __forceinline__ Bvh2Warp bvh2_warp(unsigned char *base, uint32_t warp)
{
    unsigned char *p = base + (size_t)warp * BVH2_WARP_BYTES;
    Bvh2Warp w;
    w.best = reinterpret_cast<unsigned long long *>(p);
    w.slot = reinterpret_cast<uint2 *>(p + 256u);
    w.depth = reinterpret_cast<uint32_t *>(p + 256u + 64u);
    w.stack = reinterpret_cast<uint16_t *>(w.depth + 32u);
    w.surv_s = reinterpret_cast<uint32_t *>(w.stack + 32u * BVH2_RAY_STACK);
    w.surv_c = w.surv_s + BVH2_SURV_CAP;
    return w;
}

template <bool SPHERE>
__device__ __forceinline__ void bvh2_exact(const f4 *obj_geom, Bvh2Warp ws, uint32_t lane, const uint32_t *list, uint32_t base,
                                           uint32_t n, v3 o, v3 d)
{
    const uint32_t FULL = 0xffffffffu;
    const bool has = lane < n;
    const uint32_t e = has ? list[base + lane] : (lane << 27);
    const uint32_t own = e >> 27, idx = e & 0x07ffffffu;
    const v3 ro = mk3(__shfl_sync(FULL, o.x, own), __shfl_sync(FULL, o.y, own), __shfl_sync(FULL, o.z, own));
    const v3 rd = mk3(__shfl_sync(FULL, d.x, own), __shfl_sync(FULL, d.y, own), __shfl_sync(FULL, d.z, own));
    if (has) {
        const f4 g = obj_geom[idx];
        float tt;
        const bool hit = SPHERE ? hit_sphere_exact(ro, rd, mk3(g.x, g.y, g.z), g.w, &tt) : hit_cube_exact(ro, rd, mk3(g.x, g.y, g.z), g.w, &tt);
        if (hit) atomicMin(&ws.best[own], coop_key(tt, (int)idx));
    }
    __syncwarp();
}

// Root primitives: the entries this lane's ray hit (bit k of `bits` = primitive pay[k]) go to the warp's survivor lists;
// one packed prefix sum gives every lane its positions.  n_s / n_c: list lengths, warp-uniform.
template <int N>
__device__ __forceinline__ void bvh2_emit_prims(Bvh2Warp ws, uint32_t lane, uint32_t bits, const uint32_t (&pay)[N], uint32_t &n_s, uint32_t &n_c)
{
    uint32_t cb = 0u;
#pragma unroll
    for (int k = 0; k < N; ++k) cb |= ((pay[k] >> 30) & 1u) << k;
    cb &= bits;
    const uint32_t packed = __popc(bits ^ cb) | (__popc(cb) << 16);
    const uint32_t incl = warp_scan_incl(packed, lane);
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t excl = incl - packed;
    uint32_t ps = n_s + (excl & 0xffffu), pc = BVH2_SURV_CAP + n_c + (excl >> 16);     // surv_c = surv_s + BVH2_SURV_CAP
    const uint32_t tag = lane << 27;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        if ((bits >> k) & 1u) {
            const bool cube = (cb >> k) & 1u;
            ws.surv_s[cube ? pc : ps] = tag | (pay[k] & 0x07ffffffu);
            pc += cube ? 1u : 0u; ps += cube ? 0u : 1u;
        }
    }
    n_s += total & 0xffffu; n_c += total >> 16;
    __syncwarp();
}

__device__ __forceinline__ Hit trace_bvh2(const SceneView &S, const FrameParams &P, Bvh2Warp ws, bool alive, v3 o, v3 d)
{
    const uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const SlabRay R = slab_ray_setup(P.cull, o, d);
    ws.best[lane] = ~0ull;
    ws.depth[lane] = 0u;
    uint32_t m = top_scan(P.top, P.lay.bvh2_root, R);
    if (!alive) m = 0u;
    // the root's child nodes this ray touches, as bits in FRONT-TO-BACK order for its direction octant (rdr_bvh.h)
    const uint32_t oct = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
    uint32_t mr = 0u;
    for (uint32_t mm = m & ~P.top.prim_mask; mm != 0u; mm &= mm - 1u) {
        const uint32_t k = (uint32_t)__ffs((int)mm) - 1u;
        mr |= 1u << ((uint32_t)(P.top.rank8[k] >> (5u * oct)) & 31u);
    }
    m &= P.top.prim_mask;                                         // what is left in entry order: the root's primitives
    __syncwarp();

    const f4 *nodes = reinterpret_cast<const f4 *>(P.blob + P.lay.off_nodes2);
    const uint32_t root_chunks = (P.lay.bvh2_root + 7u) >> 3;
    const uint32_t sub = lane & 3u, grp = lane >> 2;              // node stage: 4 lanes per task, one pair of entries per lane
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t n_s = 0u, n_c = 0u, chunk = 0u;                      // survivor list lengths: warp-uniform
#pragma unroll 1
    for (;;) {
        bool last = false;
        if (chunk < root_chunks) {
            // ---- the root's primitives, 8 entries at a time, to the survivor lists ----
            uint32_t pay[8];
#pragma unroll
            for (uint32_t k = 0; k < 8u; ++k) pay[k] = P.top.payload[8u * chunk + k];
            bvh2_emit_prims<8>(ws, lane, (m >> (8u * chunk)) & 0xffu, pay, n_s, n_c);
            ++chunk;
        } else {
            // ---- select: the first 8 rays that still have work contribute their NEXT node each (top of the ray's stack,
            //      else the nearest untouched child of the root) ----
            const uint32_t dep = ws.depth[lane];
            const bool work = dep != 0u || mr != 0u;
            const uint32_t active = __ballot_sync(FULL, work);
            last = active == 0u;
            if (!last) {
                const uint32_t idx = __popc(active & lt);
                if (work && idx < 8u) {
                    uint32_t node;
                    if (dep != 0u) { node = ws.stack[lane * BVH2_RAY_STACK + dep - 1u]; ws.depth[lane] = dep - 1u; }
                    else { const uint32_t r = (uint32_t)__ffs((int)mr) - 1u; mr &= mr - 1u; node = P.top.node_by_rank[oct][r]; }
                    ws.slot[idx] = make_uint2(lane, node);
                }
                const uint32_t n_act = __popc(active), n_tasks = n_act < 8u ? n_act : 8u;
                __syncwarp();
                // ---- N: 4 lanes per task, one pair of the node's entries per lane (the 4 lanes read one whole 128-byte
                //      line per LDG.256), bounded by the owner's best exact t so far ----
                const bool has = grp < n_tasks;
                const uint2 task = has ? ws.slot[grp] : make_uint2(lane, 0u);
                const uint32_t owner = task.x;
                const float qx = __shfl_sync(FULL, R.rx, owner), qy = __shfl_sync(FULL, R.ry, owner), qz = __shfl_sync(FULL, R.rz, owner);
                const float mx = __shfl_sync(FULL, R.nx, owner), my = __shfl_sync(FULL, R.ny, owner), mz = __shfl_sync(FULL, R.nz, owner);
                const f32x2 rho2 = bc2(__shfl_sync(FULL, R.rho, owner));
                // best exact t of the owner so far: high word of its key; ~0 (no hit) and a NaN hit read as NaN = no bound
                const float best = __uint_as_float(reinterpret_cast<const uint32_t *>(ws.best + owner)[1]);
                const uint32_t dbase = ws.depth[owner];           // after the pop above
                const f4 *nd = nodes + 16u * (size_t)task.y + 2u * sub;        // node2_quad(i, sub): quads (0, 1) here, (2, 3) one line on
                f4 q0, q1, q2, q3;
                ld_quads2(nd, q0, q1);
                ld_quads2(nd + 8, q2, q3);
                const uint32_t pa = __float_as_uint(q3.z), pb = __float_as_uint(q3.w);
                // payload bit 29: the box grows by the ray's rho (a sphere, or a node that contains one)
                const f32x2 sp = pk2((pa & 0x20000000u) ? 1.0f : 0.0f, (pb & 0x20000000u) ? 1.0f : 0.0f);
                const f32x2 ex = fma2(sp, rho2, pk2(q1.z, q1.w)), ey = fma2(sp, rho2, pk2(q2.x, q2.y)), ez = fma2(sp, rho2, pk2(q2.z, q2.w));
                uint32_t bits = slab_pair<true>(pk2(q0.x, q0.y), pk2(q0.z, q0.w), pk2(q1.x, q1.y), ex, ey, ez, qx, qy, qz, mx, my, mz, best);
                if (!has) bits = 0u;
                // unused entries (payload ~0: a ray that skips the cull passes every box) are dropped here
                const bool ha = (bits & 1u) && pa != 0xffffffffu, hb = (bits & 2u) && pb != 0xffffffffu;
                const bool na = ha && !(pa >> 31), nb = hb && !(pb >> 31);                       // child nodes
                const bool ca = ha && (pa >> 30) == 3u, cb = hb && (pb >> 30) == 3u;              // cubes
                const bool sa = ha && (pa >> 30) == 2u, sb = hb && (pb >> 30) == 2u;              // spheres
                // ---- P: child nodes onto the owner's stack FAR FIRST (the nearest ends on top): the node carries, for each
                //      direction octant, the rank of every entry in its near-to-far order (rdr_bvh.h); the group ORs its hit
                //      children into a mask in rank space (two butterfly shuffles) and an entry's slot is the number of hit
                //      children behind it.  Primitives go to the warp's survivor lists (positions from four ballots). ----
                const uint32_t octq = (qx < 0.0f ? 3u : 0u) + (qy < 0.0f ? 6u : 0u) + (qz < 0.0f ? 12u : 0u);     // 3 x octant of the owner
                const uint32_t ra = (__float_as_uint(q3.x) >> octq) & 7u, rb = (__float_as_uint(q3.y) >> octq) & 7u;
                uint32_t gm = (na ? 1u << ra : 0u) | (nb ? 1u << rb : 0u);
                gm |= __shfl_xor_sync(FULL, gm, 1);
                gm |= __shfl_xor_sync(FULL, gm, 2);
                const uint32_t sbase = owner * BVH2_RAY_STACK + dbase;
                __syncwarp();                                     // every lane of the group has read depth[owner] before lane 0 rewrites it
                if (na) ws.stack[sbase + __popc(gm >> (ra + 1u))] = (uint16_t)(pa & 0xffffu);
                if (nb) ws.stack[sbase + __popc(gm >> (rb + 1u))] = (uint16_t)(pb & 0xffffu);
                if (has && sub == 0u) ws.depth[owner] = dbase + __popc(gm);
                const uint32_t bsa = __ballot_sync(FULL, sa), bsb = __ballot_sync(FULL, sb);
                const uint32_t bca = __ballot_sync(FULL, ca), bcb = __ballot_sync(FULL, cb);
                const uint32_t tag = owner << 27;
                const uint32_t ps = n_s + __popc(bsa & lt) + __popc(bsb & lt), pc = n_c + __popc(bca & lt) + __popc(bcb & lt);
                if (sa) ws.surv_s[ps] = tag | (pa & 0x07ffffffu);
                if (sb) ws.surv_s[ps + (sa ? 1u : 0u)] = tag | (pb & 0x07ffffffu);
                if (ca) ws.surv_c[pc] = tag | (pa & 0x07ffffffu);
                if (cb) ws.surv_c[pc + (ca ? 1u : 0u)] = tag | (pb & 0x07ffffffu);
                n_s += __popc(bsa) + __popc(bsb); n_c += __popc(bca) + __popc(bcb);
                __syncwarp();
            }
        }
        // ---- E: exact tests on full groups of 32 survivors, and on the rest once no ray has work left ----
#pragma unroll 1
        while (n_s >= 32u || (last && n_s != 0u)) {
            const uint32_t n = n_s < 32u ? n_s : 32u;
            n_s -= n;
            bvh2_exact<true>(S.obj_geom, ws, lane, ws.surv_s, n_s, n, o, d);
        }
#pragma unroll 1
        while (n_c >= 32u || (last && n_c != 0u)) {
            const uint32_t n = n_c < 32u ? n_c : 32u;
            n_c -= n;
            bvh2_exact<false>(S.obj_geom, ws, lane, ws.surv_c, n_c, n, o, d);
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

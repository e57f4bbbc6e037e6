// rdr_bvh2.cuh -- warp-cooperative traversal of the pair-packed hierarchy (MODE 7) for scenes beyond the fused scan
// (BASELINE config 4: 100k objects, 4 MB of nodes read in place from global memory / L2).
//
// The per-lane traversal (trace_bvh, rdr_trace.cuh) keeps a 64-entry stack per lane in local memory and walks one ray
// per lane: lanes finish at different times (12 of 32 active on config 4) and every node costs 8 scalar box tests.
// Here the WARP owns one stack of (ray, node) tasks in shared memory and all 32 lanes work on it together:
//   A0   every lane tests ITS ray against the <= 32 root entries (constant bank, FFMA2 pairs: top_scan of rdr_fused.cuh);
//        the root is then consumed in chunks of 8 entries exactly like a node;
//   N    pop up to 8 tasks from the top of the stack, FOUR LANES PER TASK -- usually some other lane's ray, whose slab
//        constants come by indexed shuffle; each lane loads one pair of the node's entries (quad-major node: the 4 lanes
//        read 64 contiguous bytes per LDG.128) and tests it with FFMA2, bounded by the ray's best exact t so far;
//   P    the hit entries are split by payload bits into child nodes, spheres and cubes; ONE packed shuffle prefix sum
//        (3 x 10 bits) gives every lane its positions on the stack and in the two survivor lists;
//   E    full groups of 32 survivors (and, when the stack is empty, the rest) get the exact, reference-ordered test, one
//        per lane, folded into the owner's winner with a 64-bit atomicMin on the (t, original index) key.
// LIFO order makes the warp go depth-first, so hits arrive early and prune the rest (a task whose node cannot beat the
// ray's best is simply expanded to nothing).  The stack cannot overflow: the root pushes at most 32 x 32 tasks, a round
// pops P <= 8 tasks and pushes at most 8 P, and P is throttled so that the stack stays below its soft limit; above
// it P = 1, i.e. plain depth-first descent of one task, which adds at most 7 entries per level (8 levels).
// (One node per LANE -- 32 tasks per round, 17 scattered LDG.128 per lane -- was measured first: 77 % L1 data-pipe
// utilisation, 24 % issue utilisation, profiles/ncu_r01r_config4_bvh2_summary.txt.)
// The winner is decided by the exact tests and the (t, index) rule only, exactly as in every other search.
// Device-only; must be entered by all 32 lanes of a warp.  Needs lay.bvh2_ok.
#pragma once

#include "rdr_fused.cuh"

namespace rdr {

constexpr uint32_t BVH2_STACK_SOFT = 1024u;
constexpr uint32_t BVH2_STACK_CAP = BVH2_STACK_SOFT + 7u * 8u + 8u;
constexpr uint32_t BVH2_SURV_CAP = 31u + 256u + 1u;
constexpr uint32_t BVH2_WARP_BYTES = 32u * 8u + 4u * BVH2_STACK_CAP + 2u * 4u * BVH2_SURV_CAP;

struct Bvh2Warp {
    unsigned long long *best;     // [32] winner key per lane
    uint32_t *stack;              // [BVH2_STACK_CAP]  owner lane << 27 | node
    uint32_t *surv_s, *surv_c;    // [BVH2_SURV_CAP]   owner lane << 27 | object
};

__device__ __forceinline__ Bvh2Warp bvh2_warp(unsigned char *base, uint32_t warp)
{
    unsigned char *p = base + (size_t)warp * BVH2_WARP_BYTES;
    Bvh2Warp w;
    w.best = reinterpret_cast<unsigned long long *>(p);
    w.stack = reinterpret_cast<uint32_t *>(p + 256u);
    w.surv_s = w.stack + BVH2_STACK_CAP;
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

// P: the entries this lane found hit (bit k of `bits` = entry pay[k]) go to the warp's stack (child nodes) or survivor
// lists (spheres, cubes), by payload bits; ONE packed prefix sum (3 x 10 bits) gives every lane its positions.
// Unrolled and predicated: no dependent loads, no divergent loop.  n_t / n_s / n_c: list lengths, warp-uniform.
template <int N>
__device__ __forceinline__ void bvh2_emit(Bvh2Warp ws, uint32_t lane, uint32_t owner, uint32_t bits, const uint32_t (&pay)[N],
                                          uint32_t &n_t, uint32_t &n_s, uint32_t &n_c)
{
    uint32_t nb = 0u, cb = 0u;                                    // child-node bits, cube bits
#pragma unroll
    for (int k = 0; k < N; ++k) {
        nb |= ((pay[k] >> 31) ^ 1u) << k;
        cb |= (((pay[k] >> 30) & 1u) & (pay[k] >> 31)) << k;
    }
    nb &= bits; cb &= bits;
    const uint32_t packed = __popc(nb) | (__popc(bits ^ nb ^ cb) << 10) | (__popc(cb) << 20);
    const uint32_t incl = warp_scan_incl(packed, lane);
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t excl = incl - packed;
    // the three lists are one array (stack | surv_s | surv_c): a position is all that is selected per entry
    uint32_t pn = n_t + (excl & 1023u), ps = BVH2_STACK_CAP + n_s + ((excl >> 10) & 1023u), pc = BVH2_STACK_CAP + BVH2_SURV_CAP + n_c + (excl >> 20);
    const uint32_t tag = owner << 27;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        if ((bits >> k) & 1u) {
            const bool node = (nb >> k) & 1u, cube = (cb >> k) & 1u;
            ws.stack[node ? pn : (cube ? pc : ps)] = tag | (pay[k] & 0x07ffffffu);
            pn += node ? 1u : 0u; pc += (!node && cube) ? 1u : 0u; ps += (!node && !cube) ? 1u : 0u;
        }
    }
    n_t += total & 1023u; n_s += (total >> 10) & 1023u; n_c += total >> 20;
    __syncwarp();
}

__device__ __forceinline__ Hit trace_bvh2(const SceneView &S, const FrameParams &P, Bvh2Warp ws, bool alive, v3 o, v3 d)
{
    const uint32_t FULL = 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    const SlabRay R = slab_ray_setup(P.cull, o, d);
    ws.best[lane] = ~0ull;
    uint32_t m = top_scan(P.top, P.lay.bvh2_root, R);
    if (!alive) m = 0u;
    __syncwarp();

    const f4 *nodes = reinterpret_cast<const f4 *>(P.blob + P.lay.off_nodes2);
    const uint32_t root_chunks = (P.lay.bvh2_root + 7u) >> 3;
    const uint32_t sub = lane & 3u, grp = lane >> 2;              // node stage: 4 lanes per task, one pair of entries per lane
    uint32_t n_t = 0u, n_s = 0u, n_c = 0u, chunk = 0u;            // stack / survivor list lengths: warp-uniform
#pragma unroll 1
    for (;;) {
        const bool rooting = chunk < root_chunks;
        const bool last = !rooting && n_t == 0u;
        if (rooting) {
            // ---- the root, 8 entries at a time: the lane's own ray, payloads from the kernel parameters ----
            uint32_t pay[8];
#pragma unroll
            for (uint32_t k = 0; k < 8u; ++k) pay[k] = P.top.payload[8u * chunk + k];
            bvh2_emit<8>(ws, lane, lane, (m >> (8u * chunk)) & 0xffu, pay, n_t, n_s, n_c);
            ++chunk;
        } else if (!last) {
            // ---- N: pop up to 8 tasks (throttled near the soft limit); the 4 lanes of a group test one pair each.  The
            //      node is quad-major, so a group reads 64 contiguous bytes per load: 8 wavefronts per LDG.128 instead of the
            //      32 of one-node-per-lane (the traversal is bound by the L1 data pipe) ----
            uint32_t pop = (BVH2_STACK_SOFT > n_t ? BVH2_STACK_SOFT - n_t : 0u) >> 3;      // <= room / 7
            pop = pop < 1u ? 1u : (pop > 8u ? 8u : pop);
            if (pop > n_t) pop = n_t;
            const bool has = grp < pop;
            const uint32_t task = has ? ws.stack[n_t - 1u - grp] : (lane << 27);
            n_t -= pop;
            __syncwarp();                                         // the popped slots are overwritten by the pushes below
            const uint32_t owner = task >> 27;
            const float qx = __shfl_sync(FULL, R.rx, owner), qy = __shfl_sync(FULL, R.ry, owner), qz = __shfl_sync(FULL, R.rz, owner);
            const float mx = __shfl_sync(FULL, R.nx, owner), my = __shfl_sync(FULL, R.ny, owner), mz = __shfl_sync(FULL, R.nz, owner);
            const f32x2 rho2 = bc2(__shfl_sync(FULL, R.rho, owner));
            // best exact t of the owner so far: high word of its key; ~0 (no hit) and a NaN hit read as NaN = no bound
            const float best = __uint_as_float(reinterpret_cast<const uint32_t *>(ws.best + owner)[1]);
            const f4 *nd = nodes + 16u * (size_t)(task & 0x07ffffffu) + sub;
            const f4 q0 = nd[0], q1 = nd[4], q2 = nd[8], q3 = nd[12];
            const f32x2 sp = pk2(q3.x, q3.y);
            const f32x2 ex = fma2(sp, rho2, pk2(q1.z, q1.w)), ey = fma2(sp, rho2, pk2(q2.x, q2.y)), ez = fma2(sp, rho2, pk2(q2.z, q2.w));
            uint32_t bits = slab_pair<true>(pk2(q0.x, q0.y), pk2(q0.z, q0.w), pk2(q1.x, q1.y), ex, ey, ez, qx, qy, qz, mx, my, mz, best);
            uint32_t pay[2] = {__float_as_uint(q3.z), __float_as_uint(q3.w)};
            if (pay[0] == 0xffffffffu) bits &= ~1u;              // unused entries (a ray that skips the cull passes every box)
            if (pay[1] == 0xffffffffu) bits &= ~2u;
            if (!has) bits = 0u;
            bvh2_emit<2>(ws, lane, owner, bits, pay, n_t, n_s, n_c);
        }
        // ---- E: exact tests on full groups of 32 survivors, and on the rest once the stack is empty ----
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

// rdr_kernels.cu -- sm_100a kernels of the path-tracing sample loop.
//
//   render_kernel      render_next_sample x N (cpu.rs:193-219) fused with per_pixel (cpu.rs:233-342):
//                      one lane owns one pixel, loops over its samples and bounces, accumulates in
//                      registers and issues one 16-byte load and one 16-byte store of the accumulator.
//   resolve_kernel     print_frame_buffer (cpu.rs:221-230)
//   primary_kernel     the frame's primary table: camera ray + nearest hit per pixel, once per frame (also rdr_first_hit)
//   trace_path_kernel  debug/parity: one path with every bounce recorded
//   kat_* kernels      per-function known-answer entry points
#include <cuda_runtime.h>

#include "raydar_cuda.h"
#include "rdr_device.cuh"
#include "rdr_fused.cuh"
#include "rdr_bvh2.cuh"
#include "rdr_loop.cuh"
#include "rdr_launch.h"

namespace rdr {

// Dynamic shared memory: [scene blob (when staged)][mbarrier, 16 B][per-lane scratch words].
// Scratch = candidate masks of the scan (one word per 32-primitive chunk) or the BVH candidate queue
// (BVH_QCAP words), strided by the block size so that lanes hit distinct banks.
// A scene whose blob does not fit (FrameParams::staged == 0: large BVH scenes) is read in place from
// global memory / L2 and only the scratch words live in shared memory.
// MODE 5 (fused scan) always runs on a staged blob: returning `smem` itself (not a select between two pointers)
// lets the compiler prove the address space, so every scene access becomes an LDS with a 32-bit address.
// The fused kernels (MODE 5 / 6) need only the blob's prefix (obj_geom, material and the fused scan's sections:
// SceneLayout::fused_stage_bytes) unless they also run the per-lane twin of the scan (FULL: trace_path_kernel).
template <int MODE, bool FULL>
__device__ __forceinline__ uint32_t staged_bytes(const FrameParams &P)
{
    return ((MODE == 5 || MODE == 6) && !FULL) ? P.lay.fused_stage_bytes : P.lay.blob_bytes;
}
template <int MODE, bool FULL = false>
__device__ __forceinline__ const unsigned char *stage_scene(unsigned char *smem, const FrameParams &P)
{
    if (MODE == 7 || (MODE < 5 && !P.staged)) return P.blob;       // MODE 7 reads the scene in place (global memory / L2)
    const uint32_t bytes = staged_bytes<MODE, FULL>(P);
    stage_blob(smem, P.blob, bytes, reinterpret_cast<uint64_t *>(smem + bytes));
    return smem;
}
template <int MODE, bool FULL = false>
__device__ __forceinline__ uint32_t *scratch_base(unsigned char *smem, const FrameParams &P)
{
    return (MODE != 7 && (MODE >= 5 || P.staged)) ? reinterpret_cast<uint32_t *>(smem + staged_bytes<MODE, FULL>(P) + 16u)
                                                   : reinterpret_cast<uint32_t *>(smem);
}

// nearest hit for every lane of the warp (alive = the lane has a ray).  MODE 4 regroups the work across the warp
// and must be entered by all 32 lanes; the other searches are per-lane.
// scratch layout: MODE 4: [one word per lane][per-warp cooperative regions], otherwise [words per lane, strided].
template <int MODE>
__device__ __forceinline__ Hit trace_warp(const SceneView &S, const FrameParams &P, uint32_t *scratch0, bool alive, v3 o, v3 d)
{
    if (MODE == 7) return trace_bvh2(S, P, bvh2_warp(reinterpret_cast<unsigned char *>(scratch0 + blockDim.x), threadIdx.x >> 5), alive, o, d);
    if (MODE == 5 || MODE == 6) {
        FusedView V;
        V.pair_block = S.pair_block; V.member_geom = S.fused_geom; V.member_idx = S.fused_idx;
        return trace_fused<MODE == 5>(V, P, fused_warp(reinterpret_cast<unsigned char *>(scratch0 + blockDim.x), threadIdx.x >> 5), alive, o, d);
    }
    if (MODE == 4) {
        unsigned char *coop = reinterpret_cast<unsigned char *>(scratch0 + blockDim.x);
        return trace_cluster_coop(S, P.cull, coop_scratch(coop, threadIdx.x >> 5), alive, o, d);
    }
    Hit h; h.idx = -1; h.t = finf();
    if (alive) h = trace_any<MODE>(S, P.cull, scratch0 + threadIdx.x, blockDim.x, o, d);
    return h;
}

// ---- the sample loop (LaneState / lane_shade / trace_brute in rdr_trace.cuh) ------------------------------
// Persistent lanes in warp lock-step.  The grid is sized to the machine (SMs x resident CTAs), not to the
// image.  Each iteration:
//   1. a lane without work claims the next unclaimed pixel (one atomicAdd on a global counter; pixels are
//      handed out in row-major order, so a warp works on neighbouring pixels) and sets up its primary ray;
//   2. the warp votes (__any_sync doubles as the reconvergence point -- without it Volta+ independent thread
//      scheduling lets the lanes drift apart until each executes the scan alone);
//   3. every lane with a ray runs the scan together, then shades; a lane whose pixel is finished writes its
//      accumulator (one 16-byte store per pixel per launch, after one 16-byte load when it claimed it).
// A pixel is always processed by exactly one lane with its samples in ascending order, so results do not
// depend on the schedule (bit-identical to the per-pixel host loop).
template <int MODE, int BLOCK, int MIN_CTAS, bool COLD>
__global__ void __launch_bounds__(BLOCK, MIN_CTAS) render_kernel(const __grid_constant__ FrameParams P)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const SceneView S = scene_view(stage_scene<MODE>(smem, P), P.lay);
    uint32_t *scratch0 = scratch_base<MODE>(smem, P);
#define RDR_LOOP_STATE typename lane_state_of<BLOCK, COLD>::type
    // the cold columns follow the per-warp scratch of the cooperative search (mode_smem_bytes)
#define RDR_LOOP_COLD_BASE (reinterpret_cast<unsigned char *>(scratch0 + BLOCK) + (BLOCK / 32) * (MODE == 7 ? BVH2_WARP_BYTES : FUSED_WARP_BYTES))
#define RDR_LOOP_LANE_SCRATCH (scratch0 + threadIdx.x)
#define RDR_LOOP_TRACE(alive, o, d) trace_warp<MODE>(S, P, scratch0, alive, o, d)
#include "rdr_loop_body.inc"
#undef RDR_LOOP_STATE
#undef RDR_LOOP_COLD_BASE
#undef RDR_LOOP_LANE_SCRATCH
#undef RDR_LOOP_TRACE
}

// ---- print_frame_buffer (cpu.rs:221-230): one uchar4 (32-bit) store per pixel ------------------------
__global__ void __launch_bounds__(256) resolve_kernel(const f4 *__restrict__ accum, uchar4 *__restrict__ out,
                                                      uint32_t n_pixels, float divisor)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    const f4 a = accum[i];
    uchar4 o;
    o.x = (unsigned char)quantise(a.x, divisor); o.y = (unsigned char)quantise(a.y, divisor);
    o.z = (unsigned char)quantise(a.z, divisor); o.w = (unsigned char)quantise(a.w, divisor);
    out[i] = o;
}

// ---- multi-GPU combine: accumulator sum over NVLink peer memory fused with print_frame_buffer -----------------
// Each GPU owns 1/G of the pixels.  A thread reads its pixel from EVERY accumulator -- its own and the peers', mapped
// through NVLink peer access (one process) or CUDA IPC (one process per GPU) -- with G independent 16-byte loads in
// flight, adds them in source order (device 0 first: a fixed order, so the result is deterministic, unlike a ring or
// tree reduce), quantises (cpu.rs:221-230) and stores the pixel's RGBA8 word where the image is wanted: pinned host
// memory (every GPU writes its slice over its own PCIe link) or a peer's device image.  No rooted reduce: nothing lands
// on one GPU, 4x fewer bytes leave the GPUs than with an f32 reduce.  sum != NULL writes the f32 sum instead (parity).
// Row-stripe partition (stripe_count > 1): the owned pixels are the GPU's own stripes and n_src == 1 -- no peer traffic.
template <int N>
__global__ void __launch_bounds__(256) peer_combine_kernel(const __grid_constant__ PeerCombine C)
{
    const uint32_t step = gridDim.x * blockDim.x;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < C.count; k += step) {
        const uint32_t pixel = C.stripe_count > 1u ? stripe_pixel(C.width, C.stripe_rows, C.stripe_index, C.stripe_count, k) : C.first + k;
        float4 v[N];
#pragma unroll
        for (int g = 0; g < N; ++g) v[g] = __ldcg(reinterpret_cast<const float4 *>(C.src[g] + pixel));
        float4 a = v[0];
#pragma unroll
        for (int g = 1; g < N; ++g) { a.x = fadd(a.x, v[g].x); a.y = fadd(a.y, v[g].y); a.z = fadd(a.z, v[g].z); a.w = fadd(a.w, v[g].w); }
        for (uint32_t g = N; g < C.n_src; ++g) {            // more than N sources (N == 8 only): the rest one by one
            const float4 b = __ldcg(reinterpret_cast<const float4 *>(C.src[g] + pixel));
            a.x = fadd(a.x, b.x); a.y = fadd(a.y, b.y); a.z = fadd(a.z, b.z); a.w = fadd(a.w, b.w);
        }
        if (C.sum) { f4 o; o.x = a.x; o.y = a.y; o.z = a.z; o.w = a.w; C.sum[pixel] = o; }
        if (C.rgba) {
            uchar4 o;
            o.x = (unsigned char)quantise(a.x, C.divisor); o.y = (unsigned char)quantise(a.y, C.divisor);
            o.z = (unsigned char)quantise(a.z, C.divisor); o.w = (unsigned char)quantise(a.w, C.divisor);
            C.rgba[pixel] = o;
        }
    }
}

// ---- debug / parity kernels ---------------------------------------------------------------------------
// ---- the frame's primary table: camera ray (cpu.rs:199-202,234-251) and its nearest hit, once per pixel per frame ----
// The reference's camera ray has no jitter, so every sample of a pixel starts from the same ray and the same first hit.
// rdr_new_frame runs this kernel once; render_kernel then never sets up a camera ray or traces a primary ray, in any
// launch of the frame (a progressive frame of N one-sample launches saves N - 1 primary traces per pixel).  The table
// is also what rdr_first_hit returns.
template <int MODE>
__global__ void __launch_bounds__(RDR_BLOCK, 2) primary_kernel(const __grid_constant__ FrameParams P,
                                                              f4 *__restrict__ primary, int32_t *__restrict__ primary_idx)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const SceneView S = scene_view(stage_scene<MODE>(smem, P), P.lay);
    uint32_t *scratch0 = scratch_base<MODE>(smem, P);
    const uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = pixel < P.cam.width * P.cam.height;
    const v3 o = mk3(P.cam.pos[0], P.cam.pos[1], P.cam.pos[2]);
    const v3 d = valid ? camera_ray_dir(P.cam, pixel % P.cam.width, pixel / P.cam.width) : mk3(0.0f, 0.0f, 1.0f);
    const Hit h = trace_warp<MODE>(S, P, scratch0, valid, o, d);
    if (!valid) return;
    f4 out; out.x = d.x; out.y = d.y; out.z = d.z; out.w = h.idx >= 0 ? h.t : 0.0f;
    primary[pixel] = out;
    primary_idx[pixel] = h.idx;
}

template <int MODE>
__global__ void __launch_bounds__(RDR_BLOCK, 2) kat_trace_kernel(const __grid_constant__ FrameParams P, uint32_t n,
                                                                const float *__restrict__ rays,
                                                                int32_t *__restrict__ ids, float *__restrict__ ts)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const SceneView S = scene_view(stage_scene<MODE>(smem, P), P.lay);
    uint32_t *scratch0 = scratch_base<MODE>(smem, P);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const v3 o = valid ? mk3(rays[6 * i + 0], rays[6 * i + 1], rays[6 * i + 2]) : mk3(0.0f, 0.0f, 0.0f);
    const v3 d = valid ? mk3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]) : mk3(0.0f, 0.0f, 1.0f);
    const Hit h = trace_warp<MODE>(S, P, scratch0, valid, o, d);
    if (!valid) return;
    ids[i] = h.idx;
    ts[i] = h.idx >= 0 ? h.t : 0.0f;
}

// one path, one thread (block of 32 so the staging code is shared; lane 0 walks the path)
template <int MODE>
__global__ void __launch_bounds__(RDR_BLOCK, 2) trace_path_kernel(const __grid_constant__ FrameParams P, uint32_t x, uint32_t y,
                                                                 uint32_t sample, RdrPathStep *__restrict__ steps,
                                                                 uint32_t capacity, uint32_t *__restrict__ n_steps,
                                                                 float *__restrict__ rgba)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const SceneView S = scene_view(stage_scene<MODE, true>(smem, P), P.lay);      // the per-lane twins read the whole blob
    uint32_t *masks = scratch_base<MODE, true>(smem, P) + threadIdx.x;    // (MODE 4 - 7 walk the path with their per-lane twin)
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    *n_steps = trace_path_lane<MODE>(P, S, masks, blockDim.x, x, y, sample, steps, capacity, rgba);
}

__global__ void kat_hit_sphere_kernel(uint32_t n, const float *__restrict__ rays, const float *__restrict__ prims,
                                      float *__restrict__ t_out, int32_t *__restrict__ hit_out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float t = 0.0f;
    const bool h = hit_sphere_exact(mk3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]),
                                    mk3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]),
                                    mk3(prims[4 * i], prims[4 * i + 1], prims[4 * i + 2]), prims[4 * i + 3], &t);
    hit_out[i] = h ? 1 : 0;
    t_out[i] = h ? t : 0.0f;
}

__global__ void kat_hit_cube_kernel(uint32_t n, const float *__restrict__ rays, const float *__restrict__ prims,
                                    float *__restrict__ t_out, int32_t *__restrict__ hit_out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float t = 0.0f;
    const bool h = hit_cube_exact(mk3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]),
                                  mk3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]),
                                  mk3(prims[4 * i], prims[4 * i + 1], prims[4 * i + 2]), prims[4 * i + 3], &t);
    hit_out[i] = h ? 1 : 0;
    t_out[i] = h ? t : 0.0f;
}

// per-function KATs of the shading helpers: n records of 12 floats in, 8 floats out (include/raydar_cuda.h: RDR_KAT_*)
__global__ void kat_vec_kernel(int op, uint32_t n, const float *__restrict__ in, float *__restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *a = in + 12u * i;
    float *o = out + 8u * i;
    for (int k = 0; k < 8; ++k) o[k] = 0.0f;
    const v3 v = mk3(a[0], a[1], a[2]), w = mk3(a[3], a[4], a[5]);
    if (op == RDR_KAT_REFLECT) { const v3 r = reflect3(v, w); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
    else if (op == RDR_KAT_REFRACT) { const v3 r = refract3(v, w, a[6]); o[0] = r.x; o[1] = r.y; o[2] = r.z; }
    else if (op == RDR_KAT_CAN_REFRACT) { o[0] = can_refract3(v, w, a[6]) ? 1.0f : 0.0f; }
    else if (op == RDR_KAT_WORLD_SAMPLE) {
        World wd; wd.kind = a[3] != 0.0f ? 1u : 0u;
        for (int k = 0; k < 3; ++k) { wd.a[k] = a[4 + k]; wd.b[k] = a[7 + k]; }
        const v3 r = world_sample(wd, v); o[0] = r.x; o[1] = r.y; o[2] = r.z;
    }
    else if (op == RDR_KAT_CLOSEST_HIT) {
        const Surface sf = closest_hit(v, w, a[6], a[7] != 0.0f, mk3(a[8], a[9], a[10]), a[11]);
        o[0] = sf.p.x; o[1] = sf.p.y; o[2] = sf.p.z; o[3] = sf.n.x; o[4] = sf.n.y; o[5] = sf.n.z; o[6] = sf.front ? 1.0f : 0.0f;
    }
    else if (op == RDR_KAT_QUANTISE) { o[0] = (float)quantise(a[0], a[1]); }
    else if (op == RDR_KAT_RAND_FLOATS) {            // the rand 0.8.5 float maps on raw 32-bit words (passed as bit patterns)
        u4 wds; wds.x = f2u(a[0]); wds.y = f2u(a[1]); wds.z = f2u(a[2]); wds.w = 0u;
        o[0] = u01(wds.x); o[1] = range_pm1(wds.x);
        const v3 r = random_in_unit_sphere(wds); o[2] = r.x; o[3] = r.y; o[4] = r.z;
    }
}

__global__ void kat_camera_rays_kernel(const __grid_constant__ FrameParams P, uint32_t n, const uint32_t *__restrict__ xy,
                                       float *__restrict__ rays)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const v3 d = camera_ray_dir(P.cam, xy[2 * i], xy[2 * i + 1]);
    rays[6 * i + 0] = P.cam.pos[0]; rays[6 * i + 1] = P.cam.pos[1]; rays[6 * i + 2] = P.cam.pos[2];
    rays[6 * i + 3] = d.x; rays[6 * i + 4] = d.y; rays[6 * i + 5] = d.z;
}

__global__ void kat_rng_kernel(uint32_t seed_lo, uint32_t seed_hi, uint32_t pixel, uint32_t sample, uint32_t bounce,
                               uint32_t block, uint32_t *__restrict__ out)
{
    const u4 r = rng_block(seed_lo, seed_hi, pixel, sample, bounce, block);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// ---- launch wrappers (called from rdr_api.cpp) ---------------------------------------------------------
static inline uint32_t scratch_words(const SceneLayout &L)
{
    if (L.mode == 1u) return (uint32_t)BVH_QCAP;
    const uint32_t chunks = (L.ns_pad > L.nc_pad ? L.ns_pad : L.nc_pad) / 32u;
    return chunks > CL_SCRATCH ? chunks : CL_SCRATCH;       // flat-scan masks or cluster masks + queue
}

// dynamic shared memory of one CTA running kernel variant `mode` (the MODE template argument).
//   cold: the kernel keeps the cold / parked lane state in shared-memory columns (render_kernel<.., COLD = true>)
//   full: a fused kernel that stages the whole blob, not only its prefix (trace_path_kernel)
static size_t mode_smem_bytes(const SceneLayout &L, bool staged, uint32_t block, int mode, bool cold = false, bool full = false)
{
    const size_t warps = (block + 31u) / 32u;
    const size_t cold_bytes = cold ? (size_t)block * (COLD_WORDS + COLD_PARK_WORDS) * sizeof(float) : 0u;
    size_t scratch;
    if (mode == 7) return (size_t)block * sizeof(uint32_t) + warps * BVH2_WARP_BYTES + cold_bytes;     // never staged
    if (mode >= 5) scratch = (size_t)block * sizeof(uint32_t) + warps * FUSED_WARP_BYTES + cold_bytes; // one word per lane + per-warp regions
    else if (mode == 4) scratch = (size_t)block * sizeof(uint32_t) + warps * COOP_WARP_BYTES;
    else scratch = (size_t)scratch_words(L) * block * sizeof(uint32_t);                                // per-lane words
    const size_t blob = (mode >= 5 && !full) ? L.fused_stage_bytes : L.blob_bytes;
    return (staged ? blob + 16u : 0u) + scratch;
}

// the largest footprint any variant of this layout needs (feasibility check in rdr_api.cpp)
size_t scene_smem_bytes(const SceneLayout &L, bool staged, uint32_t block)
{
    if (L.mode == 1u) return mode_smem_bytes(L, staged, block, 2);
    size_t m = mode_smem_bytes(L, staged, block, 0);
    for (int mode = 3; mode <= 4; ++mode) { const size_t b = mode_smem_bytes(L, staged, block, mode); if (b > m) m = b; }
    return m;
}

// kernel variant (the MODE template argument): 0 = flat scan + cull, 1 = flat scan exact-everything (debug),
// 2 = BVH, 3 = two-level cluster scan (per lane), 4 = the same, warp-cooperative.
// `variant` (from rdr_api.cpp) numbers the scans of a scan-packed blob: 0, 1, 2 = cluster, 3 = cooperative cluster,
// 4 = fused scan (needs lay.fused_ok, else the cooperative scan runs).  A BVH-packed blob can only be traversed as a BVH.
static inline int mode_of(const FrameParams &P, int variant)
{
    if (P.lay.mode == 1u) return (variant == 5 && P.lay.bvh2_ok) ? 7 : 2;      // hierarchy: warp-cooperative or per-lane traversal
    if (variant > 4) variant = 4;                                   // scan lists hold no hierarchy: 5 / 6 cannot run on them
    if (variant == 4 && !(P.lay.fused_ok && P.staged)) variant = 3;
    if (variant == 4) return P.lay.fused_cap == 8u ? 5 : 6;           // fused scan: 8-member clusters resolved at compile time
    return variant >= 2 ? variant + 1 : variant;
}

template <typename K>
static cudaError_t set_smem(K kernel, size_t bytes)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    // the kernels keep their working set in shared memory: take the largest carve-out so that the resident CTA count is
    // set by the footprint, not by the default L1/shared split
    return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

#define RDR_DISPATCH(mode, KERNEL, ...)                                   \
    do {                                                                  \
        if ((mode) == 7) { KERNEL(7, __VA_ARGS__); }                      \
        else if ((mode) == 6) { KERNEL(6, __VA_ARGS__); }                 \
        else if ((mode) == 5) { KERNEL(5, __VA_ARGS__); }                 \
        else if ((mode) == 4) { KERNEL(4, __VA_ARGS__); }                 \
        else if ((mode) == 3) { KERNEL(3, __VA_ARGS__); }                 \
        else if ((mode) == 2) { KERNEL(2, __VA_ARGS__); }                 \
        else if ((mode) == 1) { KERNEL(1, __VA_ARGS__); }                 \
        else { KERNEL(0, __VA_ARGS__); }                                  \
    } while (0)

// ---- CTA shape of the render kernel ----------------------------------------------------------------------------
// The scans that keep per-lane state only (MODE 0-4) run 3 CTAs of 256 threads per SM (80 registers).  The warp-
// cooperative kernels are latency-bound on their shuffle / shared-memory chains and want as many resident warps as the
// register file allows: ONE large CTA per SM shares a single staged copy of the scene, and the COLD variants keep the
// lane state the search does not touch in shared memory.  Measured on benchmark.rscn (Msamples/s, profiles/variants_r02*):
//   768 threads @ 80 registers, state in registers   5370     (42 B of spills)
//   768 @ 80, cold + parked state in shared memory   5680
//   896 @ 72, cold + parked                          5780     <- default when the footprint fits
//  1024 @ 64, cold                                   5410     (138 B of spills)
// The shape is chosen per frame: the largest that fits the device's opt-in shared memory.
struct RenderShape { uint32_t block; bool cold; };

static size_t smem_optin_limit()
{
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 48u << 10;
    return (size_t)v;
}
// -DRDR_FUSED_CTA / -DRDR_BVH2_CTA (kernel experiments, raydar_b200.build.build_variant) force a shape: fused 1 = 896 cold,
// 3 = 768 cold, 4 = 768 registers, 0 = 3 x 256 registers; hierarchy 0 = 768, 1 = 640, 2 = 512 (registers).
#ifndef RDR_FUSED_CTA
#define RDR_FUSED_CTA -1
#endif
#ifndef RDR_BVH2_CTA
#define RDR_BVH2_CTA -1
#endif
static RenderShape render_shape(const FrameParams &P, int mode)
{
    if (mode < 5) return RenderShape{RDR_BLOCK, false};
    const size_t limit = smem_optin_limit();
    if (mode == 7) {
        const int cfg = RDR_BVH2_CTA;
        switch (cfg) {
        case 0: return RenderShape{768u, false};
        case 1: return RenderShape{640u, false};
        case 2: return RenderShape{512u, false};
        default: break;
        }
        // 24 warps with the cold / parked state in shared-memory columns when the footprint fits (6.3 KB of stacks and
        // lists per warp + 72 B of columns per lane = 207 KB), else 20 warps with the state in registers
        if (mode_smem_bytes(P.lay, false, 768u, mode, true) <= limit) return RenderShape{768u, true};
        return RenderShape{640u, false};
    }
    const int cfg = RDR_FUSED_CTA;
    if (cfg == 0) return RenderShape{RDR_BLOCK, false};
    if (cfg == 1) return RenderShape{896u, true};
#ifdef RDR_EXTRA_SHAPES
    if (cfg == 5) return RenderShape{832u, true};
    if (cfg == 6) return RenderShape{960u, true};
    if (cfg == 7) return RenderShape{1024u, true};
#endif
    if (cfg == 3) return RenderShape{768u, true};
    if (cfg == 4) return RenderShape{768u, false};
    if (mode_smem_bytes(P.lay, true, 896u, mode, true) <= limit) return RenderShape{896u, true};
    if (mode_smem_bytes(P.lay, true, 768u, mode, true) <= limit) return RenderShape{768u, true};
    return RenderShape{768u, false};
}

#ifdef RDR_EXTRA_SHAPES      // kernel experiments only (build_variant): 832 / 960-thread CTAs of the fused kernel
#define RDR_EXTRA_SHAPE_DISPATCH(shape, F) if ((shape).block == 832u) F((render_kernel<5, 832, 1, true>)); else if ((shape).block == 960u) F((render_kernel<5, 960, 1, true>)); else if ((shape).block == 1024u) F((render_kernel<5, 1024, 1, true>)); else
#else
#define RDR_EXTRA_SHAPE_DISPATCH(shape, F)
#endif
// calls F(kernel) with the render kernel instantiation for `mode` and `shape`
#define RDR_RENDER_DISPATCH(mode, shape, F)                                                           \
    do {                                                                                              \
        if ((mode) == 7) {                                                                            \
            if ((shape).cold) F((render_kernel<7, 768, 1, true>));                                    \
            else if ((shape).block == 768u) F((render_kernel<7, 768, 1, false>));                     \
            else if ((shape).block == 512u) F((render_kernel<7, 512, 1, false>));                     \
            else F((render_kernel<7, 640, 1, false>));                                                \
        }                                                                                             \
        else if ((mode) == 6) {                                                                       \
            if ((shape).cold) { if ((shape).block == 896u) F((render_kernel<6, 896, 1, true>)); else F((render_kernel<6, 768, 1, true>)); } \
            else if ((shape).block == 768u) F((render_kernel<6, 768, 1, false>));                     \
            else F((render_kernel<6, RDR_BLOCK, 3, false>));                                          \
        }                                                                                             \
        else if ((mode) == 5) {                                                                       \
            if ((shape).cold) { RDR_EXTRA_SHAPE_DISPATCH(shape, F) if ((shape).block == 896u) F((render_kernel<5, 896, 1, true>)); else F((render_kernel<5, 768, 1, true>)); } \
            else if ((shape).block == 768u) F((render_kernel<5, 768, 1, false>));                     \
            else F((render_kernel<5, RDR_BLOCK, 3, false>));                                          \
        }                                                                                             \
        else if ((mode) == 4) { F((render_kernel<4, RDR_BLOCK, 3, false>)); }                         \
        else if ((mode) == 3) { F((render_kernel<3, RDR_BLOCK, 3, false>)); }                         \
        else if ((mode) == 2) { F((render_kernel<2, RDR_BLOCK, 3, false>)); }                         \
        else if ((mode) == 1) { F((render_kernel<1, RDR_BLOCK, 3, false>)); }                         \
        else { F((render_kernel<0, RDR_BLOCK, 3, false>)); }                                          \
    } while (0)

// the smallest footprint of the fused render kernel (rdr_api.cpp: does the fused scan fit at all?)
size_t fused_smem_bytes(const SceneLayout &L) { return mode_smem_bytes(L, true, 768u, 5, false); }

cudaError_t launch_render(const FrameParams &P, int variant, int resident_ctas, cudaStream_t stream)
{
    const uint32_t n_pixels = P.owned_pixels;       // pixels this launch hands out (all of them, or the shard's stripes)
    if (n_pixels == 0u) return cudaSuccess;
    const int mode = mode_of(P, variant);
    const RenderShape shape = render_shape(P, mode);
    const uint32_t block = shape.block;
    const size_t smem = mode_smem_bytes(P.lay, P.staged != 0u, block, mode, shape.cold);
    // persistent grid: every resident CTA slot of the device, but no more CTAs than there are pixels to hand out
    uint32_t grid = (uint32_t)(resident_ctas > 0 ? resident_ctas : 1);
    const uint32_t needed = (n_pixels + block - 1u) / block;
    if (grid > needed) grid = needed;
    cudaError_t e = cudaMemsetAsync(P.pixel_counter, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
#define RDR_F(K) do { if ((e = set_smem(K, smem)) != cudaSuccess) return e; K<<<grid, block, smem, stream>>>(P); } while (0)
    RDR_RENDER_DISPATCH(mode, shape, RDR_F);
#undef RDR_F
    return cudaGetLastError();
}

// resident CTAs of render_kernel on the current device for this scene's shared-memory footprint
cudaError_t render_resident_ctas(const FrameParams &P, int variant, int *out)
{
    const int mode = mode_of(P, variant);
    const RenderShape shape = render_shape(P, mode);
    const uint32_t block = shape.block;
    const size_t smem = mode_smem_bytes(P.lay, P.staged != 0u, block, mode, shape.cold);
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
#define RDR_F(K) do { if ((e = set_smem(K, smem)) != cudaSuccess) return e; \
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, K, (int)block, smem); } while (0)
    RDR_RENDER_DISPATCH(mode, shape, RDR_F);
#undef RDR_F
    if (e != cudaSuccess) return e;
    *out = sms * (per_sm > 0 ? per_sm : 1);
    return cudaSuccess;
}

cudaError_t launch_resolve(const f4 *accum, uchar4 *out, uint32_t n_pixels, float divisor, cudaStream_t stream)
{
    if (n_pixels == 0u) return cudaSuccess;
    resolve_kernel<<<(n_pixels + 255u) / 256u, 256, 0, stream>>>(accum, out, n_pixels, divisor);
    return cudaGetLastError();
}

cudaError_t launch_peer_combine(const PeerCombine &C, cudaStream_t stream)
{
    if (C.count == 0u || C.n_src == 0u) return cudaSuccess;
    if (C.n_src > RDR_MAX_PEERS) return cudaErrorInvalidValue;
    uint32_t grid = (C.count + 255u) / 256u;
    if (grid > 148u * 8u) grid = 148u * 8u;
    switch (C.n_src) {
    case 1: peer_combine_kernel<1><<<grid, 256, 0, stream>>>(C); break;
    case 2: peer_combine_kernel<2><<<grid, 256, 0, stream>>>(C); break;
    case 3: peer_combine_kernel<3><<<grid, 256, 0, stream>>>(C); break;
    case 4: peer_combine_kernel<4><<<grid, 256, 0, stream>>>(C); break;
    case 5: peer_combine_kernel<5><<<grid, 256, 0, stream>>>(C); break;
    case 6: peer_combine_kernel<6><<<grid, 256, 0, stream>>>(C); break;
    case 7: peer_combine_kernel<7><<<grid, 256, 0, stream>>>(C); break;
    default: peer_combine_kernel<8><<<grid, 256, 0, stream>>>(C); break;
    }
    return cudaGetLastError();
}

cudaError_t launch_primary(const FrameParams &P, int variant, f4 *primary, int32_t *primary_idx, cudaStream_t stream)
{
    const uint32_t n_pixels = P.cam.width * P.cam.height;
    if (n_pixels == 0u) return cudaSuccess;
    const size_t smem = mode_smem_bytes(P.lay, P.staged != 0u, RDR_BLOCK, mode_of(P, variant));
    const uint32_t grid = (n_pixels + RDR_BLOCK - 1u) / RDR_BLOCK;
    cudaError_t e;
#define RDR_K(M, ...) do { if ((e = set_smem(primary_kernel<M>, smem)) != cudaSuccess) return e; primary_kernel<M><<<grid, RDR_BLOCK, smem, stream>>>(P, primary, primary_idx); } while (0)
    RDR_DISPATCH(mode_of(P, variant), RDR_K, 0);
#undef RDR_K
    return cudaGetLastError();
}

cudaError_t launch_kat_trace(const FrameParams &P, int variant, uint32_t n, const float *rays, int32_t *ids, float *ts,
                             cudaStream_t stream)
{
    if (n == 0u) return cudaSuccess;
    const size_t smem = mode_smem_bytes(P.lay, P.staged != 0u, RDR_BLOCK, mode_of(P, variant));
    const uint32_t grid = (n + RDR_BLOCK - 1u) / RDR_BLOCK;
    cudaError_t e;
#define RDR_K(M, ...) do { if ((e = set_smem(kat_trace_kernel<M>, smem)) != cudaSuccess) return e; kat_trace_kernel<M><<<grid, RDR_BLOCK, smem, stream>>>(P, n, rays, ids, ts); } while (0)
    RDR_DISPATCH(mode_of(P, variant), RDR_K, 0);
#undef RDR_K
    return cudaGetLastError();
}

cudaError_t launch_trace_path(const FrameParams &P, int variant, uint32_t x, uint32_t y, uint32_t sample,
                              RdrPathStep *steps, uint32_t capacity, uint32_t *n_steps, float *rgba, cudaStream_t stream)
{
    const size_t smem = mode_smem_bytes(P.lay, P.staged != 0u, 32, mode_of(P, variant), false, true);
    cudaError_t e;
#define RDR_K(M, ...) do { if ((e = set_smem(trace_path_kernel<M>, smem)) != cudaSuccess) return e; trace_path_kernel<M><<<1, 32, smem, stream>>>(P, x, y, sample, steps, capacity, n_steps, rgba); } while (0)
    RDR_DISPATCH(mode_of(P, variant), RDR_K, 0);
#undef RDR_K
    return cudaGetLastError();
}

cudaError_t launch_kat_hit(bool sphere, uint32_t n, const float *rays, const float *prims, float *t, int32_t *hit, cudaStream_t stream)
{
    if (n == 0u) return cudaSuccess;
    if (sphere) kat_hit_sphere_kernel<<<(n + 255u) / 256u, 256, 0, stream>>>(n, rays, prims, t, hit);
    else kat_hit_cube_kernel<<<(n + 255u) / 256u, 256, 0, stream>>>(n, rays, prims, t, hit);
    return cudaGetLastError();
}

cudaError_t launch_kat_vec(int op, uint32_t n, const float *in, float *out, cudaStream_t stream)
{
    if (n == 0u) return cudaSuccess;
    kat_vec_kernel<<<(n + 255u) / 256u, 256, 0, stream>>>(op, n, in, out);
    return cudaGetLastError();
}

cudaError_t launch_kat_camera_rays(const FrameParams &P, uint32_t n, const uint32_t *xy, float *rays, cudaStream_t stream)
{
    if (n == 0u) return cudaSuccess;
    kat_camera_rays_kernel<<<(n + 255u) / 256u, 256, 0, stream>>>(P, n, xy, rays);
    return cudaGetLastError();
}

cudaError_t launch_kat_rng(uint32_t seed_lo, uint32_t seed_hi, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t block,
                           uint32_t *out, cudaStream_t stream)
{
    kat_rng_kernel<<<1, 1, 0, stream>>>(seed_lo, seed_hi, pixel, sample, bounce, block, out);
    return cudaGetLastError();
}

}  // namespace rdr

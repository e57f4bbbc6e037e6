// rdr_layout.h -- scene blob layout and kernel parameters (plain structs shared by host packing,
// the kernels and the host-simulation test build).
//
// HBM / shared-memory layout.  The host packs the frame's scene into ONE contiguous blob
// (16-byte aligned sections) that every CTA stages into shared memory with a single
// cp.async.bulk (TMA 1-D bulk copy, SASS UBLKCP) completing on an mbarrier:
//
//   sphere_cull  f4[ns_pad]   (cx, cy, cz, r^2)             conservative test operands
//   cube_cull    f4[nc_pad]   (cx, cy, cz, side/2 + pad)
//   sphere_geom  f4[ns_pad]   (cx, cy, cz, radius)          exact test operands, list order
//   cube_geom    f4[nc_pad]   (cx, cy, cz, side_length)
//   obj_geom     f4[n]        (cx, cy, cz, size)            by ORIGINAL object index
//   material     f4[3*n]      (albedo.xyz, roughness) (emission.xyz, emission_strength)
//                             (metallic, transmission, ior, kind as bits)
//   sphere_idx   u32[ns_pad]  list position -> original object index
//   cube_idx     u32[nc_pad]
//
//   -- two-level ("cluster") scan, same blob --
//   top          f4[2*nt_pad] top-level entries in the hierarchy-entry format of rdr_bvh.h: a cluster of <= 8
//                             spatially close primitives (box = union) or one large primitive on its own;
//                             payload = first member position << 4 | member count
//   member_box   f4[9*n_clusters] (cx, cy, cz, +-(half-extent + pad)), negative = sphere (add the per-ray rho);
//                             8 members + 1 pad quad per cluster (144-byte stride: lanes on different clusters hit
//                             different shared-memory banks);
//                             unused slots (cx = cy = cz = 0, e = -0 is never used: count bounds the loop)
//   member_geom  f4[..]       exact-test operands (c, size) in member order
//   member_idx   u32[..]      original object index | cube bit 30
//   (within a cluster the spheres come first; the single-primitive top entries come first, spheres first)
//
//   -- the fused scan (rdr_fused.cuh) has its own clustering of the same primitives: at most 32 top-level entries,
//      so the cluster size C grows with the scene (8 up to ~250 objects, then 16, 24, 32) --
//   pair_block   f4[stride*n_top]  per cluster C/2 member pairs (A, B) x 3 quads in FFMA2 operand order, padded to an
//                             odd number of quads (lanes on different clusters spread over the banks):
//                             (cxA, cxB, cyA, cyB) (czA, czB, eA, eB) (sphereA, sphereB, desc*, 0)
//                             e = half-extent + pad (>= 0), sphere = 1.0 | 0.0 (the box grows by sphere * rho);
//                             desc (pair 0 only) = first member slot << 12 | spheres in the cluster << 6 | members
//   fused_geom   f4[C*n_top]  exact-test operands (c, size), slot = C * cluster + member (spheres first)
//   fused_idx    u32[C*n_top] original object index
//   The top-level boxes of the fused scan travel in the kernel parameters (FrameParams::top, constant bank); the
//   leading single-primitive entries (<= 4, spheres first) skip the member stage.
//
// Order in the blob: obj_geom, material and the fused scan's three sections come FIRST (SceneLayout::fused_stage_bytes):
// the fused render kernel stages only that prefix (benchmark.rscn: 20.5 KB of 37.4 KB), which leaves the shared memory
// for more resident warps; the flat-scan lists and the cluster scan's sections follow.
//
// ns_pad / nc_pad are the list lengths rounded up to 32 (one candidate-mask word per chunk).
// Every lane of a warp reads the same primitive at the same time, so all shared-memory reads in
// the scan are single-wavefront broadcasts.
#pragma once

#include "rdr_core.cuh"


namespace rdr {

// BVH mode (mode == 1) packs instead:
//   nodes        f4[16*n_nodes]   8 entries x 2 quads per node (rdr_bvh.h), per-lane traversal
//   obj_geom     f4[n]
//   material     f4[3*n]
//   nodes2       f4[16*n_nodes2]  the same primitives as a pair-packed hierarchy for the warp-cooperative traversal
//                                 (rdr_bvh2.cuh); its <= 32 root entries travel in FrameParams::top
// and is either staged the same way (small scenes) or read in place from global memory / L2 (large scenes).
struct SceneLayout {
    uint32_t mode;                       // 0 = brute-force scan lists, 1 = BVH
    uint32_t n_nodes, off_nodes;
    uint32_t bvh2_ok, bvh2_root, n_nodes2, off_nodes2;   // pair-packed hierarchy of the cooperative traversal (16 quads per node)
    uint32_t n_objects, n_spheres, n_cubes;
    uint32_t ns_pad, nc_pad;
    uint32_t off_sphere_cull, off_cube_cull, off_sphere_geom, off_cube_geom;
    uint32_t off_obj_geom, off_material, off_sphere_idx, off_cube_idx;
    uint32_t n_top, nt_pad, n_members;   // cluster scan: top entries (padded to 32), member slots
    uint32_t n_direct;                   // the first n_direct (<= 4) top entries are single primitives
    uint32_t off_top, off_member_box, off_member_geom, off_member_idx;
    // fused scan (its own clustering, see above)
    uint32_t fused_ok;                   // 1: FrameParams::top and the sections below are filled and the fused scan may run
    uint32_t fused_top;                  // top-level entries (<= 32)
    uint32_t fused_cap;                  // C: member slots per cluster (8, 16, 24 or 32)
    uint32_t fused_stride;               // quads per cluster in pair_block (3 * C/2 rounded up to odd)
    uint32_t fused_direct, fused_ns_direct;   // leading single-primitive entries (<= 4) and the spheres among them
    uint32_t off_pair_block, off_fused_geom, off_fused_idx;
    uint32_t fused_stage_bytes;          // the blob's prefix [obj_geom, material, pair_block, fused_geom, fused_idx]: all the fused kernels stage
    uint32_t blob_bytes;                 // multiple of 16
};

// Top-level boxes of the fused scan, two entries (A, B) per record in FFMA2 operand order.  They are read with
// uniform constant-bank loads (LDCU.128) straight from the kernel parameters: no shared-memory traffic in the
// uniform stage.  Unused entries have e = -1 (never hit).
constexpr uint32_t FUSED_MAX_TOP = 32u;
struct alignas(8) TopPair { float cx[2], cy[2], cz[2], ex[2], ey[2], ez[2], sphere[2]; };
struct TopParams {
    TopPair pair[FUSED_MAX_TOP / 2];
    // cooperative hierarchy only (rdr_bvh.h, Bvh2Root): root payloads and the front-to-back order tables
    uint64_t rank8[FUSED_MAX_TOP];       // bits [5 oct, 5 oct + 5): rank of root entry k for direction octant oct
    uint32_t payload[FUSED_MAX_TOP];
    uint32_t node_by_rank[8][FUSED_MAX_TOP];   // child node of the root entry with that rank (0xffffffff: none)
    uint32_t prim_mask, cube_mask;       // which root entries are primitives / cubes
};

struct FrameParams {
    Camera cam;
    World world;
    CullConsts cull;
    SceneLayout lay;
    const unsigned char *blob;           // device copy of the packed scene
    uint32_t staged;                     // 1: every CTA stages the blob into shared memory; 0: read in place (L2)
    f4 *accum;                           // W*H float RGBA accumulator (row-major, top row first)
    // primary table of the frame (primary_kernel): the camera ray has no jitter (cpu.rs:199-202), so a pixel's primary
    // direction and nearest hit are the same for every sample of every launch
    const f4 *primary;                   // W*H: primary ray direction (x, y, z), nearest-hit t (w)
    const int32_t *primary_idx;          // W*H: nearest-hit object index, -1 = miss
    uint32_t *pixel_counter;             // next unclaimed pixel of this launch (zeroed before the launch)
    uint32_t seed_lo, seed_hi;
    uint32_t max_bounces;
    uint32_t sample_begin;               // global index of the first sample of this launch
    uint32_t sample_count;               // samples per pixel in this launch
    // row-stripe sharding (rdr_set_row_stripes): this launch renders only the stripes s of stripe_rows image rows with
    // s % stripe_count == stripe_index.  stripe_count <= 1: the whole image.  owned_pixels = pixels handed out.
    uint32_t stripe_rows, stripe_index, stripe_count;
    uint32_t owned_pixels;
    TopParams top;                       // fused scan only (lay.fused_ok)
};

// ---- row-stripe ownership (image-tile sharding across GPUs, SURVEY.md 8e "alternative") --------------------
// Stripe s covers image rows [s * rows, min((s + 1) * rows, H)); shard `index` of `count` owns the stripes with
// s % count == index (round-robin, so an unevenly lit image still balances).  Pixels are row-major, so a stripe
// is one contiguous range of pixel indices and the k-th owned pixel is found with one division.
RDR_HD uint32_t stripe_owned_pixels(uint32_t width, uint32_t height, uint32_t rows, uint32_t index, uint32_t count)
{
    if (count <= 1u || rows == 0u) return width * height;
    const uint32_t n_stripes = (height + rows - 1u) / rows;
    uint32_t owned_rows = 0u;
    for (uint32_t s = index; s < n_stripes; s += count) {
        const uint32_t r0 = s * rows;
        owned_rows += (height - r0 < rows) ? height - r0 : rows;
    }
    return owned_rows * width;
}

// global pixel index (y * W + x) of the k-th pixel this shard owns, k < owned_pixels.  Only the last stripe of the
// image can be short, and a shard that owns it owns it last, so every earlier owned stripe is full.
RDR_HD uint32_t stripe_pixel(uint32_t width, uint32_t rows, uint32_t index, uint32_t count, uint32_t k)
{
    if (count <= 1u || rows == 0u) return k;
    const uint32_t per = rows * width;
    const uint32_t s = k / per;
    return (s * count + index) * per + (k - s * per);
}

struct Hit { int idx; float t; };

}  // namespace rdr

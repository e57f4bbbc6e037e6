/*
 * raydar_cuda.h -- C ABI of libraydar_cuda.so, the B200 (sm_100a) CUDA backend for Raydar's
 * per-pixel path-tracing sample loop.
 *
 * This is the drop-in boundary: the entry points mirror, one for one, what a Rust
 * `impl Renderer for CudaRenderer` (reference: src/renderer/mod.rs:25-35) has to call, next to
 * the existing `CpuRenderer` (src/renderer/cpu.rs:118-183) and `VulkanRenderer`
 * (src/renderer/vulkan.rs:100-455).  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *   - every function returns an RdrStatus (0 = ok); rdr_last_error() gives the message.
 *     Nothing in the library aborts the process (the reference panics instead: unwrap/todo!).
 *   - a handle is NOT thread-safe: one caller thread, like `&mut self` in the trait.
 *   - the scene is BORROWED per call and snapshotted into device memory by rdr_new_frame
 *     (the VulkanRenderer model, vulkan.rs:207-428).
 *   - images are row-major, top row first, RGBA, 8-bit linear, alpha 255 -- the layout of
 *     image::RgbaImage that the trait returns.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     RDR_ERR_CUDA.
 */
#ifndef RAYDAR_CUDA_H
#define RAYDAR_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum RdrStatus {
    RDR_OK = 0,
    RDR_ERR_INVALID = 1,       /* bad argument / call order (e.g. render before new_frame) */
    RDR_ERR_CUDA = 2,          /* CUDA runtime error or no device */
    RDR_ERR_UNSUPPORTED = 3,   /* World::Transparent is todo!() in the reference (world.rs:32) */
    RDR_ERR_IO = 4,            /* scene file cannot be opened / image cannot be written */
    RDR_ERR_PARSE = 5,         /* scene file is not a valid .rscn */
    RDR_ERR_NCCL = 6,
    RDR_ERR_NOMEM = 7
} RdrStatus;

enum { RDR_SPHERE = 0, RDR_CUBE = 1 };                                   /* scene/objects.rs:6-10 */
enum { RDR_WORLD_SKY = 0, RDR_WORLD_SOLID = 1, RDR_WORLD_TRANSPARENT = 2 }; /* scene/world.rs:6-14 */
enum { RDR_LOBE_MISS = 0, RDR_LOBE_DIFFUSE = 1, RDR_LOBE_SPECULAR = 2, RDR_LOBE_REFRACT = 3 };
enum { RDR_MAT_STRIDE = 11 };  /* albedo3, roughness, metallic, emission_color3, emission_strength,
                                  transmission, ior -- field order of scene/material.rs:4-13 */

/* nearest-hit search strategy.  All return the reference's trace_ray winner (cpu.rs:344-352):
 *   BRUTE    flat scan over the shared-memory SoA primitive buffer (every primitive, every ray)
 *   CLUSTER  two-level scan of the same buffer: a uniform scan over <= 128 cluster boxes, then the members
 *            of the clusters a ray touches
 *   BVH      8-wide hierarchy (shared memory when it fits, otherwise global memory / L2)
 *   COOP     CLUSTER with the per-lane stages regrouped across the warp (ballot/shuffle work distribution)
 *   FUSED    COOP rebuilt for the sm_100a issue model: packed FFMA2 box tests, top-level boxes in the constant
 *            bank, atomics-free survivor compaction, lane state the search does not need in shared-memory columns
 *            (<= 32 top-level entries of 8 / 16 / 24 / 32 members, i.e. up to ~1000 objects; COOP otherwise)
 *   BVH_COOP the hierarchy in pair-packed form (<= 32 root entries, 8-wide nodes), traversed by the whole warp: one
 *            near-first stack per ray, four lanes per node (FFMA2 pair tests), pruning by the best exact hit so far
 *   AUTO     FUSED / COOP up to 1024 objects, BVH_COOP above */
enum { RDR_ACCEL_AUTO = 0, RDR_ACCEL_BRUTE = 1, RDR_ACCEL_BVH = 2, RDR_ACCEL_CLUSTER = 3, RDR_ACCEL_COOP = 4, RDR_ACCEL_FUSED = 5, RDR_ACCEL_BVH_COOP = 6 };

typedef struct RdrRenderer RdrRenderer;    /* replaces CpuRenderer state, cpu.rs:111-116 */
typedef struct RdrScene RdrScene;          /* host-side loaded scene, scene/mod.rs:13-18 */

/* RendererConfig, renderer/mod.rs:11-23 (defaults 1024 / 12) */
typedef struct RdrConfig {
    uint32_t max_sample_count;
    uint32_t max_bounces;
} RdrConfig;

/* Flattened `&Scene`: what the Rust shim builds from scene.camera / scene.world / scene.objects. */
typedef struct RdrSceneFlat {
    uint32_t width, height;          /* camera.resolution_x/y, camera.rs:19-20 */
    float inv_proj[16];              /* camera.inverse_proj_matrix(), column-major (camera.rs:206) */
    float inv_view[16];              /* camera.inverse_view_matrix(), column-major (camera.rs:202) */
    float cam_pos[3];                /* camera.position(), camera.rs:62 */
    uint32_t world_kind;             /* RDR_WORLD_* */
    float world_a[3];                /* SkyColor.top_color | SolidColor */
    float world_b[3];                /* SkyColor.bottom_color */
    uint32_t n_objects;
    const uint32_t *kind;            /* n: RDR_SPHERE | RDR_CUBE */
    const float *geom;               /* n*4: center.xyz, radius | side_length */
    const float *material;           /* n*RDR_MAT_STRIDE */
} RdrSceneFlat;

/* Profiler, renderer/timing.rs:10-19: the four Timer durations in nanoseconds; a timer whose
 * duration() would be None reports has_* = 0 (main.rs:61-107 errors out on those). */
typedef struct RdrProfiler {
    uint64_t frame_ns, sample_ns, prepare_ns, render_ns;
    uint32_t has_frame, has_sample, has_prepare, has_render;
    double   device_render_ms;       /* extension: CUDA-event time of the sample kernels of this frame */
} RdrProfiler;

/* One bounce of a path (debug / parity mode). */
typedef struct RdrPathStep {
    int32_t  object;                 /* hit object index, -1 = miss */
    uint32_t lobe;                   /* RDR_LOBE_* */
    uint32_t front_face;
    float t;
    float position[3];
    float normal[3];
    float origin[3];                 /* next ray origin */
    float direction[3];              /* next ray direction */
    float attenuation[3];
    float light[3];
} RdrPathStep;

/* ---- lifecycle: CpuRenderer::new(config), cpu.rs:186 ------------------------------------------ */
int rdr_create(const RdrConfig *config, int device, RdrRenderer **out);
void rdr_destroy(RdrRenderer *r);
const char *rdr_last_error(const RdrRenderer *r);   /* r == NULL: last error of a failed rdr_create / scene call */

/* ---- trait Renderer, renderer/mod.rs:25-35 ---------------------------------------------------- */
/* new_frame(&mut self, &Scene)                       cpu.rs:135-140 */
int rdr_new_frame(RdrRenderer *r, const RdrSceneFlat *scene);
/* render_sample(&mut self, &Scene) -> Option<RgbaImage>   cpu.rs:142-158.
 * *produced = 0 (image untouched) once sample_count >= max_sample_count, i.e. `None`. */
int rdr_render_sample(RdrRenderer *r, uint8_t *rgba8, int *produced);
/* render_frame(&mut self, &Scene) -> RgbaImage       cpu.rs:119-133 (new_frame + all samples + resolve) */
int rdr_render_frame(RdrRenderer *r, const RdrSceneFlat *scene, uint8_t *rgba8);
/* The second half of render_frame on its own: every sample the current frame has left (cpu.rs:126-128) + the image
 * (cpu.rs:129).  rdr_render_frame == rdr_new_frame + rdr_finish_frame; a host that keeps its own Profiler (the Rust
 * shim: prepare_timer around the first call, render_timer around the second) calls the halves. */
int rdr_finish_frame(RdrRenderer *r, uint8_t *rgba8);
/* profiler(&self) -> &Profiler */
int rdr_profiler(const RdrRenderer *r, RdrProfiler *out);
uint32_t rdr_sample_count(const RdrRenderer *r);
uint32_t rdr_max_sample_count(const RdrRenderer *r);
uint32_t rdr_max_bounces(const RdrRenderer *r);
/* On a multi-GPU handle a new max_sample_count takes effect at the next rdr_new_frame (the devices' shares of the
 * frame's samples are fixed there); max_bounces applies to the next launch, as on one GPU. */
int rdr_set_max_sample_count(RdrRenderer *r, uint32_t count);
int rdr_set_max_bounces(RdrRenderer *r, uint32_t bounces);

/* ---- extensions the reference lacks (seed, sharding, throughput path) -------------------------- */
/* RNG seed (Philox key); the reference's thread_rng has none.  Takes effect at the next new_frame. */
int rdr_set_seed(RdrRenderer *r, uint64_t seed);
/* This instance renders global sample indices [first, first + max_sample_count): sample-range
 * sharding across GPUs (one instance per GPU, accumulators summed afterwards). */
int rdr_set_sample_offset(RdrRenderer *r, uint32_t first_sample);
/* This instance renders only the row stripes s (stripe_rows image rows each, top stripe = 0) with
 * s % count == index: image-tile sharding across GPUs.  The stripes of different instances are disjoint, so the
 * sum of their accumulators (x + 0 = x) is BIT-IDENTICAL to the one-instance image.  count <= 1 or
 * stripe_rows == 0: the whole image.  Takes effect at the next launch; the accumulator stays W*H. */
int rdr_set_row_stripes(RdrRenderer *r, uint32_t stripe_rows, uint32_t index, uint32_t count);
int rdr_set_accel(RdrRenderer *r, int accel);                 /* RDR_ACCEL_*; takes effect at the next rdr_new_frame */
/* Zero the accumulator and sample_count but keep the scene already resident on the device
 * (new_frame without the scene upload; the editor calls new_frame for every change, the
 * throughput path only needs a fresh accumulator). */
int rdr_reset_frame(RdrRenderer *r);
/* Render up to n more samples into the device accumulator, no resolve, no read-back. */
int rdr_render_samples(RdrRenderer *r, uint32_t n);
/* print_frame_buffer, cpu.rs:221-230: clamp(sum / divisor, 0, 1) * 255 as u8, read back to host.
 * divisor = 0 uses sample_count(). */
int rdr_resolve(RdrRenderer *r, uint32_t divisor, uint8_t *rgba8);
int rdr_read_accum(RdrRenderer *r, float *accum_rgba_f32);    /* W*H*4 floats */
/* The inverse: restore the current frame's accumulation state (W*H*4 floats = sums over `sample_count` samples) -- resume
 * a progressive render from a saved accumulator: the next launch continues with sample index sample_count, exactly as
 * if the samples had been rendered by this handle (the reference keeps this state only in memory, cpu.rs:113-114). */
int rdr_write_accum(RdrRenderer *r, const float *accum_rgba_f32, uint32_t sample_count);
/* device accumulator (float RGBA, W*H*4) and the stream the kernels run on, for an external
 * NCCL reduce (torch.distributed / ncclReduce) between per-GPU instances */
int rdr_accum_device_ptr(RdrRenderer *r, void **ptr, size_t *bytes);
int rdr_stream(RdrRenderer *r, void **cuda_stream);
int rdr_synchronize(RdrRenderer *r);
/* number of kernels launched by this handle so far */
uint64_t rdr_launch_count(const RdrRenderer *r);
/* bytes of the packed scene that rdr_new_frame uploads to the device for the current frame (0 without a frame) */
uint64_t rdr_scene_device_bytes(const RdrRenderer *r);

/* single-process multi-GPU: one sub-renderer per device, sample ranges split evenly, accumulators
 * combined by the fused peer-memory reduce + resolve (rdr_set_combine) or one ncclReduce(sum, f32) onto devices[0].
 * Progressive calls: rdr_render_sample adds exactly ONE sample per call, as the trait does (cpu.rs:142-158) -- with the
 * STRIPES partition every device renders it for its own rows (latency / G), with SAMPLES the devices take turns. */
int rdr_create_multi(const RdrConfig *config, int n_devices, const int *devices, RdrRenderer **out);
/* How a multi-GPU handle combines the per-GPU accumulators into the image:
 *   PEER  fused reduce + resolve over NVLink peer memory: every GPU sums ITS 1/G of the pixels from all accumulators
 *         (fixed order, device 0 first), quantises (cpu.rs:221-230) and writes its RGBA8 slice straight into the host
 *         image -- no rooted reduce, no separate resolve or copy.  With the STRIPES partition a GPU only reads its own
 *         accumulator.  Needs peer access between all devices of the handle.
 *   NCCL  one ncclReduce(sum, f32) onto devices[0], then resolve and copy there (the fallback, and the check of PEER)
 *   AUTO  PEER when every pair of devices has peer access, else NCCL (default) */
enum { RDR_COMBINE_AUTO = 0, RDR_COMBINE_PEER = 1, RDR_COMBINE_NCCL = 2 };
int rdr_set_combine(RdrRenderer *r, int combine);
int rdr_combine_in_use(const RdrRenderer *r);                 /* RDR_COMBINE_PEER | RDR_COMBINE_NCCL; AUTO for one device */
/* Pinned host memory for images, visible to every device: an rgba8 argument that lies in such a buffer is written by
 * the GPUs directly (PEER combine) or by one asynchronous DMA; any other pointer is accepted too and costs a staging
 * copy.  The Rust shim keeps one per renderer and copies into the RgbaImage it returns. */
int rdr_alloc_host_image(size_t bytes, uint8_t **out);
void rdr_free_host_image(uint8_t *image);

/* ---- one process per GPU (torchrun / MPI style): the same fused combine over CUDA IPC ------------------------- */
/* Every rank exports its accumulator and device image (RDR_IPC_HANDLE_BYTES each, after rdr_new_frame), the ranks
 * exchange the handles (all-gather) and attach; rdr_peer_combine then sums this rank's 1/world of the pixels over all
 * ranks' accumulators and writes the RGBA8 slice into rank 0's device image (rdr_read_image on rank 0 reads it back).
 * The caller orders the ranks: all renders finished before any combine, all combines before the next frame's reset
 * (two barriers).  Detach on every rank before a resolution change. */
enum { RDR_IPC_HANDLE_BYTES = 128 };
int rdr_ipc_export(RdrRenderer *r, void *handle);
int rdr_peer_attach(RdrRenderer *r, uint32_t rank, uint32_t world, const void *handles /* world * RDR_IPC_HANDLE_BYTES */);
int rdr_peer_combine(RdrRenderer *r, uint32_t divisor);
int rdr_peer_detach(RdrRenderer *r);
int rdr_read_image(RdrRenderer *r, uint8_t *rgba8);          /* the device image (last resolve / combine) -> host */

/* how a multi-GPU handle splits the frame (takes effect at the next new_frame):
 *   SAMPLES  every device renders the whole image for its share of the sample indices (default; perfect balance,
 *            result equal to one GPU up to f32 summation order)
 *   STRIPES  every device renders all samples of its round-robin row stripes (stripe_rows rows each, 0 = 16):
 *            the reduce adds zeros, so the result is bit-identical to one GPU */
enum { RDR_PARTITION_SAMPLES = 0, RDR_PARTITION_STRIPES = 1 };
int rdr_set_partition(RdrRenderer *r, int partition, uint32_t stripe_rows);

/* ---- parity / debug modes ------------------------------------------------------------------------ */
/* first-hit object id (-1 = miss) and t per pixel for the current frame's primary rays */
int rdr_first_hit(RdrRenderer *r, int32_t *ids, float *t);
/* one path with every bounce recorded; *n_steps <= capacity; rgba = per_pixel() result */
int rdr_trace_path(RdrRenderer *r, uint32_t x, uint32_t y, uint32_t sample,
                   RdrPathStep *steps, uint32_t capacity, uint32_t *n_steps, float rgba[4]);
/* per-function known-answer entry points (device execution of the same __device__ functions the
 * render kernel uses): rays n*6 (origin, direction), prims n*4 */
int rdr_kat_hit_sphere(RdrRenderer *r, uint32_t n, const float *rays, const float *spheres, float *t, int32_t *hit);
int rdr_kat_hit_cube(RdrRenderer *r, uint32_t n, const float *rays, const float *cubes, float *t, int32_t *hit);
/* nearest hit of n arbitrary rays against the current frame's scene: id (-1 miss), t */
int rdr_kat_trace(RdrRenderer *r, uint32_t n, const float *rays, int32_t *ids, float *t);
/* camera rays of the current frame for n pixels (xy n*2 u32) -> rays n*6 */
int rdr_kat_camera_rays(RdrRenderer *r, uint32_t n, const uint32_t *xy, float *rays);
/* the shading helpers, n records of 12 floats in -> 8 floats out (unused slots 0):
 *   REFLECT       utils/mod.rs:14-16    in v[3] n[3]                               out r[3]
 *   REFRACT       utils/mod.rs:25-35    in v[3] n[3] ratio                         out r[3]
 *   CAN_REFRACT   utils/mod.rs:37-44    in v[3] n[3] ratio                         out flag
 *   WORLD_SAMPLE  world.rs:17-34        in d[3] kind(0 sky, 1 solid) a[3] b[3]     out rgb[3]
 *   CLOSEST_HIT   cpu.rs:354-394        in o[3] d[3] t is_sphere c[3] size         out p[3] n[3] front
 *   QUANTISE      cpu.rs:224-228        in sum n                                   out the u8 value
 *   RAND_FLOATS   utils/mod.rs:47-55, cpu.rs:280 (rand 0.8.5 float maps)  in three raw u32 words as bit patterns
 *                                       out random::<f32>(w0), gen_range(-1.0..=1.0)(w0), random_in_unit_sphere(w0, w1, w2)[3] */
enum { RDR_KAT_REFLECT = 0, RDR_KAT_REFRACT = 1, RDR_KAT_CAN_REFRACT = 2, RDR_KAT_WORLD_SAMPLE = 3, RDR_KAT_CLOSEST_HIT = 4, RDR_KAT_QUANTISE = 5,
       RDR_KAT_RAND_FLOATS = 6 };
int rdr_kat_vec(RdrRenderer *r, int op, uint32_t n, const float *in, float *out);
/* raw RNG block of the spec (Philox4x32-10), computed on the device */
int rdr_kat_rng(RdrRenderer *r, uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t block, uint32_t out[4]);

/* ---- host scene pipeline (cli/mod.rs:32-40, scene/camera.rs:210-231) ----------------------------- */
int rdr_scene_load_rscn(const char *path, RdrScene **out);
int rdr_scene_default(RdrScene **out);                            /* Scene::default(), scene/mod.rs:20-68 */
/* set_resolution_x/y + update_matrices (camera.rs:141-157,210-231) */
int rdr_scene_set_resolution(RdrScene *s, uint32_t width, uint32_t height);
/* the same, but keeps the stored matrices (valid when the aspect ratio is unchanged) */
int rdr_scene_override_resolution(RdrScene *s, uint32_t width, uint32_t height);
int rdr_scene_flat(const RdrScene *s, RdrSceneFlat *out);        /* pointers valid until rdr_scene_free */
void rdr_scene_free(RdrScene *s);
/* RgbaImage::save (main.rs:19): 8-bit RGBA PNG */
int rdr_write_png(const char *path, const uint8_t *rgba8, uint32_t width, uint32_t height);

const char *rdr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RAYDAR_CUDA_H */

/*
 * raydar_oracle.h -- CPU restatement of bvpav/raydar's CPU backend (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity ORACLE for the CUDA path-tracing sample loop.  It is a plain-C,
 * scalar, f32-only restatement of the reference's algorithm:
 *
 *   src/renderer/cpu.rs:17-99    Ray, hit_sphere, hit_cube
 *   src/renderer/cpu.rs:193-230  render_next_sample, print_frame_buffer
 *   src/renderer/cpu.rs:233-342  per_pixel (camera ray, bounce loop, scatter)
 *   src/renderer/cpu.rs:344-398  trace_ray, closest_hit, miss
 *   src/utils/mod.rs:14-55       reflect, refract, can_refract, random_in_unit_sphere
 *   src/scene/world.rs:17-34     World::sample
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libraydar_cuda.so) never links, loads or
 * calls it and has no CPU fallback.
 *
 * PARITY UNPINNED: the reference is Rust (no cargo/rustc in this image, no vendored
 * crates), has no tests, no golden vectors and a non-deterministic RNG
 * (rand::thread_rng), so this restatement cannot be checked against reference outputs.
 * It is cross-checked instead against an independent numpy f32 restatement
 * (tests/np_restatement.py) and against the survey's scene anchors.
 *
 * Third-party arithmetic restated from the pinned crate versions (Cargo.lock):
 *   cgmath 0.18.0  dot = (x*x' + y*y') + z*z';  normalize = v * (1/sqrt(dot(v,v)));
 *                  Matrix4*Vector4 = ((c0*v0 + c1*v1) + c2*v2) + c3*v3;  lerp = a + (b-a)*t
 *   rand 0.8.5     random::<f32>()      = (u32 >> 8) * 2^-24
 *                  gen_range(-1.0..=1.0) = ((u32 >> 9) * 2^-23) * scale + (-1),
 *                                          scale = 2/(1-2^-23) as computed by UniformFloat::new_inclusive
 *   ordered-float 4.6.0  total order with NaN greatest; Iterator::min_by_key keeps the FIRST minimum
 *   compiler-rt __powisf2 / LLVM powi expansion: x^2 = x*x, x^5 = x*((x*x)*(x*x))
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (Rust never contracts to FMA).
 *
 * RNG: the reference draws from an OS-seeded ChaCha12 and cannot be reproduced.  Oracle
 * and CUDA path share a counter-based SPEC instead (implemented twice, independently):
 *   Philox4x32-10, key = (seed_lo, seed_hi), counter = (pixel_index, sample_index,
 *   bounce*4 + block, 0); pixel_index = y*W + x.  Per bounce:
 *     block 0: word0 -> u1 (transmission test), word1 -> u2 (fresnel / metallic test),
 *              word2 -> u3 (roughness test), word3 unused
 *     block 1: words 0..2 -> random_in_unit_sphere() #1 (diffuse)
 *     block 2: words 0..2 -> random_in_unit_sphere() #2 (specular offset)
 *     block 3: words 0..2 -> random_in_unit_sphere() #3 (refraction offset)
 *   The reference's draws are i.i.d., so assigning each call site its own slot keeps
 *   the distribution of every sample identical.
 */
#ifndef RAYDAR_ORACLE_H
#define RAYDAR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_SPHERE = 0, ORC_CUBE = 1 };
enum { ORC_WORLD_SKY = 0, ORC_WORLD_SOLID = 1, ORC_WORLD_TRANSPARENT = 2 };
enum { ORC_LOBE_MISS = 0, ORC_LOBE_DIFFUSE = 1, ORC_LOBE_SPECULAR = 2, ORC_LOBE_REFRACT = 3 };

/* material layout: 11 floats per object, the field order of src/scene/material.rs:4-13 */
enum { ORC_MAT_ALBEDO = 0, ORC_MAT_ROUGHNESS = 3, ORC_MAT_METALLIC = 4, ORC_MAT_EMISSION = 5,
       ORC_MAT_EMISSION_STRENGTH = 8, ORC_MAT_TRANSMISSION = 9, ORC_MAT_IOR = 10, ORC_MAT_STRIDE = 11 };

typedef struct {
    uint32_t width, height;       /* camera.rs:19-20 */
    float inv_proj[16];           /* column-major, as stored in .rscn (camera.rs:29) */
    float inv_view[16];           /* camera.rs:28 */
    float cam_pos[3];             /* camera.rs:15 */
    uint32_t world_kind;          /* world.rs:6-14 */
    float world_a[3];             /* SkyColor.top_color | SolidColor */
    float world_b[3];             /* SkyColor.bottom_color */
    uint32_t n_objects;
    const uint32_t *kind;         /* n: ORC_SPHERE | ORC_CUBE (objects.rs:6-10) */
    const float *geom;            /* n*4: cx,cy,cz, radius | side_length (objects.rs:40-50) */
    const float *material;        /* n*11 */
} OrcScene;

/* one bounce of a path, for the fixed-seed single-path debug comparison */
typedef struct {
    int32_t  object;              /* hit object index, -1 = miss */
    uint32_t lobe;                /* ORC_LOBE_* */
    uint32_t front_face;
    float t;
    float position[3];
    float normal[3];
    float origin[3];              /* next ray origin */
    float direction[3];           /* next ray direction (un-normalised for diffuse) */
    float attenuation[3];         /* after this bounce */
    float light[3];               /* after this bounce */
} OrcPathStep;

typedef struct {
    uint64_t samples;
    uint64_t trace_calls;
    uint64_t primitive_tests;
    uint64_t alive_at_bounce[64];
    uint64_t lobe_count[4];
    uint64_t exhausted;           /* paths that used all max_bounces */
} OrcStats;

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_rng_block(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t block, uint32_t out[4]);
float orc_u01(uint32_t word);                 /* rand::random::<f32>() */
float orc_range_pm1(uint32_t word);           /* gen_range(-1.0..=1.0) */
void orc_random_in_unit_sphere(const uint32_t words[3], float out[3]);   /* utils/mod.rs:47-55 */

void orc_camera_ray(const OrcScene *s, uint32_t x, uint32_t y, float origin[3], float dir[3]);  /* cpu.rs:199-251 */
int  orc_hit_sphere(const float o[3], const float d[3], const float center[3], float radius, float *t);   /* cpu.rs:34-62 */
int  orc_hit_cube(const float o[3], const float d[3], const float center[3], float side, float *t);       /* cpu.rs:64-98 */
int  orc_trace(const OrcScene *s, const float o[3], const float d[3], float *t);                          /* cpu.rs:344-352; -1 = miss */
void orc_closest_hit(const OrcScene *s, int obj, const float o[3], const float d[3], float t,
                     float pos[3], float normal[3], uint32_t *front_face);                                /* cpu.rs:354-394 */
void orc_reflect(const float v[3], const float n[3], float out[3]);                                       /* utils/mod.rs:14-16 */
void orc_refract(const float v[3], const float n[3], float ratio, float out[3]);                          /* utils/mod.rs:25-35 */
int  orc_can_refract(const float v[3], const float n[3], float ratio);                                    /* utils/mod.rs:37-44 */
void orc_world_sample(const OrcScene *s, const float d[3], float out[3]);                                 /* world.rs:17-34 */

/* batched forms for hypothesis/KAT tests: rays n*6 (o,d), prims n*4, out_t n, out_hit n */
void orc_hit_sphere_batch(uint32_t n, const float *rays, const float *spheres, float *out_t, int32_t *out_hit);
void orc_hit_cube_batch(uint32_t n, const float *rays, const float *cubes, float *out_t, int32_t *out_hit);

/* first-hit object id (-1 miss) and t per pixel, primary rays only */
void orc_first_hit(const OrcScene *s, int32_t *ids, float *ts, int n_threads);

/* one path, recording every bounce; returns number of steps written (<= max_bounces); rgba = per_pixel() result */
uint32_t orc_trace_path(const OrcScene *s, uint32_t x, uint32_t y, uint32_t sample, uint64_t seed,
                        uint32_t max_bounces, OrcPathStep *steps, float rgba[4]);

/* accum[W*H*4] += samples [sample_begin, sample_end) in order (cpu.rs:193-219); rows optional sub-range.
 * n_threads == 1 is the reference-faithful single-threaded loop; > 1 splits rows with OpenMP. */
void orc_render(const OrcScene *s, uint64_t seed, uint32_t sample_begin, uint32_t sample_end,
                uint32_t max_bounces, uint32_t row_begin, uint32_t row_end, float *accum,
                int n_threads, OrcStats *stats);

/* Oracle-side BVH: NOT part of the reference (its trace_ray is the linear scan above).  Returns orc_trace's winner
 * (checked in tests/test_oracle_bvh.py); exists so that a CPU path can be timed at 100k objects.  tests: exact
 * primitive tests performed (may be NULL). */
typedef struct OrcBvh OrcBvh;
OrcBvh *orc_bvh_build(const OrcScene *s);
void orc_bvh_free(OrcBvh *b);
int  orc_trace_bvh(const OrcScene *s, const OrcBvh *b, const float o[3], const float d[3], float *t, uint64_t *tests);
/* orc_render for the pixel rectangle [x0, x0+w) x [y0, y0+h) into a w*h*4 buffer; bvh NULL = linear scan */
void orc_render_region(const OrcScene *s, const OrcBvh *bvh, uint64_t seed, uint32_t sample_begin, uint32_t sample_end,
                       uint32_t max_bounces, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h, float *accum,
                       int n_threads, OrcStats *stats);

/* print_frame_buffer (cpu.rs:221-230) */
void orc_resolve(const float *accum, uint64_t n_pixels, uint32_t sample_count, uint8_t *rgba8);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif

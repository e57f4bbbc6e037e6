// rdr_core.cuh -- per-lane device functions of the path-tracing sample loop (sm_100a).
//
// Everything a lane does to one ray lives here: RNG, camera ray, exact intersections, the
// conservative cull tests, closest_hit, scatter, world sample, resolve quantisation.
//
// Floating point policy.  Raydar's CPU backend is plain f32 Rust: one IEEE rounding per
// operation, never an FMA.  First-hit ids must be bit-exact against it, so every operation whose
// result can reach an output is written with an explicit rounding intrinsic (__fmul_rn,
// __fadd_rn, __fdiv_rn, __fsqrt_rn): nvcc never contracts or reassociates those.  FMAs appear
// only where they are written out (fma()), in the conservative cull tests, whose results
// only decide which primitives get the exact test.
//
// The functions are __host__ __device__ so that tests/hostsim can compile this very header with
// g++ -ffp-contract=off -mfma and check the per-lane logic on a machine without a GPU.  That
// build is test infrastructure; libraydar_cuda.so never runs it.
#pragma once

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RDR_HD __host__ __device__ __forceinline__
#define RDR_HD_NOINLINE __host__ __device__ __noinline__
#define RDR_UNROLL _Pragma("unroll")
#define RDR_NOUNROLL _Pragma("unroll 1")
#define RDR_UNROLL4 _Pragma("unroll 4")
#else
#define RDR_UNROLL
#define RDR_NOUNROLL
#define RDR_UNROLL4
#define RDR_HD_NOINLINE inline
#define RDR_HD inline
#include <math.h>
#endif

// The render kernel's hot loop is ~34 KB of code against a 32 KB L1.5 instruction cache, so helpers that
// are used at several call sites of the shading code are kept OUT OF LINE on the device: one copy of Philox and one of
// normalize instead of two and four (measured on B200: +3.3 %, profiles/variants_r03i.txt).
#define RDR_HD_SHARED RDR_HD_NOINLINE

namespace rdr {

// ---- one-rounding IEEE f32 operations ---------------------------------------------------------
#if defined(__CUDA_ARCH__)
RDR_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
RDR_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
RDR_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
RDR_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
RDR_HD float fsqrt(float a) { return __fsqrt_rn(a); }
RDR_HD float frcp(float a) { return __frcp_rn(a); }         // == 1.0f / a correctly rounded (IEEE division of 1 by a), shorter than div.rn
RDR_HD float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
// PTX min.f32 / max.f32: a NaN operand is ignored, -0 < +0 (Rust f32::min/max ignore NaN too)
RDR_HD float fmin(float a, float b) { return fminf(a, b); }
RDR_HD float fmax(float a, float b) { return fmaxf(a, b); }
RDR_HD uint32_t f2u(float f) { return __float_as_uint(f); }
RDR_HD float u2f(uint32_t u) { return __uint_as_float(u); }
RDR_HD uint32_t mulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
RDR_HD int ffs32(uint32_t m) { return __ffs((int)m) - 1; }
#else
RDR_HD float fadd(float a, float b) { return a + b; }
RDR_HD float fsub(float a, float b) { return a - b; }
RDR_HD float fmul(float a, float b) { return a * b; }
RDR_HD float fdiv(float a, float b) { return a / b; }
RDR_HD float fsqrt(float a) { return __builtin_sqrtf(a); }
RDR_HD float frcp(float a) { return 1.0f / a; }
RDR_HD float fma(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
RDR_HD float fmin(float a, float b)
{
    if (a != a) return b;
    if (b != b) return a;
    if (a == 0.0f && b == 0.0f) return __builtin_signbit(a) ? a : b;
    return a < b ? a : b;
}
RDR_HD float fmax(float a, float b)
{
    if (a != a) return b;
    if (b != b) return a;
    if (a == 0.0f && b == 0.0f) return __builtin_signbit(a) ? b : a;
    return a > b ? a : b;
}
RDR_HD uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
RDR_HD float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
RDR_HD uint32_t mulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
RDR_HD int ffs32(uint32_t m) { return __builtin_ffs((int)m) - 1; }
#endif

RDR_HD float fabs_(float a) { return u2f(f2u(a) & 0x7fffffffu); }
RDR_HD float fneg(float a) { return u2f(f2u(a) ^ 0x80000000u); }
RDR_HD bool isnan_(float a) { return a != a; }
RDR_HD float finf() { return u2f(0x7f800000u); }

struct v3 { float x, y, z; };
// 16-byte aligned quad: one LDS.128 / LDG.128 / STG.128 per access on the device
struct alignas(16) f4 { float x, y, z, w; };
RDR_HD v3 mk3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
RDR_HD v3 add3(v3 a, v3 b) { return mk3(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }
RDR_HD v3 sub3(v3 a, v3 b) { return mk3(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }
RDR_HD v3 mul3(v3 a, v3 b) { return mk3(fmul(a.x, b.x), fmul(a.y, b.y), fmul(a.z, b.z)); }
RDR_HD v3 scale3(v3 a, float s) { return mk3(fmul(a.x, s), fmul(a.y, s), fmul(a.z, s)); }
RDR_HD v3 neg3(v3 a) { return mk3(fneg(a.x), fneg(a.y), fneg(a.z)); }
// cgmath 0.18 InnerSpace::dot: (x*x' + y*y') + z*z'
RDR_HD float dot3(v3 a, v3 b) { return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z)); }
// cgmath normalize: v * (1 / sqrt(dot(v, v)))
RDR_HD_SHARED v3 normalize3(v3 a) { return scale3(a, frcp(fsqrt(dot3(a, a)))); }

// ---- RNG spec: Philox4x32-10, key = seed, counter = (pixel, sample, bounce*4 + block, 0) -----
struct u4 { uint32_t x, y, z, w; };

RDR_HD u4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    RDR_UNROLL
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = mulhi(M1, c2), lo1 = M1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    u4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

RDR_HD_SHARED u4 rng_block(uint32_t seed_lo, uint32_t seed_hi, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t block)
{
    return philox4x32_10(pixel, sample, bounce * 4u + block, 0u, seed_lo, seed_hi);
}

// rand 0.8.5: random::<f32>() = (u32 >> 8) * 2^-24
RDR_HD float u01(uint32_t w) { return fmul((float)(w >> 8), 5.9604644775390625e-08f); }
// rand 0.8.5: gen_range(-1.0..=1.0) = ((u32 >> 9) * 2^-23) * scale + (-1), scale = 2/(1 - 2^-23)
// rounded to f32 (0x40000001); new_inclusive's shrink loop leaves it unchanged.
RDR_HD float range_pm1(uint32_t w)
{
    const float scale = u2f(0x40000001u);
    return fadd(fmul(fmul((float)(w >> 9), 1.1920928955078125e-07f), scale), -1.0f);
}
// utils/mod.rs:47-55: a uniform point of [-1,1]^3, normalised (not uniform on the sphere)
RDR_HD v3 random_in_unit_sphere(u4 w) { return normalize3(mk3(range_pm1(w.x), range_pm1(w.y), range_pm1(w.z))); }

// ---- camera ray: cpu.rs:199-202 (uv) and cpu.rs:234-251 -----------------------------------------
struct Camera {
    float inv_proj[16];   // column-major
    float inv_view[16];
    float pos[3];
    uint32_t width, height;
};

// Matrix4 * Vector4 (cgmath): ((c0*v0 + c1*v1) + c2*v2) + c3*v3
RDR_HD float mat_row(const float *m, int r, float v0, float v1, float v2, float v3_)
{
    return fadd(fadd(fadd(fmul(m[0 + r], v0), fmul(m[4 + r], v1)), fmul(m[8 + r], v2)), fmul(m[12 + r], v3_));
}

RDR_HD v3 camera_ray_dir(const Camera &cam, uint32_t x, uint32_t y)
{
    float u = fdiv((float)x, (float)cam.width);
    float v = fsub(1.0f, fdiv((float)y, (float)cam.height));
    float cx = fsub(fmul(u, 2.0f), 1.0f);
    float cy = fsub(fmul(v, 2.0f), 1.0f);
    float c0 = mat_row(cam.inv_proj, 0, cx, cy, -1.0f, -1.0f);
    float c1 = mat_row(cam.inv_proj, 1, cx, cy, -1.0f, -1.0f);
    float c2 = mat_row(cam.inv_proj, 2, cx, cy, -1.0f, -1.0f);
    float c3 = mat_row(cam.inv_proj, 3, cx, cy, -1.0f, -1.0f);
    float q0 = fdiv(c0, c3), q1 = fdiv(c1, c3), q2 = fdiv(c2, c3), q3 = fdiv(c3, c3);
    v3 w = mk3(mat_row(cam.inv_view, 0, q0, q1, q2, q3),
               mat_row(cam.inv_view, 1, q0, q1, q2, q3),
               mat_row(cam.inv_view, 2, q0, q1, q2, q3));
    return neg3(normalize3(w));
}

// ---- exact intersections: cpu.rs:34-62 and cpu.rs:64-98 -----------------------------------------
// returns true and *t on a hit; operation order is the reference's, one rounding per operation
RDR_HD bool hit_sphere_exact(v3 o, v3 d, v3 sc, float radius, float *t)
{
    float a = dot3(d, d);
    float k = fsub(dot3(o, d), dot3(d, sc));
    float c = fsub(fadd(fsub(dot3(o, o), fmul(2.0f, dot3(o, sc))), dot3(sc, sc)), fmul(radius, radius));
    float disc = fsub(fmul(k, k), fmul(a, c));
    if (disc < 0.0f) return false;
    float sq = fsqrt(disc);
    float t1 = fdiv(fsub(fneg(k), sq), a);
    float t2 = fdiv(fadd(fneg(k), sq), a);
    if (t1 >= 0.0f) { *t = t1; return true; }
    if (t2 >= 0.0f) { *t = t2; return true; }
    return false;
}

RDR_HD bool hit_cube_exact(v3 o, v3 d, v3 c, float side, float *t)
{
    float h = fmul(side, 0.5f);
    float t1x = fdiv(fsub(fsub(c.x, h), o.x), d.x), t2x = fdiv(fsub(fadd(c.x, h), o.x), d.x);
    float t1y = fdiv(fsub(fsub(c.y, h), o.y), d.y), t2y = fdiv(fsub(fadd(c.y, h), o.y), d.y);
    float t1z = fdiv(fsub(fsub(c.z, h), o.z), d.z), t2z = fdiv(fsub(fadd(c.z, h), o.z), d.z);
    float tmin = fmax(fmax(fmin(t1x, t2x), fmin(t1y, t2y)), fmin(t1z, t2z));
    float tmax = fmin(fmin(fmax(t1x, t2x), fmax(t1y, t2y)), fmax(t1z, t2z));
    if (tmax < 0.0f) return false;
    if (tmin > tmax) return false;
    *t = (tmin < 0.0f) ? tmax : tmin;
    return true;
}

// Winner selection of trace_ray (cpu.rs:344-352): min_by_key(OrderedFloat(t)) keeps the FIRST
// minimum in object order, NaN sorts last.  Expressed as a lexicographic (t, index) minimum so
// that candidates may be visited in any order.
RDR_HD bool hit_better(float t, int idx, float best_t, int best_idx)
{
    if (best_idx < 0) return true;
    if (isnan_(t)) return isnan_(best_t) && idx < best_idx;
    if (isnan_(best_t)) return true;
    return t < best_t || (t == best_t && idx < best_idx);
}

// ---- closest_hit: cpu.rs:354-394 -----------------------------------------------------------------
struct Surface { v3 p, n; bool front; };

RDR_HD float signum_(float x) { return isnan_(x) ? x : u2f((f2u(x) & 0x80000000u) | 0x3f800000u); }

RDR_HD Surface closest_hit(v3 o, v3 d, float t, bool is_sphere, v3 c, float size)
{
    Surface s;
    s.p = add3(o, scale3(d, t));
    v3 n;
    if (is_sphere) {
        n = normalize3(sub3(s.p, c));
    } else {
        v3 l = sub3(s.p, c);
        float half_side = fmul(size, 0.5f);           // == size / 2.0 (cpu.rs:372), exactly: both are the correctly rounded half
        float xd = fabs_(fsub(fabs_(l.x), half_side));
        float yd = fabs_(fsub(fabs_(l.y), half_side));
        float zd = fabs_(fsub(fabs_(l.z), half_side));
        if (xd < yd && xd < zd) n = mk3(signum_(l.x), 0.0f, 0.0f);
        else if (yd < zd)       n = mk3(0.0f, signum_(l.y), 0.0f);
        else                    n = mk3(0.0f, 0.0f, signum_(l.z));
    }
    s.front = dot3(n, d) <= 0.0f;
    if (!s.front) n = neg3(n);
    s.n = n;
    return s;
}

// ---- utils/mod.rs:14-44 ----------------------------------------------------------------------------
RDR_HD v3 reflect3(v3 v, v3 n) { return sub3(v, scale3(scale3(n, dot3(v, n)), 2.0f)); }

RDR_HD v3 refract3(v3 v, v3 n, float ratio)
{
    float cos_theta = fmin(dot3(v, neg3(n)), 1.0f);
    v3 perp = scale3(add3(v, scale3(n, cos_theta)), ratio);
    float s = fneg(fsqrt(fabs_(fsub(1.0f, dot3(perp, perp)))));
    return add3(perp, scale3(n, s));
}

RDR_HD bool can_refract3(v3 v, v3 n, float ratio)
{
    float cos_theta = fmin(dot3(v, neg3(n)), 1.0f);
    float sin_theta = fsqrt(fsub(1.0f, fmul(cos_theta, cos_theta)));
    return fmul(ratio, sin_theta) <= 1.0f;
}

// ---- World::sample, world.rs:17-34 -------------------------------------------------------------------
struct World { uint32_t kind; float a[3]; float b[3]; };

RDR_HD v3 world_sample(const World &w, v3 d)
{
    v3 a = mk3(w.a[0], w.a[1], w.a[2]);
    if (w.kind != 0u) return a;                       // SolidColor
    v3 up = mk3(0.0f, 1.0f, 0.0f);
    float cosine = fdiv(dot3(d, up), fmul(fsqrt(dot3(d, d)), fsqrt(dot3(up, up))));
    v3 bottom = mk3(w.b[0], w.b[1], w.b[2]);
    return add3(bottom, scale3(sub3(a, bottom), fmul(fadd(cosine, 1.0f), 0.5f)));
}

// ---- material & scatter: cpu.rs:262-333 -----------------------------------------------------------
struct Material {
    v3 albedo; float roughness; float metallic; v3 emission; float emission_strength; float transmission; float ior;
};

struct Scatter { v3 origin, dir; uint32_t lobe; };

// One bounce off `s`.  RNG slots are the spec of oracle/raydar_oracle.h: block 0 = (u1,u2,u3),
// block 1/2/3 = the diffuse / specular / refraction random vectors.  Only the vector of the
// lobe that is actually taken is generated (counter-based RNG: unused draws cost nothing).
RDR_HD Scatter scatter(v3 rd, const Surface &s, const Material &m,
                       uint32_t seed_lo, uint32_t seed_hi, uint32_t pixel, uint32_t sample, uint32_t bounce)
{
    const float roughness = fmul(m.roughness, m.roughness);
    const u4 b0 = rng_block(seed_lo, seed_hi, pixel, sample, bounce, 0u);
    const bool transmission_ray = u01(b0.x) < m.transmission;
    const float u2 = u01(b0.y);

    uint32_t lobe;
    float ior = m.ior;
    v3 rdn = rd;
    if (transmission_ray) {
        if (s.front) ior = frcp(ior);
        rdn = normalize3(rd);
        float cos_theta = fmin(dot3(rdn, neg3(s.n)), 1.0f);
        float q = fdiv(fsub(ior, 1.0f), fadd(ior, 1.0f));
        float r0 = fmul(q, q);
        float w = fsub(1.0f, cos_theta);
        float w2 = fmul(w, w);
        float w5 = fmul(w, fmul(w2, w2));
        float refl = fadd(r0, fmul(fsub(1.0f, r0), w5));
        lobe = (refl < u2 && can_refract3(rdn, s.n, ior)) ? 3u : 2u;
    } else if (u2 < m.metallic) {
        lobe = 2u;
    } else {
        lobe = (u01(b0.z) < roughness) ? 1u : 2u;
    }

    const v3 r = random_in_unit_sphere(rng_block(seed_lo, seed_hi, pixel, sample, bounce, lobe));
    v3 dir;
    if (lobe == 1u) {
        dir = add3(s.n, r);
        if (dot3(dir, s.n) < 0.0f) dir = neg3(dir);
    } else {
        // specular and refracted lobes share normalize(base + r * roughness): one copy of that code for both
        // (lanes of a warp that took different lobes run it together)
        const v3 base = (lobe == 2u) ? reflect3(rd, s.n) : refract3(rdn, s.n, ior);
        dir = normalize3(add3(base, scale3(r, roughness)));
    }

    Scatter out;
    const v3 offset = transmission_ray ? dir : s.n;
    out.origin = add3(s.p, scale3(offset, 0.0001f));
    out.dir = (dot3(dir, dir) < 1e-10f) ? s.n : dir;
    out.lobe = lobe;
    return out;
}

// ---- print_frame_buffer, cpu.rs:221-230: ((sum / n).clamp(0,1) * 255) as u8 ---------------------
RDR_HD uint32_t quantise(float sum, float n)
{
    float v = fdiv(sum, n);
    if (v < 0.0f) v = 0.0f; else if (v > 1.0f) v = 1.0f;   // f32::clamp keeps NaN
    v = fmul(v, 255.0f);
    if (isnan_(v)) return 0u;                                // `as u8`: NaN -> 0, saturating, truncating
    if (v <= 0.0f) return 0u;
    if (v >= 255.0f) return 255u;
    return (uint32_t)(int)v;
}

// =====================================================================================================
// Conservative cull tests.
//
// trace_ray visits every object (cpu.rs:344-352).  The winner is the lexicographic (t, index)
// minimum over the objects whose exact test hits, so an object may be skipped whenever it
// provably cannot hit or provably cannot beat the current best.  The tests below bound the
// AS-WRITTEN f32 result of hit_sphere / hit_cube (including its rounding noise, which for the
// sphere's expanded quadratic is large) from cheap FMA arithmetic plus explicit margins.  Every
// reject condition is a comparison that is false on NaN, so a NaN keeps the object.
//
// Notation: u = 2^-24.
//   sphere, as written: k = o.d - d.c, c = o.o - 2 o.c + c.c - r^2, disc = k^2 - a c.
//     |k - d.(o-c)|        <=  4u |d| (|o|+|c|)
//     |c - (|o-c|^2-r^2)|  <=  6u ((|o|+|c|)^2 + r^2)
//     |disc - true|        <= 20u a S,  S = (|o|+|c|)^2 + r^2 <= 2|o|^2 + (2|c|^2 + r^2)
//   the FMA forms below add <= 8u a S.  Margins used: M = 2^-19 a S_ray (= 32u a S_ray) on
//   disc, Ms = 2^-19 S_ray on c, ek = 2^-19 sqrt(a S_ray) on k, with
//   S_ray = 2|o|^2 + max_spheres(2|c|^2 + r^2).
//   cube, as written: t = ((c -+ h) - o) / d per axis, three roundings: |dt| <= 3u B' / |d_i|,
//     B' = |c_i| + h + |o_i|; the FMA form adds <= 5u B' / |d_i|.  Cubes are padded by
//     pad = 2^-18 B (= 64u B) with B >= B' for every origin inside the scene bound; rays whose
//     origin leaves the bound, or with a zero/denormal direction component, take the exact test on
//     everything (RayCull::degenerate).
// tests/test_cull_conservative.py hammers these bounds with adversarial random rays on the CPU.
// =====================================================================================================
struct CullConsts {
    float sphere_q_max;    // max over spheres of 2|c|^2 + r^2
    float origin_bound;    // rays with |o|_inf above this skip the cull
    float sphere_r_min;    // smallest sphere radius (BVH: per-ray inflation of sphere boxes)
};

struct RayCull {
    v3 inv, od, ainv;      // 1/d, o/d, |1/d|
    float a;               // d.d
    float M, Ms, ek;       // sphere margins (see above)
    float s_ray;           // 2|o|^2 + q_max
    bool degenerate;
};

// the sphere part of the set-up (no divisions): a, margins, and whether the ray must skip the cull
RDR_HD void sphere_margins(v3 o, v3 d, const CullConsts &cc, RayCull &rc)
{
    rc.a = fma(d.x, d.x, fma(d.y, d.y, fmul(d.z, d.z)));
    float oo = fma(o.x, o.x, fma(o.y, o.y, fmul(o.z, o.z)));
    float s_ray = fma(2.0f, oo, cc.sphere_q_max);
    rc.s_ray = s_ray;
    rc.Ms = fmul(1.9073486328125e-06f, s_ray);                    // 2^-19 S (32u; proven need 28u, observed 5.3u)
    rc.M = fmul(rc.a, rc.Ms);
    rc.ek = fmul(1.9073486328125e-06f, fsqrt(fmul(rc.a, s_ray))); // 2^-19 sqrt(a S)
    float omax = fmax(fmax(fabs_(o.x), fabs_(o.y)), fabs_(o.z));
    // all comparisons written so that NaN/inf anywhere => degenerate
    rc.degenerate = !(omax <= cc.origin_bound) || !(rc.a > 1e-30f) || !(rc.a < 1e30f) || !(s_ray < 1e30f);
}

RDR_HD RayCull make_ray_cull(v3 o, v3 d, const CullConsts &cc)
{
    RayCull rc;
    rc.inv = mk3(frcp(d.x), frcp(d.y), frcp(d.z));
    rc.od = mk3(fmul(o.x, rc.inv.x), fmul(o.y, rc.inv.y), fmul(o.z, rc.inv.z));
    rc.ainv = mk3(fabs_(rc.inv.x), fabs_(rc.inv.y), fabs_(rc.inv.z));
    sphere_margins(o, d, cc, rc);
    float imax = fmax(fmax(rc.ainv.x, rc.ainv.y), rc.ainv.z);
    rc.degenerate = rc.degenerate || !(imax < 1e30f);
    return rc;
}

// sphere cull datum: (cx, cy, cz, r^2).  true = may hit (run the exact test)
RDR_HD bool sphere_may_hit(v3 o, v3 d, const RayCull &rc, float cx, float cy, float cz, float r2)
{
    float lx = fsub(cx, o.x), ly = fsub(cy, o.y), lz = fsub(cz, o.z);
    float bp = fma(lx, d.x, fma(ly, d.y, fmul(lz, d.z)));          // d.(c-o) = -k
    float cc = fma(lx, lx, fma(ly, ly, fma(lz, lz, fneg(r2))));    // |c-o|^2 - r^2
    float D = fma(fneg(rc.a), cc, fma(bp, bp, rc.M));              // disc + margin
    bool reject = (D < 0.0f) || ((bp < fneg(rc.ek)) && (cc > rc.Ms));
    return !reject;
}

// cube cull datum: (cx, cy, cz, h + pad).  best = current best exact t (+inf if none):
// a cube whose padded entry distance exceeds it cannot win (ties need tn == best: kept).
// One comparison decides: reject when max(tn, 0) > min(tf, best), which covers "behind the origin"
// (tf < 0 <= max(tn, 0)), "missed" (tn > tf) and "cannot win" (tn > best >= 0).
RDR_HD bool cube_may_hit(const RayCull &rc, float cx, float cy, float cz, float hp, float best)
{
    float tcx = fma(cx, rc.inv.x, fneg(rc.od.x));
    float tcy = fma(cy, rc.inv.y, fneg(rc.od.y));
    float tcz = fma(cz, rc.inv.z, fneg(rc.od.z));
    float tn = fmax(fmax(fma(fneg(hp), rc.ainv.x, tcx), fma(fneg(hp), rc.ainv.y, tcy)), fmax(fma(fneg(hp), rc.ainv.z, tcz), 0.0f));
    float tf = fmin(fmin(fma(hp, rc.ainv.x, tcx), fma(hp, rc.ainv.y, tcy)), fmin(fma(hp, rc.ainv.z, tcz), best));
    return !(tn > tf);
}

// ---- BVH traversal helpers -------------------------------------------------------------------------------
// Ray set-up for box tests against the 8-wide hierarchy (rdr_bvh.h).  Unlike the scan there is no O(N)
// fallback worth taking for axis-parallel rays, so a zero / denormal direction component is handled by
// clamping |1/d| to 1e30: the as-written slab of such an axis is (-inf, +inf) when the origin is inside it
// and empty otherwise; the clamped slab is (-(e -+ delta) 1e30, +(e +- delta) 1e30), i.e. unbounded for every
// hit distance below ~1e20 because the boxes are padded by >= 1e-6 (documented limit of the BVH path).
// rho: per-ray inflation of sphere-flagged boxes, covering (a) the cancellation noise of the reference's
// expanded sphere quadratic, r_eff^2 <= r^2 + Ms with Ms = 2^-19 (2|o|^2 + q_max), so
// r_eff - r <= sqrt(r_min^2 + Ms) - r_min, and (b) the error of its root, <= 2^-19 sqrt(S_ray) in distance.
struct RayBvh {
    RayCull rc;
    float rho;
};

RDR_HD RayBvh make_ray_bvh(v3 o, v3 d, const CullConsts &cc)
{
    RayBvh rb;
    rb.rc = make_ray_cull(o, d, cc);
    const float big = 1e30f;
    bool clamped = false;
    if (!(rb.rc.ainv.x < big)) { rb.rc.inv.x = u2f((f2u(d.x) & 0x80000000u) | f2u(big)); clamped = true; }
    if (!(rb.rc.ainv.y < big)) { rb.rc.inv.y = u2f((f2u(d.y) & 0x80000000u) | f2u(big)); clamped = true; }
    if (!(rb.rc.ainv.z < big)) { rb.rc.inv.z = u2f((f2u(d.z) & 0x80000000u) | f2u(big)); clamped = true; }
    if (clamped) {
        rb.rc.od = mk3(fmul(o.x, rb.rc.inv.x), fmul(o.y, rb.rc.inv.y), fmul(o.z, rb.rc.inv.z));
        rb.rc.ainv = mk3(fabs_(rb.rc.inv.x), fabs_(rb.rc.inv.y), fabs_(rb.rc.inv.z));
        const float omax = fmax(fmax(fabs_(o.x), fabs_(o.y)), fabs_(o.z));
        // same conditions as make_ray_cull minus the 1/d test; a NaN direction component stays degenerate
        rb.rc.degenerate = !(omax <= cc.origin_bound) || !(rb.rc.a > 1e-30f) || !(rb.rc.a < 1e30f) || !(rb.rc.s_ray < 1e30f) ||
                           isnan_(d.x) || isnan_(d.y) || isnan_(d.z);
    }
    rb.rho = fadd(fsub(fsqrt(fma(cc.sphere_r_min, cc.sphere_r_min, rb.rc.Ms)), cc.sphere_r_min),
                  fmul(1.9073486328125e-06f, fsqrt(rb.rc.s_ray)));
    rb.rho = fmul(rb.rho, 1.0001f);
    return rb;
}

// one hierarchy entry: (cx, cy, cz, ex) (ey, ez, payload, sphere flag).  Returns true when the padded box may
// contain a hit that can still win (entry distance not beyond `best`); *tn_out = its conservative entry distance.
RDR_HD bool bvh_entry_may_hit(const RayBvh &rb, f4 q0, f4 q1, float best, float *tn_out)
{
    const RayCull &rc = rb.rc;
    const float ex = fma(q1.w, rb.rho, q0.w), ey = fma(q1.w, rb.rho, q1.x), ez = fma(q1.w, rb.rho, q1.y);
    const float tcx = fma(q0.x, rc.inv.x, fneg(rc.od.x));
    const float tcy = fma(q0.y, rc.inv.y, fneg(rc.od.y));
    const float tcz = fma(q0.z, rc.inv.z, fneg(rc.od.z));
    const float tn = fmax(fmax(fma(fneg(ex), rc.ainv.x, tcx), fma(fneg(ey), rc.ainv.y, tcy)), fmax(fma(fneg(ez), rc.ainv.z, tcz), 0.0f));
    const float tf = fmin(fmin(fma(ex, rc.ainv.x, tcx), fma(ey, rc.ainv.y, tcy)), fmin(fma(ez, rc.ainv.z, tcz), best));
    *tn_out = tn;
    return !(tn > tf);
}

}  // namespace rdr

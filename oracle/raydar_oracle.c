/*
 * raydar_oracle.c -- CPU restatement of bvpav/raydar's CPU backend.  TEST INFRASTRUCTURE ONLY;
 * see raydar_oracle.h for scope, citations and the "parity unpinned" statement.
 *
 * Every arithmetic expression below keeps the operation order of the Rust source (and of the
 * cgmath 0.18 / rand 0.8.5 generic code it instantiates).  Compile with -ffp-contract=off.
 */
#include "raydar_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(__FMA__) && !defined(ORC_ALLOW_FMA_ISA)
/* -mfma alone does not contract with -ffp-contract=off, but keep the build honest. */
#endif

typedef struct { float x, y, z; } v3;

/* ---- cgmath 0.18 restatements ------------------------------------------------------------- */
static inline v3 V(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmuls(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vdiv(v3 a, v3 b) { return V(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
/* InnerSpace::dot = mul_element_wise(..).sum(), sum = (x + y) + z */
static inline float vdot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float vmag2(v3 a) { return vdot(a, a); }
static inline float vmag(v3 a) { return sqrtf(vmag2(a)); }
/* InnerSpace::normalize = normalize_to(1) = self * (1 / magnitude) */
static inline v3 vnormalize(v3 a) { return vmuls(a, 1.0f / vmag(a)); }
static inline v3 ld3(const float *p) { return V(p[0], p[1], p[2]); }
static inline void st3(float *p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }

/* Rust f32::min / f32::max: a NaN operand is ignored.  Signed zeros are unspecified in Rust;
 * we fix -0 < +0 (IEEE 754-2019 minimumNumber, which is what PTX min.f32/max.f32 do). */
static inline float rmin(float a, float b)
{
    if (a != a) return b;
    if (b != b) return a;
    if (a == 0.0f && b == 0.0f) return signbit(a) ? a : b;
    return a < b ? a : b;
}
static inline float rmax(float a, float b)
{
    if (a != a) return b;
    if (b != b) return a;
    if (a == 0.0f && b == 0.0f) return signbit(a) ? b : a;
    return a > b ? a : b;
}
/* Rust f32::signum: NaN -> NaN, otherwise copysign(1, x) (so +0 -> 1, -0 -> -1) */
static inline float rsignum(float x) { return (x != x) ? x : copysignf(1.0f, x); }

/* Matrix4<f32> * Vector4<f32>: ((c0*v0 + c1*v1) + c2*v2) + c3*v3, m column-major */
static inline void mat4_mul_vec4(const float m[16], const float v[4], float out[4])
{
    for (int r = 0; r < 4; ++r)
        out[r] = ((m[0 + r] * v[0] + m[4 + r] * v[1]) + m[8 + r] * v[2]) + m[12 + r] * v[3];
}

/* ---- RNG spec: Philox4x32-10 (Salmon et al., SC'11) ----------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_rng_block(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce, uint32_t block, uint32_t out[4])
{
    uint32_t ctr[4] = { pixel, sample, bounce * 4u + block, 0u };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    orc_philox4x32_10(ctr, key, out);
}

/* rand 0.8.5 Standard for f32: 24 random mantissa bits, [0,1) */
float orc_u01(uint32_t word) { return (float)(word >> 8) * (1.0f / 16777216.0f); }

/* rand 0.8.5 UniformFloat<f32>::new_inclusive(-1, 1).sample():
 *   max_rand = 1 - 2^-23; scale = (high - low) / max_rand, decreased while scale*max_rand + low > high;
 *   value0_1 = bits(u32 >> 9 | exponent 0) - 1.0;  result = value0_1 * scale + low */
static float range_pm1_scale(void)
{
    const float low = -1.0f, high = 1.0f;
    const float max_rand = 1.0f - 1.1920929e-7f;
    float scale = (high - low) / max_rand;
    for (;;) {
        volatile float probe = scale * max_rand + low;
        if (!(probe > high)) break;
        uint32_t bits; memcpy(&bits, &scale, 4); bits -= 1; memcpy(&scale, &bits, 4);
    }
    return scale;
}
float orc_range_pm1(uint32_t word)
{
    static float cached_scale = 0.0f;                  /* same value from every thread: benign race */
    if (cached_scale == 0.0f) cached_scale = range_pm1_scale();
    const float scale = cached_scale;
    float value0_1 = (float)(word >> 9) * (1.0f / 8388608.0f);
    return value0_1 * scale + -1.0f;
}

/* utils/mod.rs:47-55 */
void orc_random_in_unit_sphere(const uint32_t words[3], float out[3])
{
    v3 p = V(orc_range_pm1(words[0]), orc_range_pm1(words[1]), orc_range_pm1(words[2]));
    st3(out, vnormalize(p));
}

/* ---- camera ray: cpu.rs:199-202 (uv) + cpu.rs:234-251 ---------------------------------------- */
void orc_camera_ray(const OrcScene *s, uint32_t x, uint32_t y, float origin[3], float dir[3])
{
    float u = (float)x / (float)s->width;
    float v = 1.0f - (float)y / (float)s->height;
    float clip[4] = { u * 2.0f - 1.0f, v * 2.0f - 1.0f, -1.0f, -1.0f };
    float cs[4], ws[4];
    mat4_mul_vec4(s->inv_proj, clip, cs);
    float w = cs[3];
    cs[0] = cs[0] / w; cs[1] = cs[1] / w; cs[2] = cs[2] / w; cs[3] = cs[3] / w;
    mat4_mul_vec4(s->inv_view, cs, ws);
    v3 d = vneg(vnormalize(V(ws[0], ws[1], ws[2])));
    origin[0] = s->cam_pos[0]; origin[1] = s->cam_pos[1]; origin[2] = s->cam_pos[2];
    st3(dir, d);
}

/* ---- intersections: cpu.rs:34-98 ------------------------------------------------------------- */
int orc_hit_sphere(const float o_[3], const float d_[3], const float center[3], float radius, float *t)
{
    v3 o = ld3(o_), d = ld3(d_), sc = ld3(center);
    float a = vdot(d, d);
    float k = vdot(o, d) - vdot(d, sc);
    float c = vdot(o, o) - 2.0f * vdot(o, sc) + vdot(sc, sc) - radius * radius;
    float disc = k * k - a * c;
    if (disc < 0.0f) return 0;
    float sq = sqrtf(disc);
    float t1 = (-k - sq) / a;
    float t2 = (-k + sq) / a;
    if (t1 >= 0.0f) { *t = t1; return 1; }
    if (t2 >= 0.0f) { *t = t2; return 1; }
    return 0;
}

int orc_hit_cube(const float o_[3], const float d_[3], const float center[3], float side, float *t)
{
    v3 o = ld3(o_), d = ld3(d_), c = ld3(center);
    v3 half = V(side * 0.5f, side * 0.5f, side * 0.5f);
    v3 mn = vsub(c, half);
    v3 mx = vadd(c, half);
    v3 t1 = vdiv(vsub(mn, o), d);
    v3 t2 = vdiv(vsub(mx, o), d);
    float tmin = rmax(rmax(rmin(t1.x, t2.x), rmin(t1.y, t2.y)), rmin(t1.z, t2.z));
    float tmax = rmin(rmin(rmax(t1.x, t2.x), rmax(t1.y, t2.y)), rmax(t1.z, t2.z));
    if (tmax < 0.0f) return 0;
    if (tmin > tmax) return 0;
    *t = (tmin < 0.0f) ? tmax : tmin;
    return 1;
}

static inline int hit_object(const OrcScene *s, uint32_t i, const float o[3], const float d[3], float *t)
{
    const float *g = s->geom + 4 * (size_t)i;
    return s->kind[i] == ORC_SPHERE ? orc_hit_sphere(o, d, g, g[3], t) : orc_hit_cube(o, d, g, g[3], t);
}

/* OrderedFloat total order: NaN is greater than everything and equal to itself */
static inline int ordered_less(float a, float b)
{
    if (a != a) return 0;
    if (b != b) return 1;
    return a < b;
}

/* cpu.rs:344-352: filter_map over ALL objects, min_by_key keeps the first minimum */
int orc_trace(const OrcScene *s, const float o[3], const float d[3], float *t_out)
{
    int best = -1;
    float best_t = 0.0f;
    for (uint32_t i = 0; i < s->n_objects; ++i) {
        float t;
        if (!hit_object(s, i, o, d, &t)) continue;
        if (best < 0 || ordered_less(t, best_t)) { best = (int)i; best_t = t; }
    }
    if (best >= 0) *t_out = best_t;
    return best;
}


/* ---- oracle-side BVH (NOT in the reference: cpu.rs:344-352 is a linear scan) ---------------------------------------
 * A binary median-split hierarchy over the objects' boxes that returns exactly orc_trace's winner: the boxes only
 * decide which objects get the exact test above; the winner is the (t, index) minimum with NaN last, and a subtree is
 * skipped only when its (conservatively early) entry distance lies strictly beyond the best exact t so far, so equal-t
 * candidates with a lower index are still visited.  Conservative: the slab test runs in double precision on boxes
 * grown per ray by bounds of the exact tests' own f32 noise -- the as-written sphere quadratic cancels badly, so a
 * sphere can report a hit up to sqrt(r^2 + eps S) from its centre, S = (|o| + |c|)^2 + r^2 (margins are 4x the bound
 * DESIGN.md 4.2 derives).  Rays with a zero, denormal or non-finite component take the linear scan.
 * Used to time a CPU path at 100k objects (bench.py) and checked against orc_trace in tests/test_oracle_bvh.py. */
typedef struct { float lo[3], hi[3]; int32_t left, right; uint32_t first, count; uint32_t has_sphere; } OrcBvhNode;
struct OrcBvh {
    OrcBvhNode *nodes; uint32_t n_nodes;
    uint32_t *order;              /* leaf object indices */
    double c_max;                 /* max |centre| + extent over the scene */
    double r_min;                 /* smallest sphere radius */
    double q_max;                 /* max over spheres of 2 |c|^2 + r^2 */
};

static void obj_box(const OrcScene *s, uint32_t i, float lo[3], float hi[3])
{
    const float *g = s->geom + 4 * (size_t)i;
    const float e = s->kind[i] == ORC_SPHERE ? fabsf(g[3]) : fabsf(g[3]) * 0.5f;
    for (int a = 0; a < 3; ++a) { lo[a] = g[a] - e; hi[a] = g[a] + e; }
}

static int32_t bvh_build_rec(const OrcScene *s, OrcBvh *b, uint32_t first, uint32_t count)
{
    const int32_t me = (int32_t)b->n_nodes++;
    OrcBvhNode *n = &b->nodes[me];
    n->first = first; n->count = count; n->left = n->right = -1; n->has_sphere = 0;
    float clo[3] = { INFINITY, INFINITY, INFINITY }, chi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (int a = 0; a < 3; ++a) { n->lo[a] = INFINITY; n->hi[a] = -INFINITY; }
    for (uint32_t k = 0; k < count; ++k) {
        const uint32_t i = b->order[first + k];
        float lo[3], hi[3];
        obj_box(s, i, lo, hi);
        if (s->kind[i] == ORC_SPHERE) n->has_sphere = 1;
        for (int a = 0; a < 3; ++a) {
            if (lo[a] < n->lo[a]) n->lo[a] = lo[a];
            if (hi[a] > n->hi[a]) n->hi[a] = hi[a];
            const float c = s->geom[4 * (size_t)i + a];
            if (c < clo[a]) clo[a] = c;
            if (c > chi[a]) chi[a] = c;
        }
    }
    if (count <= 4) return me;
    int axis = 0;
    if (chi[1] - clo[1] > chi[axis] - clo[axis]) axis = 1;
    if (chi[2] - clo[2] > chi[axis] - clo[axis]) axis = 2;
    if (!(chi[axis] > clo[axis])) return me;                       /* coincident centres: one leaf */
    /* median split by centre: nth_element by simple quickselect */
    uint32_t *o = b->order + first;
    uint32_t lo_i = 0, hi_i = count - 1, mid = count / 2;
    while (lo_i < hi_i) {
        const float pivot = s->geom[4 * (size_t)o[(lo_i + hi_i) / 2] + axis];
        uint32_t i = lo_i, j = hi_i;
        while (i <= j) {
            while (s->geom[4 * (size_t)o[i] + axis] < pivot) ++i;
            while (s->geom[4 * (size_t)o[j] + axis] > pivot) --j;
            if (i <= j) { const uint32_t t = o[i]; o[i] = o[j]; o[j] = t; ++i; if (j == 0) break; --j; }
        }
        if (mid <= j) hi_i = j; else if (mid >= i) lo_i = i; else break;
    }
    const int32_t l = bvh_build_rec(s, b, first, mid);
    const int32_t r = bvh_build_rec(s, b, first + mid, count - mid);
    n = &b->nodes[me];                                              /* (the array does not move: sized up front) */
    n->left = l; n->right = r;
    return me;
}

OrcBvh *orc_bvh_build(const OrcScene *s)
{
    OrcBvh *b = (OrcBvh *)calloc(1, sizeof *b);
    const uint32_t n = s->n_objects;
    b->nodes = (OrcBvhNode *)calloc(2 * (size_t)n + 1, sizeof *b->nodes);
    b->order = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    b->r_min = INFINITY; b->q_max = 0.0; b->c_max = 0.0;
    for (uint32_t i = 0; i < n; ++i) {
        b->order[i] = i;
        const float *g = s->geom + 4 * (size_t)i;
        const double c2 = (double)g[0] * g[0] + (double)g[1] * g[1] + (double)g[2] * g[2];
        const double ext = fabs((double)g[3]);
        if (sqrt(c2) + ext > b->c_max) b->c_max = sqrt(c2) + ext;
        if (s->kind[i] == ORC_SPHERE) {
            if (ext < b->r_min) b->r_min = ext;
            if (2.0 * c2 + ext * ext > b->q_max) b->q_max = 2.0 * c2 + ext * ext;
        }
    }
    if (!(b->r_min < INFINITY)) b->r_min = 0.0;
    if (n) bvh_build_rec(s, b, 0, n);
    return b;
}

void orc_bvh_free(OrcBvh *b)
{
    if (!b) return;
    free(b->nodes); free(b->order); free(b);
}

/* (t, i) strictly better than the best so far under trace_ray's rule: smaller t, NaN last, first index on ties */
static inline int hit_better(float t, int i, float bt, int bi)
{
    if (bi < 0) return 1;
    if (ordered_less(t, bt)) return 1;
    if (ordered_less(bt, t)) return 0;
    return i < bi;
}

int orc_trace_bvh(const OrcScene *s, const OrcBvh *b, const float o[3], const float d[3], float *t_out, uint64_t *tests)
{
    int ok = 1;
    for (int a = 0; a < 3; ++a) if (!(fabsf(d[a]) > 1e-30f && fabsf(d[a]) < 1e30f) || !(fabsf(o[a]) < 1e30f)) ok = 0;
    if (!ok || b->n_nodes == 0) { if (tests) *tests = s->n_objects; return orc_trace(s, o, d, t_out); }
    const double oo = (double)o[0] * o[0] + (double)o[1] * o[1] + (double)o[2] * o[2];
    const double s_ray = 2.0 * oo + b->q_max;
    const double eps = 1.0 / 131072.0;                                   /* 2^-17 */
    const double rho = sqrt(b->r_min * b->r_min + eps * s_ray) - b->r_min + eps * sqrt(s_ray);
    const double pad = eps * (b->c_max + sqrt(oo)) + 1e-30;
    double inv[3], nod[3];
    for (int a = 0; a < 3; ++a) { inv[a] = 1.0 / (double)d[a]; nod[a] = -(double)o[a] * inv[a]; }
    int best = -1; float best_t = 0.0f;
    uint64_t n_tests = 0;
    int32_t stack[128]; int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
        const OrcBvhNode *n = &b->nodes[stack[--sp]];
        const double grow = pad + (n->has_sphere ? rho : 0.0);
        double tn = 0.0, tf = INFINITY;
        for (int a = 0; a < 3; ++a) {
            double t1 = ((double)n->lo[a] - grow) * inv[a] + nod[a], t2 = ((double)n->hi[a] + grow) * inv[a] + nod[a];
            if (t1 > t2) { const double t = t1; t1 = t2; t2 = t; }
            if (t1 > tn) tn = t1;
            if (t2 < tf) tf = t2;
        }
        tn *= (1.0 - 1e-9); tf *= (1.0 + 1e-9);
        if (tn > tf) continue;
        if (best >= 0 && best_t == best_t && tn > (double)best_t) continue;      /* strictly beyond the best exact t: ties are visited */
        if (n->left < 0) {
            for (uint32_t k = 0; k < n->count; ++k) {
                const uint32_t i = b->order[n->first + k];
                float t;
                ++n_tests;
                if (!hit_object(s, i, o, d, &t)) continue;
                if (hit_better(t, (int)i, best_t, best)) { best = (int)i; best_t = t; }
            }
        } else if (sp + 2 <= 128) {
            stack[sp++] = n->left; stack[sp++] = n->right;
        } else {                                                          /* cannot happen for a median-split tree of < 2^60 objects */
            if (tests) *tests = s->n_objects;
            return orc_trace(s, o, d, t_out);
        }
    }
    if (tests) *tests = n_tests;
    if (best >= 0) *t_out = best_t;
    return best;
}

/* cpu.rs:354-394 */
void orc_closest_hit(const OrcScene *s, int obj, const float o_[3], const float d_[3], float t,
                     float pos[3], float normal[3], uint32_t *front_face)
{
    v3 o = ld3(o_), d = ld3(d_);
    const float *g = s->geom + 4 * (size_t)obj;
    v3 c = ld3(g);
    v3 p = vadd(o, vmuls(d, t));                       /* Ray::at, cpu.rs:30-32 */
    v3 n;
    if (s->kind[obj] == ORC_SPHERE) {
        n = vnormalize(vsub(p, c));
    } else {
        v3 l = vsub(p, c);
        float half_side = g[3] / 2.0f;
        float xd = fabsf(fabsf(l.x) - half_side);
        float yd = fabsf(fabsf(l.y) - half_side);
        float zd = fabsf(fabsf(l.z) - half_side);
        if (xd < yd && xd < zd) n = V(rsignum(l.x), 0.0f, 0.0f);
        else if (yd < zd)       n = V(0.0f, rsignum(l.y), 0.0f);
        else                    n = V(0.0f, 0.0f, rsignum(l.z));
    }
    int front = vdot(n, d) <= 0.0f;
    if (!front) n = vneg(n);
    st3(pos, p); st3(normal, n); *front_face = (uint32_t)front;
}

/* utils/mod.rs:14-16: v - n * dot(v,n) * 2 */
static inline v3 reflect3(v3 v, v3 n) { return vsub(v, vmuls(vmuls(n, vdot(v, n)), 2.0f)); }
void orc_reflect(const float v[3], const float n[3], float out[3]) { st3(out, reflect3(ld3(v), ld3(n))); }

/* utils/mod.rs:25-35 (the unit-length asserts are not reproduced: the oracle never traps) */
static inline v3 refract3(v3 v, v3 n, float ratio)
{
    float cos_theta = rmin(vdot(v, vneg(n)), 1.0f);
    v3 perp = vmuls(vadd(v, vmuls(n, cos_theta)), ratio);
    float s = -sqrtf(fabsf(1.0f - vmag2(perp)));
    v3 par = vmuls(n, s);
    return vadd(perp, par);
}
void orc_refract(const float v[3], const float n[3], float ratio, float out[3]) { st3(out, refract3(ld3(v), ld3(n), ratio)); }

/* utils/mod.rs:37-44 */
static inline int can_refract3(v3 v, v3 n, float ratio)
{
    float cos_theta = rmin(vdot(v, vneg(n)), 1.0f);
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    return ratio * sin_theta <= 1.0f;
}
int orc_can_refract(const float v[3], const float n[3], float ratio) { return can_refract3(ld3(v), ld3(n), ratio); }

/* world.rs:17-34 */
static inline v3 world_sample3(const OrcScene *s, v3 d)
{
    if (s->world_kind == ORC_WORLD_SKY) {
        v3 up = V(0.0f, 1.0f, 0.0f);
        float cosine = vdot(d, up) / (vmag(d) * vmag(up));
        v3 top = ld3(s->world_a), bottom = ld3(s->world_b);
        /* VectorSpace::lerp: self + (other - self) * amount */
        return vadd(bottom, vmuls(vsub(top, bottom), (cosine + 1.0f) * 0.5f));
    }
    /* SolidColor; Transparent is todo!() in the reference (callers reject it before rendering) */
    return ld3(s->world_a);
}
void orc_world_sample(const OrcScene *s, const float d[3], float out[3]) { st3(out, world_sample3(s, ld3(d))); }

void orc_hit_sphere_batch(uint32_t n, const float *rays, const float *spheres, float *out_t, int32_t *out_hit)
{
    for (uint32_t i = 0; i < n; ++i) {
        float t = 0.0f;
        out_hit[i] = orc_hit_sphere(rays + 6 * (size_t)i, rays + 6 * (size_t)i + 3, spheres + 4 * (size_t)i, spheres[4 * (size_t)i + 3], &t);
        out_t[i] = out_hit[i] ? t : 0.0f;
    }
}
void orc_hit_cube_batch(uint32_t n, const float *rays, const float *cubes, float *out_t, int32_t *out_hit)
{
    for (uint32_t i = 0; i < n; ++i) {
        float t = 0.0f;
        out_hit[i] = orc_hit_cube(rays + 6 * (size_t)i, rays + 6 * (size_t)i + 3, cubes + 4 * (size_t)i, cubes[4 * (size_t)i + 3], &t);
        out_t[i] = out_hit[i] ? t : 0.0f;
    }
}

void orc_first_hit(const OrcScene *s, int32_t *ids, float *ts, int n_threads)
{
    (void)n_threads;
    const int64_t H = s->height, W = s->width;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads > 0 ? n_threads : 1)
#endif
    for (int64_t y = 0; y < H; ++y) {
        for (int64_t x = 0; x < W; ++x) {
            float o[3], d[3], t = 0.0f;
            orc_camera_ray(s, (uint32_t)x, (uint32_t)y, o, d);
            int id = orc_trace(s, o, d, &t);
            ids[y * W + x] = id;
            ts[y * W + x] = id >= 0 ? t : 0.0f;
        }
    }
}

/* ---- per_pixel: cpu.rs:233-342 ---------------------------------------------------------------- */
static void per_pixel(const OrcScene *s, const OrcBvh *bvh, uint32_t x, uint32_t y, uint32_t sample, uint64_t seed,
                      uint32_t max_bounces, OrcPathStep *steps, uint32_t *n_steps, float rgba[4], OrcStats *st)
{
    const uint32_t pixel = y * s->width + x;
    float of[3], df[3];
    orc_camera_ray(s, x, y, of, df);
    v3 ro = ld3(of), rd = ld3(df);

    v3 light = V(0.0f, 0.0f, 0.0f);
    v3 atten = V(1.0f, 1.0f, 1.0f);
    uint32_t written = 0;
    uint32_t bounce;
    if (st) st->samples++;

    for (bounce = 0; bounce < max_bounces; ++bounce) {
        float o3[3], d3[3], t = 0.0f;
        st3(o3, ro); st3(d3, rd);
        if (st) { st->trace_calls++; if (bounce < 64) st->alive_at_bounce[bounce]++; }
        uint64_t tests = s->n_objects;
        int obj = bvh ? orc_trace_bvh(s, bvh, o3, d3, &t, &tests) : orc_trace(s, o3, d3, &t);
        if (st) st->primitive_tests += tests;
        if (obj >= 0) {
            float pf[3], nf[3]; uint32_t front;
            orc_closest_hit(s, obj, o3, d3, t, pf, nf, &front);
            v3 P = ld3(pf), n = ld3(nf);
            const float *m = s->material + ORC_MAT_STRIDE * (size_t)obj;
            float roughness = m[ORC_MAT_ROUGHNESS] * m[ORC_MAT_ROUGHNESS];
            float metallic = m[ORC_MAT_METALLIC];
            float transmission = m[ORC_MAT_TRANSMISSION];

            uint32_t b0[4], b1[4], b2[4];
            orc_rng_block(seed, pixel, sample, bounce, 0, b0);
            orc_rng_block(seed, pixel, sample, bounce, 1, b1);
            orc_rng_block(seed, pixel, sample, bounce, 2, b2);
            float r1[3], r2[3];
            orc_random_in_unit_sphere(b1, r1);
            orc_random_in_unit_sphere(b2, r2);

            v3 diffuse = vadd(n, ld3(r1));
            if (vdot(diffuse, n) < 0.0f) diffuse = vneg(diffuse);

            v3 perfect = reflect3(rd, n);
            v3 off = vmuls(ld3(r2), roughness);
            v3 specular = vnormalize(vadd(perfect, off));

            int transmission_ray = orc_u01(b0[0]) < transmission;
            v3 dir; uint32_t lobe;
            if (transmission_ray) {
                float ior = m[ORC_MAT_IOR];
                if (front) ior = 1.0f / ior;
                v3 rdn = vnormalize(rd);
                float cos_theta = rmin(vdot(rdn, vneg(n)), 1.0f);
                float q = (ior - 1.0f) / (ior + 1.0f);
                float r0 = q * q;                                   /* powi(2) */
                float w = 1.0f - cos_theta;
                float w2 = w * w;
                float w5 = w * (w2 * w2);                           /* powi(5) */
                float refl = r0 + (1.0f - r0) * w5;
                if (refl < orc_u01(b0[1]) && can_refract3(rdn, n, ior)) {
                    uint32_t b3[4]; float r3[3];
                    orc_rng_block(seed, pixel, sample, bounce, 3, b3);
                    orc_random_in_unit_sphere(b3, r3);
                    v3 refracted = refract3(rdn, n, ior);
                    dir = vnormalize(vadd(refracted, vmuls(ld3(r3), roughness)));
                    lobe = ORC_LOBE_REFRACT;
                } else {
                    dir = specular; lobe = ORC_LOBE_SPECULAR;
                }
            } else if (orc_u01(b0[1]) < metallic) {
                dir = specular; lobe = ORC_LOBE_SPECULAR;
            } else {
                if (orc_u01(b0[2]) < roughness) { dir = diffuse; lobe = ORC_LOBE_DIFFUSE; }
                else { dir = specular; lobe = ORC_LOBE_SPECULAR; }
            }

            v3 offset = transmission_ray ? dir : n;
            ro = vadd(P, vmuls(offset, 0.0001f));
            rd = dir;
            if (vmag2(rd) < 1e-10f) rd = n;

            atten = vmul(atten, ld3(m + ORC_MAT_ALBEDO));
            light = vadd(light, vmuls(ld3(m + ORC_MAT_EMISSION), m[ORC_MAT_EMISSION_STRENGTH]));

            if (st) st->lobe_count[lobe]++;
            if (steps) {
                OrcPathStep *ps = &steps[written++];
                ps->object = obj; ps->lobe = lobe; ps->front_face = front; ps->t = t;
                st3(ps->position, P); st3(ps->normal, n); st3(ps->origin, ro); st3(ps->direction, rd);
                st3(ps->attenuation, atten); st3(ps->light, light);
            }
        } else {
            light = vadd(light, vmul(world_sample3(s, rd), atten));
            if (st) st->lobe_count[ORC_LOBE_MISS]++;
            if (steps) {
                OrcPathStep *ps = &steps[written++];
                memset(ps, 0, sizeof *ps);
                ps->object = -1; ps->lobe = ORC_LOBE_MISS;
                st3(ps->origin, ro); st3(ps->direction, rd);
                st3(ps->attenuation, atten); st3(ps->light, light);
            }
            break;
        }
    }
    if (st && bounce == max_bounces) st->exhausted++;
    if (n_steps) *n_steps = written;
    rgba[0] = light.x; rgba[1] = light.y; rgba[2] = light.z; rgba[3] = 1.0f;
}

uint32_t orc_trace_path(const OrcScene *s, uint32_t x, uint32_t y, uint32_t sample, uint64_t seed,
                        uint32_t max_bounces, OrcPathStep *steps, float rgba[4])
{
    uint32_t n = 0;
    per_pixel(s, NULL, x, y, sample, seed, max_bounces, steps, &n, rgba, NULL);
    return n;
}

static void stats_merge(OrcStats *dst, const OrcStats *src)
{
    dst->samples += src->samples; dst->trace_calls += src->trace_calls;
    dst->primitive_tests += src->primitive_tests; dst->exhausted += src->exhausted;
    for (int i = 0; i < 64; ++i) dst->alive_at_bounce[i] += src->alive_at_bounce[i];
    for (int i = 0; i < 4; ++i) dst->lobe_count[i] += src->lobe_count[i];
}

/* cpu.rs:193-219: one sample for every pixel, row-major, accumulated in sample order.
 * The per-pixel sum order (sample_begin, sample_begin+1, ...) is what the reference produces;
 * splitting rows over threads does not change any pixel's value. */
void orc_render(const OrcScene *s, uint64_t seed, uint32_t sample_begin, uint32_t sample_end,
                uint32_t max_bounces, uint32_t row_begin, uint32_t row_end, float *accum,
                int n_threads, OrcStats *stats)
{
    const int64_t W = s->width;
    if (row_end > s->height) row_end = s->height;
    if (n_threads <= 1) {
        OrcStats local; memset(&local, 0, sizeof local);
        for (uint32_t smp = sample_begin; smp < sample_end; ++smp)
            for (int64_t y = row_begin; y < (int64_t)row_end; ++y)
                for (int64_t x = 0; x < W; ++x) {
                    float c[4];
                    per_pixel(s, NULL, (uint32_t)x, (uint32_t)y, smp, seed, max_bounces, NULL, NULL, c, stats ? &local : NULL);
                    float *px = accum + 4 * (y * W + x);
                    px[0] = px[0] + c[0]; px[1] = px[1] + c[1]; px[2] = px[2] + c[2]; px[3] = px[3] + c[3];
                }
        if (stats) stats_merge(stats, &local);
        return;
    }
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads)
#endif
    {
        OrcStats local; memset(&local, 0, sizeof local);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 2)
#endif
        for (int64_t y = row_begin; y < (int64_t)row_end; ++y)
            for (uint32_t smp = sample_begin; smp < sample_end; ++smp)
                for (int64_t x = 0; x < W; ++x) {
                    float c[4];
                    per_pixel(s, NULL, (uint32_t)x, (uint32_t)y, smp, seed, max_bounces, NULL, NULL, c, stats ? &local : NULL);
                    float *px = accum + 4 * (y * W + x);
                    px[0] = px[0] + c[0]; px[1] = px[1] + c[1]; px[2] = px[2] + c[2]; px[3] = px[3] + c[3];
                }
        if (stats) {
#ifdef _OPENMP
#pragma omp critical
#endif
            stats_merge(stats, &local);
        }
    }
}


/* the same accumulation for the pixel rectangle [x0, x0 + w) x [y0, y0 + h) only, into a w*h*4 buffer; bvh may be NULL
 * (linear scan, the reference's algorithm).  Bounded samples of large frames for bench.py's CPU legs. */
void orc_render_region(const OrcScene *s, const OrcBvh *bvh, uint64_t seed, uint32_t sample_begin, uint32_t sample_end,
                       uint32_t max_bounces, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h, float *accum,
                       int n_threads, OrcStats *stats)
{
    if (x0 + w > s->width) w = s->width > x0 ? s->width - x0 : 0;
    if (y0 + h > s->height) h = s->height > y0 ? s->height - y0 : 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
#endif
    {
        OrcStats local; memset(&local, 0, sizeof local);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int64_t y = 0; y < (int64_t)h; ++y)
            for (uint32_t smp = sample_begin; smp < sample_end; ++smp)
                for (int64_t x = 0; x < (int64_t)w; ++x) {
                    float c[4];
                    per_pixel(s, bvh, x0 + (uint32_t)x, y0 + (uint32_t)y, smp, seed, max_bounces, NULL, NULL, c, stats ? &local : NULL);
                    float *px = accum + 4 * (y * (int64_t)w + x);
                    px[0] = px[0] + c[0]; px[1] = px[1] + c[1]; px[2] = px[2] + c[2]; px[3] = px[3] + c[3];
                }
        if (stats) {
#ifdef _OPENMP
#pragma omp critical
#endif
            stats_merge(stats, &local);
        }
    }
}

/* cpu.rs:221-230: ((sum / n).clamp(0,1) * 255) as u8 -- `as` truncates, saturates, NaN -> 0 */
static inline uint8_t quantise(float sum, float n)
{
    float v = sum / n;
    if (v < 0.0f) v = 0.0f; else if (v > 1.0f) v = 1.0f;     /* f32::clamp keeps NaN */
    v = v * 255.0f;
    if (v != v) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}
void orc_resolve(const float *accum, uint64_t n_pixels, uint32_t sample_count, uint8_t *rgba8)
{
    float n = (float)sample_count;
    for (uint64_t i = 0; i < n_pixels * 4; ++i) rgba8[i] = quantise(accum[i], n);
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

"""Arithmetic the hot path takes from third-party crates that are NOT in /root/reference (Cargo.lock pins: rand 0.8.5,
cgmath 0.18.0, ordered-float 4.6.0; `f32::powi` = compiler-rt `__powisf2` / LLVM's constant-exponent expansion).

No network here, so nothing below was fetched: each case restates the PUBLISHED formula of the pinned version (file and
item named in the docstring, recalled from the published sources) in exact integer / binary-fraction arithmetic,
derives the edge values from it, and checks the oracle (CPU) and the CUDA path (device KAT entry points) against them.
It does not lift "parity unpinned" -- only a run of the Rust reference could -- but it pins the oracle to the crates'
documented behaviour rather than to a second restatement by the same author."""
import ctypes as C
from fractions import Fraction

import numpy as np
import pytest

F32_EPS = Fraction(1, 2 ** 23)


def f32(x):
    return np.float32(x)


def frac_to_f32(q: Fraction) -> np.float32:
    """round-to-nearest-even of an exact rational to f32 (via float64 is exact enough here: all cases have < 53 significant bits)"""
    return np.float32(float(q))


EDGE_WORDS = [0, 1, 0xFF, 0x100, 0x1FF, 0x200, 0x7FFFFFFF, 0x80000000, 0x800000FF, 0xFFFFFE00, 0xFFFFFEFF, 0xFFFFFF00, 0xFFFFFFFF,
              0x12345678, 0xDEADBEEF]


def standard_f32(word: int) -> np.float32:
    """rand 0.8.5 src/distributions/float.rs, `impl Distribution<f32> for Standard`: "Multiply-based method; 24 random
    bits; [0, 1) interval": value >> (32 - 24) scaled by 1 / 2^24.  (Its own test there: all-zero bits -> 0.0, the lowest
    kept bit -> EPSILON / 2, all-one bits -> 1 - EPSILON / 2.)"""
    return frac_to_f32(Fraction(word >> 8, 2 ** 24))


def uniform_inclusive_pm1(word: int) -> np.float32:
    """rand 0.8.5 src/distributions/uniform.rs, `UniformFloat<f32>`: `new_inclusive(low, high)` sets
    max_rand = (u32::MAX >> 9).into_float_with_exponent(0) - 1 = 1 - 2^-23, scale = (high - low) / max_rand, lowered one
    ulp at a time while scale * max_rand + low > high; `sample` = (bits >> 9 as the mantissa of a float in [1, 2)) - 1,
    times scale, plus low -- two roundings, no FMA.  gen_range(-1.0..=1.0) is sample_single_inclusive = that."""
    max_rand = f32(1.0) - f32(2.0 ** -23)
    scale = f32(2.0) / max_rand                                  # rounds to the f32 above 2.0
    assert scale.view(np.uint32) == 0x40000001
    assert not (scale * max_rand + f32(-1.0) > f32(1.0))         # the loop does not lower it
    value0_1 = frac_to_f32(Fraction(word >> 9, 2 ** 23))         # exact
    return f32(f32(value0_1 * scale) + f32(-1.0))


def test_rand_085_float_maps_oracle(orc):
    L = orc.lib()
    assert f32(L.orc_u01(0)) == 0.0
    assert f32(L.orc_u01(1 << 8)) == f32(float(F32_EPS / 2))
    assert f32(L.orc_u01(0xFFFFFFFF)) == f32(1.0) - f32(float(F32_EPS / 2))
    for w in EDGE_WORDS:
        assert f32(L.orc_u01(w)).view(np.uint32) == standard_f32(w).view(np.uint32), hex(w)
        assert f32(L.orc_range_pm1(w)).view(np.uint32) == uniform_inclusive_pm1(w).view(np.uint32), hex(w)
    assert f32(L.orc_range_pm1(0)) == -1.0 and f32(L.orc_range_pm1(0xFFFFFFFF)) == 1.0      # the closed range is reached, never exceeded
    rng = np.random.default_rng(2)
    for w in rng.integers(0, 2 ** 32, 2000, dtype=np.uint64):
        w = int(w)
        assert f32(L.orc_u01(w)).view(np.uint32) == standard_f32(w).view(np.uint32)
        v = f32(L.orc_range_pm1(w))
        assert v.view(np.uint32) == uniform_inclusive_pm1(w).view(np.uint32) and -1.0 <= v <= 1.0


def cgmath_normalize(v):
    """cgmath 0.18 src/structure.rs `InnerSpace::normalize`: `self * (one / self.magnitude())`, magnitude = sqrt(dot(self,
    self)), and vector.rs `dot` for Vector3 = x*x + y*y + z*z summed left to right -- one rounding per operation."""
    x, y, z = (f32(c) for c in v)
    d = f32(f32(f32(x * x) + f32(y * y)) + f32(z * z))
    inv = f32(f32(1.0) / np.sqrt(d, dtype=np.float32))
    return np.array([x * inv, y * inv, z * inv], np.float32)


def test_random_in_unit_sphere_is_a_normalised_cube_point(orc):
    """utils/mod.rs:47-55 = three gen_range(-1.0..=1.0) draws, then cgmath normalize."""
    L = orc.lib()
    rng = np.random.default_rng(3)
    out = (C.c_float * 3)()
    for _ in range(500):
        words = [int(w) for w in rng.integers(0, 2 ** 32, 3, dtype=np.uint64)]
        L.orc_random_in_unit_sphere((C.c_uint32 * 3)(*words), out)
        want = cgmath_normalize([uniform_inclusive_pm1(w) for w in words])
        assert np.array_equal(np.array(out, np.float32).view(np.uint32), want.view(np.uint32))


def test_cgmath_matrix4_times_vector4_and_lerp(orc, default_scene):
    """cgmath 0.18 src/matrix.rs `impl Mul<Vector4<S>> for Matrix4<S>`: column-major, result = c0*x + c1*y + c2*z + c3*w
    summed left to right per component; vector.rs `lerp`: self + (other - self) * amount.  Checked through the camera ray
    (cpu.rs:234-251) and World::sample (world.rs:24-27)."""
    s = default_scene
    ip, iv = s.inv_proj.astype(np.float32).reshape(4, 4), s.inv_view.astype(np.float32).reshape(4, 4)     # rows of the reshape = columns

    def mat_vec(m, v):
        return np.array([f32(f32(f32(m[0][r] * v[0]) + f32(m[1][r] * v[1])) + f32(m[2][r] * v[2])) + f32(m[3][r] * v[3]) for r in range(4)], np.float32)

    for (x, y) in [(0, 0), (853, 479), (427, 240), (100, 333), (1, 478)]:
        u = f32(x) / f32(s.width); v = f32(1.0) - f32(y) / f32(s.height)
        clip = np.array([f32(2.0) * u - f32(1.0), f32(2.0) * v - f32(1.0), -1.0, -1.0], np.float32)
        cs = mat_vec(ip, clip); cs = np.array([c / cs[3] for c in cs], np.float32)
        wd = mat_vec(iv, cs)
        want = -cgmath_normalize(wd[:3])
        o, d = orc.camera_ray(s, x, y)
        assert np.array_equal(d.view(np.uint32), want.view(np.uint32)), (x, y)
    L = orc.lib(); cs_ = s.c_struct(); out = (C.c_float * 3)()
    for d in ([0.3, 0.8, -0.2], [0, -2, 0], [5, 0.01, 1]):
        d = np.array(d, np.float32)
        cos = f32(d[1]) / f32(np.sqrt(f32(f32(f32(d[0] * d[0]) + f32(d[1] * d[1])) + f32(d[2] * d[2])), dtype=np.float32) * f32(1.0))
        amount = f32(f32(cos + f32(1.0)) * f32(0.5))
        want = np.array([f32(b + f32(f32(a - b) * amount)) for a, b in zip(s.world_a, s.world_b)], np.float32)
        L.orc_world_sample(C.byref(cs_), (C.c_float * 3)(*d.tolist()), out)
        assert np.array_equal(np.array(out, np.float32).view(np.uint32), want.view(np.uint32))


def test_powi_order_in_the_fresnel_term(orc, default_scene):
    """`f32::powi(n)` with a constant n: compiler-rt `__powisf2` (and LLVM's expansion of llvm.powi with a constant
    exponent) square-and-multiply from the low bit: powi(2) = x*x, powi(5) = x * ((x*x) * (x*x)) -- the Schlick term of
    cpu.rs:289-293.  Observed through the lobe choice of a glass bounce: the draw u2 sits between the two candidate
    roundings for inputs where (x*x*x*x)*x and x*((x*x)*(x*x)) differ."""
    rng = np.random.default_rng(4)
    n_diff = 0
    for _ in range(20000):
        x = f32(rng.uniform(0.01, 1.0))
        x2 = f32(x * x)
        a = f32(x * f32(x2 * x2))                           # __powisf2 order
        b = f32(f32(f32(f32(x * x) * x) * x) * x)           # naive left-to-right product
        n_diff += int(a.view(np.uint32) != b.view(np.uint32))
    assert n_diff > 1000          # the order is observable, so the restatement must pick the right one (oracle header, scatter)


@pytest.mark.gpu
def test_rand_085_float_maps_device(rb, orc):
    """the same maps and random_in_unit_sphere computed on the device (rdr_kat_vec RDR_KAT_RAND_FLOATS)"""
    r = rb.Renderer(rb.RendererConfig(1, 1))
    rng = np.random.default_rng(5)
    words = np.concatenate([np.array([[w, (w * 7 + 3) & 0xFFFFFFFF, (w ^ 0x55555555)] for w in EDGE_WORDS], np.uint64),
                            rng.integers(0, 2 ** 32, (3000, 3), dtype=np.uint64)]).astype(np.uint32)
    rec = np.zeros((len(words), 12), np.float32)
    rec[:, :3] = words.view(np.float32)
    got = r.kat_vec(rb.KAT_RAND_FLOATS, rec)
    for k, (w0, w1, w2) in enumerate(words.tolist()):
        assert got[k, 0].view(np.uint32) == standard_f32(w0).view(np.uint32), k
        assert got[k, 1].view(np.uint32) == uniform_inclusive_pm1(w0).view(np.uint32), k
        cube = [uniform_inclusive_pm1(w) for w in (w0, w1, w2)]
        if any(c != 0 for c in cube):
            assert np.array_equal(got[k, 2:5].view(np.uint32), cgmath_normalize(cube).view(np.uint32)), k
    r.close()

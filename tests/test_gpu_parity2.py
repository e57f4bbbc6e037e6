"""-m gpu, second file: device KATs of the shading helpers, the headline configuration's accumulator against the oracle,
convergence on benchmark.rscn, resolve edge values, sample indices near 2^32, the committed golden fixtures on the GPU,
and the AUTO fallbacks the advisor asked for."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.npz")


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def renderer(rb):
    r = rb.Renderer(rb.RendererConfig(max_sample_count=8, max_bounces=12))
    yield r
    r.close()


def same_f32(a, b):
    """bitwise equal, except that any NaN equals any NaN: x86 and the GPU produce different quiet-NaN payloads for 0/0,
    and Rust gives no guarantee about them either"""
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _unit(rng, n):
    v = rng.normal(size=(n, 3)).astype(np.float32)
    return (v / np.sqrt((v * v).sum(1, keepdims=True, dtype=np.float32))).astype(np.float32)


def test_reflect_refract_can_refract_kats(rb, renderer, orc):
    """utils/mod.rs:14-44 on the device against the oracle: random vectors, the cos(theta) > 1 clamp (|v| slightly above
    1 and v = -n), grazing incidence, and ratios on both sides of total internal reflection."""
    L = orc.lib()
    rng = np.random.default_rng(31)
    n = 4000
    v, nn = _unit(rng, n), _unit(rng, n)
    ratio = rng.choice(np.array([1 / 1.5, 1.5, 1.0, 1 / 1.3, 1.8, 2.4, 0.4], np.float32), n)
    v[:200] = -nn[:200]                                                  # head-on: cos = 1 (or a hair above after rounding)
    v[200:400] = (-nn[200:400] * np.float32(1.0000005)).astype(np.float32)   # cos > 1: the min(.., 1) clamp
    t = _unit(rng, 200); t -= nn[400:600] * (t * nn[400:600]).sum(1, keepdims=True); v[400:600] = t.astype(np.float32)   # grazing (not unit: reflect is linear)
    # the TIR boundary: sin(theta) * ratio within a few ulp of 1
    for k in range(600, 1000):
        r = np.float32(ratio[k] if ratio[k] > 1 else 1.5); ratio[k] = r
        s = np.float32(1.0) / r * np.float32(1 + (k % 5 - 2) * 6e-8)
        c = np.sqrt(max(0.0, 1.0 - float(s) ** 2))
        a = _unit(rng, 1)[0]; a -= nn[k] * (a * nn[k]).sum(); a /= np.linalg.norm(a)
        v[k] = (a * s - nn[k] * c).astype(np.float32)
    rec = np.zeros((n, 12), np.float32); rec[:, 0:3] = v; rec[:, 3:6] = nn; rec[:, 6] = ratio
    got_refl = renderer.kat_vec(rb.KAT_REFLECT, rec)
    got_refr = renderer.kat_vec(rb.KAT_REFRACT, rec)
    got_can = renderer.kat_vec(rb.KAT_CAN_REFRACT, rec)
    out = (C.c_float * 3)()
    n_tir = 0
    for k in range(n):
        L.orc_reflect(_f3(v[k]), _f3(nn[k]), out)
        assert same_f32(np.array(out, np.float32), got_refl[k, :3]), k
        L.orc_refract(_f3(v[k]), _f3(nn[k]), float(ratio[k]), out)
        assert same_f32(np.array(out, np.float32), got_refr[k, :3]), k
        can = L.orc_can_refract(_f3(v[k]), _f3(nn[k]), float(ratio[k]))
        assert can == int(got_can[k, 0]), k
        n_tir += 0 if can else 1
    assert 100 < n_tir < n - 100                                          # both outcomes exercised


def test_world_sample_and_closest_hit_kats(rb, renderer, orc, default_scene):
    """World::sample (world.rs:17-34) incl. non-unit, axis-aligned and zero directions (0/0 = NaN propagates as in the
    reference); closest_hit (cpu.rs:354-394) incl. cube face ties (edges, corners) and +-0 components of P - c."""
    import copy
    L = orc.lib()
    rng = np.random.default_rng(32)
    n = 1500
    d = (rng.normal(size=(n, 3)) * rng.uniform(0.01, 30, (n, 1))).astype(np.float32)
    d[:6] = np.array([[0, 1, 0], [0, -1, 0], [1, 0, 0], [0, 0, -3], [0, 0, 0], [1e-30, 0, 0]], np.float32)
    sky = copy.copy(default_scene)
    solid = copy.copy(default_scene); solid.world_kind = rb.WORLD_SOLID; solid.world_a = np.array([0.2, 0.4, 0.6], np.float32)
    for scene, kind in ((sky, 0.0), (solid, 1.0)):
        rec = np.zeros((n, 12), np.float32); rec[:, 0:3] = d; rec[:, 3] = kind; rec[:, 4:7] = scene.world_a; rec[:, 7:10] = scene.world_b
        got = renderer.kat_vec(rb.KAT_WORLD_SAMPLE, rec)
        s = scene.c_struct(); out = (C.c_float * 3)()
        for k in range(n):
            L.orc_world_sample(C.byref(s), _f3(d[k]), out)
            assert same_f32(np.array(out, np.float32), got[k, :3]), (kind, k)
    # closest_hit through a two-object scene: object 0 a sphere, object 1 a cube
    sc = copy.copy(default_scene)
    sc.kind = np.array([orc.SPHERE, orc.CUBE], np.uint32)
    sc.geom = np.array([[0.5, -1.0, 2.0, 1.25], [-1.0, 0.5, 0.25, 2.0]], np.float32)
    sc.material = sc.material[:2].copy()
    s = sc.c_struct()
    m = 3000
    o = rng.uniform(-4, 4, (m, 3)).astype(np.float32)
    obj = rng.integers(0, 2, m)
    # aim at a point of the primitive's surface so that P lands on (or within rounding of) it
    target = np.zeros((m, 3), np.float32)
    for k in range(m):
        c, size = sc.geom[obj[k], :3], sc.geom[obj[k], 3]
        if obj[k] == 0:
            target[k] = c + _unit(rng, 1)[0] * size
        else:
            p = rng.uniform(-1, 1, 3) * size / 2
            mode = k % 4
            p[rng.integers(0, 3)] = rng.choice([-1, 1]) * size / 2                    # a face
            if mode >= 1: p[rng.integers(0, 3)] = rng.choice([-1, 1]) * size / 2      # likely an edge: two equal distances
            if mode == 2: p[:] = rng.choice([-1, 1], 3) * size / 2                     # a corner: three-way tie
            if mode == 3: p[rng.integers(0, 3)] = 0.0                                 # l_i = +-0: signum(+-0) = +-1
            target[k] = c + p.astype(np.float32)
    dd = (target - o).astype(np.float32)
    t = np.ones(m, np.float32)                                                        # P = o + d * 1
    o[::9] = target[::9]; t[::9] = 0.0                                                # t = 0 with P exactly on the surface point
    dd[::9] = (rng.normal(size=(len(dd[::9]), 3))).astype(np.float32)
    dd[5::9, 1] = -0.0
    rec = np.zeros((m, 12), np.float32)
    rec[:, 0:3] = o; rec[:, 3:6] = dd; rec[:, 6] = t; rec[:, 7] = (obj == 0); rec[:, 8:11] = sc.geom[obj, :3]; rec[:, 11] = sc.geom[obj, 3]
    got = renderer.kat_vec(rb.KAT_CLOSEST_HIT, rec)
    p_o = (C.c_float * 3)(); n_o = (C.c_float * 3)(); ff = C.c_uint32(0)
    faces = set()
    for k in range(m):
        L.orc_closest_hit(C.byref(s), int(obj[k]), _f3(o[k]), _f3(dd[k]), float(t[k]), p_o, n_o, C.byref(ff))
        assert same_f32(np.array(p_o, np.float32), got[k, 0:3]), k
        assert same_f32(np.array(n_o, np.float32), got[k, 3:6]), k                      # bitwise: the sign of a zero component too
        assert int(ff.value) == int(got[k, 6]), k
        if obj[k] == 1:
            faces.add(tuple(np.array(n_o, np.float32).tolist()))
    assert len(faces) == 6


def test_quantise_kat_edge_values(rb, renderer, orc):
    """print_frame_buffer's `(clamp(sum / n, 0, 1) * 255) as u8` (cpu.rs:224-228) on the device for +-inf, NaN, negative,
    huge, denormal and exact-boundary sums, and n = 0 (0/0 = NaN -> 0, x/0 = inf -> 255)."""
    sums = np.array([0.0, -0.0, 1.0, 0.5, 254.999 / 255, 1e-45, -1e-45, -3.5, 7.0, 1e38, -1e38, np.inf, -np.inf, np.nan,
                     0.999999, 1.0000001, 3.0, 2.9999998, 128.0, 127.99999], np.float32)
    for n in (1, 3, 128, 0):
        rec = np.zeros((len(sums), 12), np.float32); rec[:, 0] = sums; rec[:, 1] = n
        got = renderer.kat_vec(rb.KAT_QUANTISE, rec)[:, 0].astype(np.uint8)
        acc = np.zeros((len(sums), 4), np.float32); acc[:, 0] = sums
        want = orc.resolve(acc, n)[:, 0]
        assert np.array_equal(got, want), n


def test_headline_configuration_accumulator_vs_oracle(rb, orc, benchmark_scene):
    """benchmark.rscn at 1920 x 1080 (the bench workload's frame) x 2 spp on ACCEL_AUTO -- the kernel, CTA shape and
    clustering the headline number is measured with: accumulator and RGBA8 image bit-identical to the oracle."""
    scene = benchmark_scene.with_resolution(1920, 1080)
    spp, seed = 2, 0x5EED
    want = orc.render(scene, seed, 0, spp, 12, n_threads=orc.max_threads())
    r = rb.Renderer(rb.RendererConfig(spp, 12)); r.set_seed(seed)
    img = r.render_frame(scene)
    assert np.array_equal(u32(want), u32(r.read_accum()))
    assert np.array_equal(orc.resolve(want, spp), img)
    # and through a pinned host image (the bench's e2e buffer)
    pinned = rb.HostImage(1080, 1920)
    r.render_frame(scene, out=pinned.array)
    assert np.array_equal(pinned.array, img)
    pinned.close()
    r.close()


def test_converged_psnr_benchmark_scene(rb, orc, benchmark_scene):
    """North-star gate on scenes/benchmark.rscn: independent RNG streams (GPU seed 1, CPU seed 2) converge to the same
    image -- PSNR >= 40 dB, per-channel mean within 0.5 %.  48 x 27 so that the CPU side (24576 spp, 32 M samples)
    finishes in seconds on the box; the GPU side is cheap and runs 262144 spp, so the residual noise is the CPU's."""
    scene = benchmark_scene.with_resolution(48, 27)
    spp_g, spp_c = 262144, 24576
    r = rb.Renderer(rb.RendererConfig(spp_g, 12)); r.set_seed(1)
    r.render_frame(scene)
    gpu = r.read_accum()[..., :3] / spp_g
    cpu = orc.render(scene, 2, 0, spp_c, 12, n_threads=orc.max_threads())[..., :3] / spp_c
    for c in range(3):
        assert abs(gpu[..., c].mean() - cpu[..., c].mean()) / cpu[..., c].mean() < 5e-3
    mse = float(((np.clip(gpu, 0, 1) - np.clip(cpu, 0, 1)) ** 2).mean())
    psnr = 10 * np.log10(1.0 / mse)
    assert psnr >= 40.0, psnr
    r.close()


def test_sample_indices_near_2_32(rb, orc, default_scene):
    """rdr_set_sample_offset close to the top of the u32 sample index (the Philox counter word): no truncation on the way."""
    scene = default_scene.with_resolution(96, 54)
    first = 0xFFFFFFF0
    want = orc.render(scene, 5, first, first + 8, 12, n_threads=orc.max_threads())
    r = rb.Renderer(rb.RendererConfig(8, 12)); r.set_seed(5); r.set_sample_offset(first)
    r.render_frame(scene)
    assert np.array_equal(u32(want), u32(r.read_accum()))
    r.close()


def test_golden_fixtures_on_the_gpu(rb, default_scene, benchmark_scene):
    """tests/golden/oracle_golden.npz (committed; made by tests/golden/make_golden.py) against the CUDA path: first-hit
    images, accumulators, RGBA8 images and the recorded single paths."""
    import hashlib
    g = np.load(GOLDEN)
    seed = 0x5EED0001
    digest = lambda a: np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)
    r = rb.Renderer(rb.RendererConfig(8, 12)); r.set_seed(seed)
    for name, scene in (("default", default_scene), ("benchmark1080", benchmark_scene.with_resolution(1920, 1080))):
        r.new_frame(scene)
        ids, ts = r.first_hit()
        assert np.array_equal(digest(ids), g[f"{name}_first_hit_ids_sha256"])
        assert np.array_equal(digest(ts), g[f"{name}_first_hit_t_sha256"])
        assert np.array_equal(ids[::16, ::16], g[f"{name}_first_hit_ids_sub"])
    for name, scene, spp in (("default", default_scene.with_resolution(214, 120), 8),
                             ("benchmark", benchmark_scene.with_resolution(160, 90), 4)):
        r.set_max_sample_count(spp)
        img = r.render_frame(scene)
        assert np.array_equal(u32(r.read_accum()), u32(g[f"{name}_accum"]))
        assert np.array_equal(img, g[f"{name}_rgba8"])
        ints, f32s = g[f"{name}_paths_int"], g[f"{name}_paths_f32"]
        row = 0
        while row < len(ints):
            x, y, sample = (int(v) for v in ints[row, :3])
            steps, _ = r.trace_path(x, y, sample)
            for b, s in enumerate(steps):
                assert tuple(ints[row, 3:7]) == (b, s.object, s.lobe, s.front_face)
                got = np.array([s.t, *s.position, *s.normal, *s.origin, *s.direction, *s.attenuation, *s.light], np.float32)
                assert np.array_equal(u32(got), u32(f32s[row])), (name, x, y, sample, b)
                row += 1
    r.close()


def test_auto_falls_back_to_the_hierarchy_when_the_scan_does_not_fit(rb, orc):
    """<= 1024 objects whose clustering needs more than 128 top-level entries (many primitives far larger
    than the median each stay alone) must not fail in AUTO: the frame is packed as a hierarchy instead."""
    import synth_scenes as ss
    base = ss.config4(600, 96, 54)
    geom = base.geom.copy()
    geom[1:300, 3] *= 40.0                                   # 299 large primitives -> their own top-level entries
    import copy
    scene = copy.copy(base); scene.geom = geom
    r = rb.Renderer(rb.RendererConfig(2, 12)); r.set_seed(4)
    r.new_frame(scene)                                       # AUTO
    ids_o, t_o = orc.first_hit(scene)
    ids_g, t_g = r.first_hit()
    assert np.array_equal(ids_o, ids_g) and np.array_equal(u32(t_o), u32(t_g))
    r.render_frame(scene)
    assert np.array_equal(u32(orc.render(scene, 4, 0, 2, 12, n_threads=orc.max_threads())), u32(r.read_accum()))
    # rdr_set_accel between frames of one handle: the new search applies from the next new_frame, the current frame
    # keeps the layout it was packed for
    r.set_accel(rb.ACCEL_BVH)
    r.render_samples(0)
    ids_g2, _ = r.first_hit()
    assert np.array_equal(ids_o, ids_g2)
    small = ss.config4(200, 96, 54)
    want = orc.render(small, 4, 0, 2, 12, n_threads=orc.max_threads())
    for accel in (rb.ACCEL_FUSED, rb.ACCEL_BVH, rb.ACCEL_BVH_COOP, rb.ACCEL_BRUTE, rb.ACCEL_AUTO):
        r.set_accel(accel)
        if accel == rb.ACCEL_BVH:                            # set, but no new frame yet: the fused frame stays valid
            r.new_frame(small); r.set_accel(rb.ACCEL_FUSED); r.render_samples(2)
            assert np.array_equal(u32(want), u32(r.read_accum()))
            r.set_accel(accel)
        r.render_frame(small)
        assert np.array_equal(u32(want), u32(r.read_accum())), accel
    r.close()


def test_resume_from_a_saved_accumulator_and_resolve_of_edge_sums(rb, orc, default_scene):
    """rdr_write_accum: (1) a progressive frame saved after k samples and restored on another handle continues to the same
    accumulator bit for bit (the reference keeps this state only in memory, cpu.rs:113-114); (2) resolve_kernel itself
    (not only the quantiser KAT) on accumulators holding +-inf, NaN, negative and huge sums."""
    scene = default_scene.with_resolution(96, 54)
    a = rb.Renderer(rb.RendererConfig(6, 12)); a.set_seed(3)
    a.render_frame(scene); want = a.read_accum()
    a.new_frame(scene); a.render_samples(2); saved = a.read_accum()
    b = rb.Renderer(rb.RendererConfig(6, 12)); b.set_seed(3)
    b.new_frame(scene); b.write_accum(saved, 2)
    assert b.sample_count() == 2
    img = b.finish_frame()
    assert np.array_equal(u32(b.read_accum()), u32(want)) and b.sample_count() == 6
    assert np.array_equal(img, orc.resolve(want, 6))
    with pytest.raises(rb.RaydarError):
        b.write_accum(saved, 7)                               # more samples than the frame has
    edge = np.zeros((54, 96, 4), np.float32)
    vals = np.array([np.inf, -np.inf, np.nan, -3.5, 1e38, -1e38, 0.0, -0.0, 2.9999998, 3.0, 1e-45, 765.0, 764.9999], np.float32)
    edge.reshape(-1)[:len(vals) * 40] = np.tile(vals, 40)
    b.new_frame(scene); b.write_accum(edge, 3)
    assert np.array_equal(b.resolve(), orc.resolve(edge, 3))
    assert np.array_equal(b.resolve(0), orc.resolve(edge, 3))
    b.write_accum(edge, 0)                                    # n = 0: 0/0 = NaN -> 0, x/0 = +-inf -> 255 / 0
    assert np.array_equal(b.resolve(), orc.resolve(edge, 0))
    a.close(); b.close()

"""-m gpu: the CUDA path (through the C ABI) against the oracle.  Bit-exact everywhere: the kernels
evaluate the reference's f32 expression trees with one rounding per operation."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def renderer(rb):
    r = rb.Renderer(rb.RendererConfig(max_sample_count=8, max_bounces=12))
    yield r
    r.close()


def random_rays(rng, n, scale=10.0):
    o = rng.uniform(-scale, scale, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d *= rng.uniform(0.05, 2.0, (n, 1)).astype(np.float32)          # bounce directions are not unit length
    return np.concatenate([o, d], axis=1).astype(np.float32)


def test_rng_kat(renderer, orc):
    for args in [(0, 0, 0, 0, 0), (0x5EED, 12345, 7, 3, 2), (0xFFFFFFFFFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 11, 3)]:
        assert renderer.kat_rng(*args) == orc.rng_block(*args)


@pytest.mark.parametrize("sphere", [True, False])
def test_intersection_kat(renderer, orc, sphere):
    rng = np.random.default_rng(7 + sphere)
    n = 200_000
    rays = random_rays(rng, n)
    prims = np.concatenate([rng.uniform(-10, 10, (n, 3)), rng.uniform(0.05, 6.0, (n, 1))], axis=1).astype(np.float32)
    # a share of rays aimed at their primitive (hits, grazing hits), and axis-aligned directions (division by zero)
    aim = rng.random(n) < 0.6
    tgt = prims[:, :3] + rng.normal(size=(n, 3)).astype(np.float32) * prims[:, 3:4] * 0.6
    rays[aim, 3:] = (tgt - rays[:, :3])[aim] * rng.uniform(0.1, 1.5, (aim.sum(), 1)).astype(np.float32)
    axis = rng.random(n) < 0.05
    rays[axis, 3 + rng.integers(0, 3)] = 0.0
    inside = rng.random(n) < 0.05
    rays[inside, :3] = prims[inside, :3]
    h_o, t_o = (orc.hit_sphere_batch if sphere else orc.hit_cube_batch)(rays, prims)
    h_g, t_g = renderer.kat_hit(sphere, rays, prims)
    assert h_o.sum() > n // 10
    assert np.array_equal(h_o, h_g)
    assert np.array_equal(u32(t_o), u32(t_g))


@pytest.mark.parametrize("name,res", [("default", None), ("benchmark", (1920, 1080)), ("benchmark", None)])
@pytest.mark.parametrize("cull", [True, False, "bvh", "cluster", "coop", "fused", "bvh2"])
def test_first_hit_ids_bit_exact(rb, renderer, orc, default_scene, benchmark_scene, name, res, cull):
    scene = default_scene if name == "default" else benchmark_scene
    if res:
        scene = scene.with_resolution(*res)
    if cull is False and scene.width > 1920:
        pytest.skip("exact-everything reference run only up to 1080p")
    ids_o, t_o = orc.first_hit(scene)
    renderer.debug_set_cull(cull is not False)
    renderer.set_accel({"bvh": rb.ACCEL_BVH, "cluster": rb.ACCEL_CLUSTER, "coop": rb.ACCEL_COOP, "fused": rb.ACCEL_FUSED, "bvh2": rb.ACCEL_BVH_COOP}.get(cull, rb.ACCEL_BRUTE))
    try:
        renderer.new_frame(scene)
        ids_g, t_g = renderer.first_hit()
    finally:
        renderer.debug_set_cull(True)
        renderer.set_accel(rb.ACCEL_AUTO)
    assert np.array_equal(ids_o, ids_g)
    assert np.array_equal(u32(t_o), u32(t_g))


def test_camera_rays(renderer, orc, default_scene, benchmark_scene):
    for scene in (default_scene, benchmark_scene):
        renderer.new_frame(scene)
        rng = np.random.default_rng(3)
        xy = np.stack([rng.integers(0, scene.width, 500), rng.integers(0, scene.height, 500)], axis=1).astype(np.uint32)
        rays = renderer.kat_camera_rays(xy)
        for (x, y), ray in zip(xy, rays):
            o, d = orc.camera_ray(scene, int(x), int(y))
            assert np.array_equal(u32(o), u32(ray[:3])) and np.array_equal(u32(d), u32(ray[3:]))


@pytest.mark.parametrize("accel", ["brute", "bvh", "cluster", "coop", "fused", "bvh2"])
def test_trace_arbitrary_rays(rb, renderer, orc, benchmark_scene, accel):
    rng = np.random.default_rng(11)
    rays = random_rays(rng, 20_003, scale=8.0)            # not a multiple of 32: partial warps
    rays[:, 1] = np.abs(rays[:, 1])                       # origins above the floor
    rays[::50, 3 + (np.arange(len(rays[::50])) % 3)] = 0.0   # axis-parallel components (division by zero in the slabs)
    rays[::77, :3] *= 1e4                                 # origins far outside the scene bound (no culling)
    renderer.set_accel({"bvh": rb.ACCEL_BVH, "cluster": rb.ACCEL_CLUSTER, "coop": rb.ACCEL_COOP, "fused": rb.ACCEL_FUSED, "bvh2": rb.ACCEL_BVH_COOP}.get(accel, rb.ACCEL_BRUTE))
    try:
        renderer.new_frame(benchmark_scene.with_resolution(64, 36))
        ids_g, t_g = renderer.kat_trace(rays)
    finally:
        renderer.set_accel(rb.ACCEL_AUTO)
    for i in range(0, len(rays), 7):
        idx, t = orc.trace(benchmark_scene, rays[i, :3], rays[i, 3:])
        assert idx == ids_g[i]
        if idx >= 0:
            assert u32(np.float32(t)) == u32(t_g[i])


@pytest.mark.parametrize("accel", ["brute", "bvh", "cluster", "coop", "fused", "bvh2"])
@pytest.mark.parametrize("name", ["default", "benchmark"])
def test_single_path_debug_mode(rb, orc, default_scene, benchmark_scene, name, accel):
    scene = (default_scene if name == "default" else benchmark_scene.with_resolution(480, 270))
    seed = 0xC0FFEE
    r = rb.Renderer(rb.RendererConfig(max_sample_count=4, max_bounces=12))
    r.set_seed(seed)
    r.set_accel({"bvh": rb.ACCEL_BVH, "cluster": rb.ACCEL_CLUSTER, "coop": rb.ACCEL_COOP, "fused": rb.ACCEL_FUSED, "bvh2": rb.ACCEL_BVH_COOP}.get(accel, rb.ACCEL_BRUTE))
    r.new_frame(scene)
    rng = np.random.default_rng(5)
    fields = ["position", "normal", "origin", "direction", "attenuation", "light"]
    n_steps = 0
    for _ in range(300):
        x, y, s = int(rng.integers(0, scene.width)), int(rng.integers(0, scene.height)), int(rng.integers(0, 1000))
        steps_o, rgba_o = orc.trace_path(scene, x, y, s, seed, 12)
        steps_g, rgba_g = r.trace_path(x, y, s)
        assert len(steps_o) == len(steps_g)
        for a, b in zip(steps_o, steps_g):
            assert (a.object, a.lobe, a.front_face) == (b.object, b.lobe, b.front_face)
            assert u32(np.float32(a.t)) == u32(np.float32(b.t))
            for f in fields:
                assert np.array_equal(u32(np.array(getattr(a, f)[:])), u32(np.array(getattr(b, f)[:]))), f
            n_steps += 1
        # north-star tolerance: fixed-seed single path within 1e-4 relative -- met exactly
        assert np.array_equal(u32(rgba_o), u32(rgba_g))
    assert n_steps > 400
    r.close()


@pytest.mark.parametrize("accel", ["brute", "bvh", "cluster", "coop", "fused", "bvh2"])
@pytest.mark.parametrize("name,res,spp", [("default", (427, 240), 16), ("benchmark", (320, 180), 8)])
def test_accumulator_bit_exact(rb, orc, default_scene, benchmark_scene, name, res, spp, accel):
    scene = (default_scene if name == "default" else benchmark_scene).with_resolution(*res)
    seed = 99
    acc_o = orc.render(scene, seed, 0, spp, 12, n_threads=orc.max_threads())
    r = rb.Renderer(rb.RendererConfig(max_sample_count=spp, max_bounces=12))
    r.set_seed(seed)
    r.set_accel({"bvh": rb.ACCEL_BVH, "cluster": rb.ACCEL_CLUSTER, "coop": rb.ACCEL_COOP, "fused": rb.ACCEL_FUSED, "bvh2": rb.ACCEL_BVH_COOP}.get(accel, rb.ACCEL_BRUTE))
    img = r.render_frame(scene)
    acc_g = r.read_accum()
    assert np.array_equal(u32(acc_o), u32(acc_g))
    assert np.array_equal(orc.resolve(acc_o, spp), img)
    assert r.sample_count() == spp
    r.close()


def test_progressive_equals_one_shot_and_sharding(rb, benchmark_scene):
    """Size-independent properties at 1080p: (1) render_sample x n == render_frame(n) bit for bit,
    (2) two sample-range shards sum to the single render up to f32 summation order."""
    scene = benchmark_scene.with_resolution(1920, 1080)
    spp = 4
    a = rb.Renderer(rb.RendererConfig(spp, 12)); a.set_seed(5)
    img_a = a.render_frame(scene); acc_a = a.read_accum()
    b = rb.Renderer(rb.RendererConfig(spp, 12)); b.set_seed(5)
    b.new_frame(scene)
    last = None
    for i in range(spp):
        last = b.render_sample(scene)
        assert last is not None and b.sample_count() == i + 1
    assert b.render_sample(scene) is None
    assert np.array_equal(u32(acc_a), u32(b.read_accum()))
    assert np.array_equal(img_a, last)
    shards = []
    for g in range(2):
        c = rb.Renderer(rb.RendererConfig(spp // 2, 12)); c.set_seed(5); c.set_sample_offset(g * spp // 2)
        c.new_frame(scene); c.render_samples(spp // 2); shards.append(c.read_accum()); c.close()
    total = shards[0] + shards[1]
    assert np.allclose(total, acc_a, rtol=1e-5, atol=1e-5)
    assert np.array_equal(total[..., 3], acc_a[..., 3])
    a.close(); b.close()


@pytest.mark.parametrize("count,rows", [(2, 16), (3, 7), (8, 16)])
def test_row_stripe_shards_sum_bit_identical(rb, orc, default_scene, count, rows):
    """Image-tile sharding (rdr_set_row_stripes): every shard renders all samples of its round-robin row stripes into
    a zeroed W x H accumulator; the shards are disjoint, so their sum is bit-identical to the whole-image render
    and to the oracle (107 x 61: the last stripe is short)."""
    scene = default_scene.with_resolution(107, 61)
    spp, seed = 5, 99
    want = orc.render(scene, seed, 0, spp, 12, n_threads=orc.max_threads())
    total = np.zeros_like(want)
    for index in range(count):
        r = rb.Renderer(rb.RendererConfig(spp, 12)); r.set_seed(seed); r.set_row_stripes(rows, index, count)
        r.new_frame(scene); r.render_samples(spp)
        acc = r.read_accum()
        mine = np.zeros(61, bool); mine[[y for y in range(61) if (y // rows) % count == index]] = True
        assert np.all(u32(acc[~mine]) == 0)
        assert np.array_equal(u32(acc[mine]), u32(want[mine]))
        total += acc
        r.close()
    assert np.array_equal(u32(total), u32(want))


def test_row_stripes_1080p_fused_and_reset(rb, benchmark_scene):
    """The same property at the bench resolution on the fused scan, and rdr_set_row_stripes(…, count = 1) restores the
    whole image on the same handle."""
    scene = benchmark_scene.with_resolution(1920, 1080)
    spp = 2
    a = rb.Renderer(rb.RendererConfig(spp, 12)); a.set_seed(5)
    a.render_frame(scene); whole = a.read_accum()
    total = np.zeros_like(whole)
    for index in range(2):
        a.set_row_stripes(16, index, 2)
        a.new_frame(scene); a.render_samples(spp)
        total += a.read_accum()
    assert np.array_equal(u32(total), u32(whole))
    a.set_row_stripes(0, 0, 1)
    a.new_frame(scene); a.render_samples(spp)
    assert np.array_equal(u32(a.read_accum()), u32(whole))
    with pytest.raises(rb.RaydarError):
        a.set_row_stripes(16, 2, 2)
    a.close()


def test_converged_psnr_independent_streams(rb, orc, default_scene):
    """North-star gate: at high spp the converged GPU image reaches PSNR >= 40 dB against the CPU backend's
    converged image, with bounded per-channel mean error.  The two sides use DIFFERENT RNG seeds (independent
    streams), so this checks convergence to the same expectation; the same-stream comparison
    (test_accumulator_bit_exact) is bit-exact.  64x36 so that the CPU side finishes in seconds at 32768 spp."""
    scene = default_scene.with_resolution(64, 36)
    spp = 32768
    r = rb.Renderer(rb.RendererConfig(spp, 12)); r.set_seed(1)
    r.render_frame(scene)
    gpu = r.read_accum()[..., :3] / spp
    cpu = orc.render(scene, 2, 0, spp, 12, n_threads=orc.max_threads())[..., :3] / spp
    for c in range(3):
        assert abs(gpu[..., c].mean() - cpu[..., c].mean()) / cpu[..., c].mean() < 2e-3
    mse = float(((np.clip(gpu, 0, 1) - np.clip(cpu, 0, 1)) ** 2).mean())
    psnr = 10 * np.log10(1.0 / mse)
    assert psnr >= 40.0, psnr
    r.close()


def test_bvh_config4_100k_objects(rb, orc):
    """BASELINE config 4: 100k random Spheres/Cubes + ground cube (hierarchy in global memory / L2, not staged).
    First-hit ids and t against the O(N) linear scan of the oracle at a resolution it finishes in seconds, then the
    accumulator of a small multi-bounce render."""
    import synth_scenes as ss
    scene = ss.config4(100_000, 160, 90)
    ids_o, t_o = orc.first_hit(scene)
    r = rb.Renderer(rb.RendererConfig(2, 12)); r.set_seed(3)          # ACCEL_AUTO -> cooperative hierarchy (n > threshold)
    small = scene.with_resolution(48, 27)
    want = orc.render(small, 3, 0, 2, 12, n_threads=orc.max_threads())
    assert len(np.unique(ids_o)) > 3000
    for accel in (rb.ACCEL_AUTO, rb.ACCEL_BVH):                       # warp-cooperative and per-lane traversal
        r.set_accel(accel)
        r.new_frame(scene)
        ids_g, t_g = r.first_hit()
        assert np.array_equal(ids_o, ids_g), accel
        assert np.array_equal(u32(t_o), u32(t_g)), accel
        r.render_frame(small)
        assert np.array_equal(u32(want), u32(r.read_accum())), accel
    # the brute-force scan cannot hold this scene in shared memory: explicit request is an error, not a fallback
    r.set_accel(rb.ACCEL_BRUTE)
    with pytest.raises(rb.RaydarError) as e:
        r.new_frame(scene)
    assert e.value.status == rb.ERR_UNSUPPORTED
    r.close()


def test_bvh_config5_glass_metal_32_bounces(rb, orc):
    """BASELINE config 5: glass/metal lattice in a closed box, 32 bounces; scan and hierarchy against the oracle."""
    import synth_scenes as ss
    scene = ss.config5(192, 108)
    want = orc.render(scene, 9, 0, 2, 32, n_threads=orc.max_threads())
    for accel in (rb.ACCEL_BRUTE, rb.ACCEL_BVH, rb.ACCEL_CLUSTER, rb.ACCEL_COOP, rb.ACCEL_FUSED, rb.ACCEL_BVH_COOP):
        r = rb.Renderer(rb.RendererConfig(2, 32)); r.set_seed(9); r.set_accel(accel)
        r.render_frame(scene)
        assert np.array_equal(u32(want), u32(r.read_accum())), accel
        r.close()


@pytest.mark.parametrize("n", [40, 300, 700, 1000])
def test_fused_scan_cluster_sizes(rb, orc, n):
    """The fused scan keeps <= 32 top-level entries by growing the clusters (8, 16, 24, 32 members): random scenes of
    n objects + a ground cube, first hits, arbitrary rays (incl. degenerate ones) and a small multi-bounce render
    against the oracle's linear scan."""
    import synth_scenes as ss
    scene = ss.config4(n, 96, 54)
    r = rb.Renderer(rb.RendererConfig(2, 12)); r.set_seed(21); r.set_accel(rb.ACCEL_FUSED)
    r.new_frame(scene)
    ids_o, t_o = orc.first_hit(scene)
    ids_g, t_g = r.first_hit()
    assert np.array_equal(ids_o, ids_g) and np.array_equal(u32(t_o), u32(t_g))
    rng = np.random.default_rng(n)
    rays = random_rays(rng, 4001, scale=60.0)
    rays[:, 1] = np.abs(rays[:, 1]) * 0.5
    rays[::50, 3 + (np.arange(len(rays[::50])) % 3)] = 0.0
    rays[::77, :3] *= 1e4
    ids_g, t_g = r.kat_trace(rays)
    for i in range(0, len(rays), 5):
        idx, t = orc.trace(scene, rays[i, :3], rays[i, 3:])
        assert idx == ids_g[i], i
        if idx >= 0:
            assert u32(np.float32(t)) == u32(t_g[i])
    want = orc.render(scene, 21, 0, 2, 12, n_threads=orc.max_threads())
    r.render_frame(scene)
    assert np.array_equal(u32(want), u32(r.read_accum()))
    r.close()


def test_renderer_trait_session(rb, orc, default_scene, benchmark_scene):
    """The trait's session semantics on one handle (renderer/mod.rs:25-35): timers, getters, a resolution and scene change
    between frames (frame_buffer() reallocation, cpu.rs:404-411), setters in mid-frame (inspector.rs:178-204), the
    SolidColor world, and a non-16:9 frame through an orthographic-style matrix pair (the path has no special case)."""
    import copy
    r = rb.Renderer(rb.RendererConfig(4, 12)); r.set_seed(11)
    a = default_scene.with_resolution(160, 90)
    r.render_frame(a)
    p = r.profiler()                                            # main.rs:61-107 fails if any of the four timers is unset
    assert p.has_frame and p.has_prepare and p.has_render and p.has_sample and p.device_render_ms > 0
    assert (r.sample_count(), r.max_sample_count(), r.max_bounces()) == (4, 4, 12)
    assert np.array_equal(u32(orc.render(a, 11, 0, 4, 12)), u32(r.read_accum()))
    b = benchmark_scene.with_resolution(200, 112)
    r.set_max_sample_count(2); r.set_max_bounces(5)
    img = r.render_frame(b)
    assert img.shape == (112, 200, 4) and img[..., 3].min() == 255
    assert np.array_equal(u32(orc.render(b, 11, 0, 2, 5)), u32(r.read_accum()))
    r.set_max_sample_count(5)                                   # raising the budget continues the same accumulation
    n = 0
    while r.render_sample(b) is not None:
        n += 1
    assert n == 3 and r.sample_count() == 5
    want = orc.render(b, 11, 0, 5, 5)
    assert np.array_equal(u32(want), u32(r.read_accum()))
    assert np.array_equal(orc.resolve(want, 5), r.resolve())
    c = copy.copy(a); c.world_kind = rb.WORLD_SOLID; c.world_a = np.array([0.25, 0.5, 0.75], np.float32)
    r.set_max_sample_count(3); r.set_max_bounces(12)
    r.render_frame(c)
    assert np.array_equal(u32(orc.render(c, 11, 0, 3, 12)), u32(r.read_accum()))
    e = copy.copy(default_scene).with_resolution(97, 131)       # odd, portrait frame: partial warps, matrices used as stored
    e.inv_proj = np.diag([3.0, 4.0, 1.0, 1.0]).astype(np.float32).T.reshape(-1)     # an orthographic-style inverse projection
    r.render_frame(e)
    assert np.array_equal(u32(orc.render(e, 11, 0, 3, 12)), u32(r.read_accum()))
    ids_o, t_o = orc.first_hit(e); ids_g, t_g = r.first_hit()
    assert np.array_equal(ids_o, ids_g) and np.array_equal(u32(t_o), u32(t_g))
    r.close()


def test_edge_cases(rb, orc, default_scene):
    # zero objects: every sample is the sky
    empty = default_scene.with_resolution(32, 16)
    import copy
    empty = copy.copy(empty)
    empty.kind = empty.kind[:0]; empty.geom = empty.geom[:0]; empty.material = empty.material[:0]
    r = rb.Renderer(rb.RendererConfig(3, 12)); r.set_seed(1)
    img = r.render_frame(empty)
    acc_o = orc.render(empty, 1, 0, 3, 12)
    assert np.array_equal(u32(acc_o), u32(r.read_accum()))
    assert np.array_equal(orc.resolve(acc_o, 3), img)
    # max_bounces = 0 and 1
    for mb in (0, 1, 2):
        r.set_max_bounces(mb)
        s = default_scene.with_resolution(64, 36)
        r.render_frame(s)
        assert np.array_equal(u32(orc.render(s, 1, 0, 3, mb)), u32(r.read_accum())), mb
    # max_sample_count = 0: the reference divides by zero -> all-zero image
    r.set_max_sample_count(0); r.set_max_bounces(12)
    img0 = r.render_frame(default_scene.with_resolution(16, 8))
    assert img0.max() == 0
    # transparent world is todo!() in the reference -> error status, no abort
    t = copy.copy(default_scene); t.world_kind = rb.WORLD_TRANSPARENT
    with pytest.raises(rb.RaydarError) as e:
        r.new_frame(t)
    assert e.value.status == rb.ERR_UNSUPPORTED
    r.close()

"""Oracle checks (CPU): RNG known answers, the second (numpy) restatement, scene anchors from the survey,
golden fixtures.  The reference has no tests or golden vectors for this path and cannot run here
(parity unpinned), so these are the strongest pins available."""
import os

import numpy as np
import pytest

import np_restatement as npr

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.npz")


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_philox_random123_known_answers(orc):
    # Random123 kat_vectors, philox4x32 10 rounds
    assert orc.philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert orc.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert orc.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_rng_block_counter_layout(orc):
    assert orc.rng_block(0x1122334455667788, 5, 6, 2, 3) == orc.philox([5, 6, 2 * 4 + 3, 0], [0x55667788, 0x11223344])


def test_rand_float_maps(orc):
    L = orc.lib()
    assert L.orc_u01(0) == 0.0
    assert L.orc_u01(0xFFFFFFFF) == np.float32(1.0 - 2.0 ** -24)            # [0, 1)
    assert L.orc_u01(0x000000FF) == 0.0                                      # low 8 bits discarded
    assert L.orc_range_pm1(0) == -1.0
    assert L.orc_range_pm1(0xFFFFFFFF) == 1.0                                # inclusive upper end
    assert L.orc_range_pm1(0x80000000) == np.float32(np.float32(0.5) * np.float32(2.0000002) - np.float32(1.0))
    w = np.random.default_rng(0).integers(0, 2 ** 32, 1000, dtype=np.uint64)
    vals = np.array([L.orc_range_pm1(int(x)) for x in w])
    assert vals.min() >= -1 and vals.max() <= 1 and abs(vals.mean()) < 0.1


def test_random_in_unit_sphere_is_normalised_cube_point(orc):
    import ctypes as C
    words = (C.c_uint32 * 3)(0x12345678, 0x9ABCDEF0, 0x0F1E2D3C)
    out = (C.c_float * 3)()
    orc.lib().orc_random_in_unit_sphere(words, out)
    p = np.array([orc.lib().orc_range_pm1(w) for w in words], np.float32)
    inv = np.float32(1.0) / np.sqrt((p[0] * p[0] + p[1] * p[1]) + p[2] * p[2])
    assert np.array_equal(u32(np.array(out, np.float32)), u32(p * inv))


@pytest.mark.parametrize("sphere", [True, False])
def test_intersections_match_numpy_restatement(orc, sphere):
    rng = np.random.default_rng(42 + sphere)
    n = 300_000
    o = rng.uniform(-10, 10, (n, 3)).astype(np.float32)
    d = (rng.normal(size=(n, 3)) * rng.uniform(0.05, 2.0, (n, 1))).astype(np.float32)
    c = rng.uniform(-10, 10, (n, 3)).astype(np.float32)
    s = rng.uniform(0.05, 6.0, n).astype(np.float32)
    aim = rng.random(n) < 0.6
    tgt = c + rng.normal(size=(n, 3)).astype(np.float32) * s[:, None] * np.float32(0.6)
    d[aim] = ((tgt - o) * rng.uniform(0.1, 1.5, (n, 1)).astype(np.float32))[aim]
    zero = rng.random(n) < 0.05
    d[zero, rng.integers(0, 3)] = 0.0                                          # +-inf / NaN slabs
    inside = rng.random(n) < 0.05
    o[inside] = c[inside]
    rays = np.concatenate([o, d], 1); prims = np.concatenate([c, s[:, None]], 1)
    h_c, t_c = (orc.hit_sphere_batch if sphere else orc.hit_cube_batch)(rays, prims)
    h_n, t_n = (npr.hit_sphere if sphere else npr.hit_cube)(o, d, c, s)
    assert h_c.sum() > n // 10
    assert np.array_equal(h_c.astype(bool), h_n)
    assert np.array_equal(u32(t_c), u32(t_n))


def test_camera_rays_match_numpy_restatement(orc, default_scene, benchmark_scene):
    for scene in (default_scene, benchmark_scene.with_resolution(480, 270)):
        d_np = npr.camera_rays(scene.width, scene.height, scene.inv_proj, scene.inv_view)
        rng = np.random.default_rng(1)
        for _ in range(400):
            x, y = int(rng.integers(0, scene.width)), int(rng.integers(0, scene.height))
            o, d = orc.camera_ray(scene, x, y)
            assert np.array_equal(u32(d), u32(d_np[y, x]))
            assert np.array_equal(o, scene.cam_pos)


def test_first_hit_matches_numpy_restatement(orc, default_scene, benchmark_scene):
    for scene in (default_scene.with_resolution(427, 240), benchmark_scene.with_resolution(480, 270)):
        ids_c, t_c = orc.first_hit(scene)
        ids_n, t_n = npr.first_hit(scene)
        assert np.array_equal(ids_c, ids_n)
        assert np.array_equal(u32(t_c), u32(t_n))


def test_survey_anchors(orc, default_scene, benchmark_scene):
    """anchors measured independently at survey time."""
    o, d = orc.camera_ray(default_scene, default_scene.width // 2, default_scene.height // 2)
    fwd = default_scene.extra["target"] - default_scene.extra["position"]
    fwd = fwd / np.linalg.norm(fwd)
    assert np.allclose(d, fwd, atol=2e-7 * 4)
    ids, _ = orc.first_hit(benchmark_scene.with_resolution(1920, 1080))
    assert (ids < 0).sum() == 0
    u, c = np.unique(ids, return_counts=True)
    assert len(u) == 135
    share = {int(a): round(100.0 * b / ids.size, 2) for a, b in zip(u, c)}
    assert share[0] == 54.28 and share[1] == 2.98 and share[47] == 2.37 and share[55] == 1.47 and share[85] == 1.47


def test_tie_break_first_minimum_wins(orc, default_scene):
    """min_by_key keeps the first minimum (cpu.rs:349): two coincident cubes -> the lower index."""
    import copy
    s = copy.copy(default_scene)
    s.kind = np.array([1, 1, 0], np.uint32)
    s.geom = np.array([[0, 0, 5, 2], [0, 0, 5, 2], [0, 0, 5, 1]], np.float32)
    s.material = np.tile(default_scene.material[1], (3, 1))
    idx, t = orc.trace(s, [0, 0, 0], [0, 0, 1])
    assert idx == 0 and t == 4.0
    s.geom = np.array([[0, 0, 5, 2], [0, 0, 5, 2.5], [0, 0, 4.5, 1]], np.float32)
    idx, t = orc.trace(s, [0, 0, 0], [0, 0, 1])
    assert idx == 2 and t == 3.5


def test_resolve_semantics(orc):
    """print_frame_buffer (cpu.rs:221-230): clamp, *255, truncating saturating cast, NaN -> 0."""
    acc = np.array([[[0.0, 2.0, 8.0, 4.0], [-1.0, np.nan, np.inf, 1.9999]]], np.float32)
    out = orc.resolve(acc, 4)
    assert out.tolist() == [[[0, 127, 255, 255], [0, 0, 255, 127]]]
    assert orc.resolve(acc, 0).tolist() == [[[0, 255, 255, 255], [0, 0, 255, 255]]]      # x/0: 0/0 = NaN -> 0, +/0 = inf -> 255


def test_workload_statistics(orc, benchmark_scene):
    """mean trace_ray calls per sample on benchmark.rscn @12 bounces ~ 3.1."""
    s = benchmark_scene.with_resolution(240, 135)
    _, st = orc.render(s, 1, 0, 4, 12, n_threads=orc.max_threads(), want_stats=True)
    b = st.trace_calls / st.samples
    assert 3.0 < b < 3.2
    assert st.primitive_tests == st.trace_calls * 183
    alive = [st.alive_at_bounce[i] / st.samples for i in range(4)]
    assert alive[0] == 1.0 and alive[1] == 1.0 and 0.5 < alive[2] < 0.56 and 0.24 < alive[3] < 0.28


def test_render_thread_count_does_not_change_pixels(orc, default_scene):
    s = default_scene.with_resolution(107, 60)
    a = orc.render(s, 3, 0, 3, 12, n_threads=1)
    b = orc.render(s, 3, 0, 3, 12, n_threads=4)
    assert np.array_equal(u32(a), u32(b))
    # sample ranges compose: [0,2) then [2,3) on the same accumulator == [0,3)
    c = orc.render(s, 3, 0, 2, 12)
    c = orc.render(s, 3, 2, 3, 12, accum=c)
    assert np.array_equal(u32(a), u32(c))


def test_golden_fixtures(orc, default_scene, benchmark_scene):
    """tests/golden/oracle_golden.npz (made by tests/golden/make_golden.py from this oracle): guards the oracle
    against drift; the GPU tests compare against the same fixtures on the box."""
    g = np.load(GOLDEN)
    from golden import make_golden
    fresh = make_golden.compute(orc, default_scene, benchmark_scene)
    assert set(g.files) == set(fresh)
    for k in g.files:
        a, b = g[k], fresh[k]
        if a.dtype == np.float32:
            assert np.array_equal(u32(a), u32(b)), k
        else:
            assert np.array_equal(a, b), k

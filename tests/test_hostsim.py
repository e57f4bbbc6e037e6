"""CPU checks of the PRODUCT's per-lane code: raydar_b200/csrc/rdr_core.cuh / rdr_trace.cuh / rdr_pack.h are
__host__ __device__ and compiled here for the host (tests/hostsim) so that what the CUDA kernels execute --
conservative cull + exact test, winner selection, sample-refill loop, scatter, RNG -- is compared with the
oracle without a GPU.  The GPU run of the same code is tests/test_gpu_parity.py."""
import copy

import numpy as np
import pytest

import synth_scenes as ss


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_rng_matches_oracle(hs, orc):
    rng = np.random.default_rng(0)
    for _ in range(200):
        seed = int(rng.integers(0, 2 ** 63)); args = [int(v) for v in rng.integers(0, 2 ** 32, 2)] + [int(rng.integers(0, 64)), int(rng.integers(0, 4))]
        assert hs.rng_block(seed, *args) == orc.rng_block(seed, *args)


@pytest.mark.parametrize("cull", [True, False, 2, 3])    # flat scan + cull, flat scan exact-everything, BVH, cluster scan
def test_first_hit_bit_exact_1080p(hs, orc, benchmark_scene, cull):
    scene = benchmark_scene.with_resolution(1920, 1080)
    ids_o, t_o = orc.first_hit(scene)
    ids, ts, st = hs.first_hit(scene, cull)
    assert np.array_equal(ids, ids_o)
    assert np.array_equal(u32(ts), u32(t_o))
    if cull:   # the cull / the hierarchy keep ~1.4-1.8 of 183 primitives per primary ray
        assert (st.sphere_exact + st.cube_exact) / st.traces < 3.0
    if cull in (2, 3):
        assert 1.0 <= st.nodes_visited / st.traces < 12.0


def test_first_hit_default_scene(hs, orc, default_scene):
    ids_o, t_o = orc.first_hit(default_scene)
    for mode in (True, 2, 3):
        ids, ts, _ = hs.first_hit(default_scene, mode)
        assert np.array_equal(ids, ids_o) and np.array_equal(u32(ts), u32(t_o))


@pytest.mark.parametrize("name,res,spp,bounces", [("benchmark", (240, 135), 6, 12), ("default", (214, 120), 8, 12),
                                                  ("benchmark", (96, 54), 4, 32), ("default", (64, 36), 4, 1)])
def test_accumulator_bit_exact(hs, orc, default_scene, benchmark_scene, name, res, spp, bounces):
    scene = (default_scene if name == "default" else benchmark_scene).with_resolution(*res)
    want = orc.render(scene, 77, 0, spp, bounces, n_threads=orc.max_threads())
    for cull in (True, False, 2, 3):
        got, _ = hs.render(scene, 77, 0, spp, bounces, use_cull=cull)
        assert np.array_equal(u32(got), u32(want)), (name, cull)
    assert np.array_equal(hs.resolve(want, spp), orc.resolve(want, spp))


def test_sample_ranges_compose(hs, orc, default_scene):
    scene = default_scene.with_resolution(107, 60)
    whole, _ = hs.render(scene, 5, 0, 6, 12)
    part, _ = hs.render(scene, 5, 0, 2, 12)
    part, _ = hs.render(scene, 5, 2, 4, 12, accum=part)
    assert np.array_equal(u32(whole), u32(part))


def test_single_path_debug_mode(hs, orc, default_scene, benchmark_scene):
    fields = ["position", "normal", "origin", "direction", "attenuation", "light"]
    for scene in (default_scene, benchmark_scene.with_resolution(480, 270)):
        rng = np.random.default_rng(9)
        for _ in range(150):
            x, y, s = int(rng.integers(0, scene.width)), int(rng.integers(0, scene.height)), int(rng.integers(0, 500))
            steps_o, rgba_o = orc.trace_path(scene, x, y, s, 31337, 12)
            steps_h, rgba_h = hs.trace_path(scene, x, y, s, 31337, 12, use_cull=(True, 2, 3)[s % 3])
            assert len(steps_o) == len(steps_h)
            for a, b in zip(steps_o, steps_h):
                assert (a.object, a.lobe, a.front_face) == (b.object, b.lobe, b.front_face)
                assert u32(np.float32(a.t)) == u32(np.float32(b.t))
                for f in fields:
                    assert np.array_equal(u32(np.array(getattr(a, f)[:])), u32(np.array(getattr(b, f)[:])))
            assert np.array_equal(u32(rgba_o), u32(rgba_h))


def test_golden_paths_and_accum(hs, default_scene, benchmark_scene):
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.npz"))
    from golden.make_golden import SEED
    for name, scene, spp in (("default", default_scene.with_resolution(214, 120), 8), ("benchmark", benchmark_scene.with_resolution(160, 90), 4)):
        acc, _ = hs.render(scene, SEED, 0, spp, 12)
        assert np.array_equal(u32(acc), u32(g[f"{name}_accum"]))
        assert np.array_equal(hs.resolve(acc, spp), g[f"{name}_rgba8"])


def _scene_with(default_scene, kind, geom):
    s = copy.copy(default_scene)
    s.kind = np.asarray(kind, np.uint32)
    s.geom = np.asarray(geom, np.float32).reshape(-1, 4)
    s.material = np.tile(default_scene.material[1], (len(kind), 1))
    return s


def test_ties_and_list_order(hs, orc, default_scene):
    """Spheres are scanned before cubes on the device; the winner must still be the reference's first minimum
    in ORIGINAL object order (cpu.rs:349)."""
    rng = np.random.default_rng(4)
    # grid-aligned unit cubes sharing faces + spheres touching them: many exactly equal t values
    kinds, geom = [], []
    for i in range(60):
        kinds.append(int(rng.integers(0, 2)))
        c = rng.integers(-3, 4, 3).astype(np.float32)
        geom.append([c[0], c[1], c[2] + 8, 1.0 if kinds[-1] else 0.5])
    scene = _scene_with(default_scene, kinds, geom)
    n = 40_000
    o = np.zeros((n, 3), np.float32)
    d = np.concatenate([rng.integers(-4, 5, (n, 2)) / np.float32(8.0), np.ones((n, 1))], 1).astype(np.float32)
    rays = np.concatenate([o, d], 1)
    ids, ts, _ = hs.trace(scene, rays, True)
    ids2, ts2, _ = hs.trace(scene, rays, False)
    assert np.array_equal(ids, ids2) and np.array_equal(u32(ts), u32(ts2))
    for mode in (2, 3):
        ids3, ts3, _ = hs.trace(scene, rays, mode)
        assert np.array_equal(ids, ids3) and np.array_equal(u32(ts), u32(ts3))
    for i in range(0, n, 37):
        idx, t = orc.trace(scene, rays[i, :3], rays[i, 3:])
        assert idx == ids[i]
        if idx >= 0:
            assert u32(np.float32(t)) == u32(ts[i])
    assert (ids >= 0).sum() > n // 4


def test_degenerate_rays_take_exact_path(hs, orc, benchmark_scene):
    """Axis-aligned directions (a zero component -> +-inf slabs, NaN from 0/0) and far-away origins skip the cull."""
    rng = np.random.default_rng(8)
    n = 3000
    o = rng.uniform(-8, 8, (n, 3)).astype(np.float32); o[:, 1] = np.abs(o[:, 1])
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[np.arange(n), rng.integers(0, 3, n)] = 0.0
    d[: n // 3, :] = 0.0; d[np.arange(n // 3), rng.integers(0, 3, n // 3)] = rng.choice([-1.0, 1.0], n // 3)
    o[-200:] *= 1e4                                                                                # outside the scene bound
    rays = np.concatenate([o, d], 1)
    ids, ts, st = hs.trace(benchmark_scene, rays, True)
    assert st.degenerate >= n - 10
    # the hierarchy clamps 1/d instead (no O(N) fallback for axis-parallel rays); only far-away origins are degenerate
    for mode in (2, 3):
        ids_b, ts_b, st_b = hs.trace(benchmark_scene, rays, mode)
        assert np.array_equal(ids, ids_b) and np.array_equal(u32(ts), u32(ts_b))
        assert 150 <= st_b.degenerate <= 250
    for i in range(0, n, 3):
        idx, t = orc.trace(benchmark_scene, rays[i, :3], rays[i, 3:])
        assert idx == ids[i]
        if idx >= 0:
            assert u32(np.float32(t)) == u32(ts[i])


def test_auto_packs_a_hierarchy_when_the_scan_clustering_does_not_fit(hs, rb):
    """AUTO must not fail for <= 1024 objects whose clustering needs more than 128 top-level entries."""
    import copy
    base = ss.config4(1023, 64, 36)                          # 1024 objects incl. the ground cube: at the AUTO threshold
    st, mode, n_top = hs.pack_mode(base, rb.ACCEL_AUTO)
    assert st == rb.OK
    big = copy.copy(base); big.geom = base.geom.copy(); big.geom[1:300, 3] *= 40.0     # 299 large primitives stay alone
    st, mode, _ = hs.pack_mode(big, rb.ACCEL_AUTO)
    assert (st, mode) == (rb.OK, 1)
    st, _, _ = hs.pack_mode(big, rb.ACCEL_FUSED)             # an explicit scan request still reports the limit
    assert st == rb.ERR_UNSUPPORTED
    small = ss.config4(200, 64, 36)
    assert hs.pack_mode(small, rb.ACCEL_AUTO)[:2] == (rb.OK, 0)
    assert hs.pack_mode(small, rb.ACCEL_BVH_COOP)[:2] == (rb.OK, 1)

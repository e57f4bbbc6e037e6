"""The product's warp-cooperative fused scan (raydar_b200/csrc/rdr_fused.cuh: the nearest-hit search behind
RDR_ACCEL_AUTO for scenes up to ~1000 objects, i.e. the kernel of the headline benchmark) is device-only code.  Here it
runs on the CPU: tests/hostsim compiles the same header against a warp emulator (tests/hostsim/warp_emu.h: 32 fibers are
the 32 lanes; every __shfl_sync / __syncwarp is a rendezvous) and the winners are compared with the reference rule --
nearest t, first minimum on ties (cpu.rs:344-352) -- as restated by the oracle.  This covers what the per-lane twins in
test_hostsim.py cannot: the task compaction, the survivor lists, the packed box tests and the atomicMin winner fold."""
import numpy as np
import pytest

import synth_scenes as ss


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def random_rays(rng, n, scale):
    o = rng.uniform(-scale, scale, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d *= rng.uniform(0.05, 2.0, (n, 1)).astype(np.float32)           # bounce directions are not unit length
    return np.concatenate([o, d], axis=1).astype(np.float32)


def primary_rays(orc, scene, step=1):
    xs, ys = np.meshgrid(np.arange(0, scene.width, step), np.arange(0, scene.height, step))
    rays = np.zeros((xs.size, 6), np.float32)
    for i, (x, y) in enumerate(zip(xs.ravel(), ys.ravel())):
        o, d = orc.camera_ray(scene, int(x), int(y))
        rays[i, :3] = o; rays[i, 3:] = d
    return rays, xs.ravel(), ys.ravel()


def check_against_linear_scan(hs, orc, scene, rays, oracle_every):
    ids_f, t_f = hs.trace_fused(scene, rays)
    ids_l, t_l, _ = hs.trace(scene, rays, use_cull=False)             # exact test on every primitive, reference order
    assert np.array_equal(ids_f, ids_l)
    assert np.array_equal(u32(t_f), u32(t_l))
    for i in range(0, len(rays), oracle_every):                       # and the oracle itself on a subset
        idx, t = orc.trace(scene, rays[i, :3], rays[i, 3:])
        assert idx == ids_f[i], i
        if idx >= 0:
            assert u32(np.float32(t)) == u32(t_f[i])
    return ids_f


def test_fused_scan_benchmark_scene_arbitrary_rays(hs, orc, benchmark_scene):
    info = hs.fused_info(benchmark_scene)
    assert info["fused_ok"] == 1 and info["fused_cap"] == 8 and info["fused_top"] <= 32 and info["fused_direct"] >= 1
    assert info["fused_stage_bytes"] < info["blob_bytes"]              # the fused kernels stage only the prefix
    rng = np.random.default_rng(11)
    rays = random_rays(rng, 6_003, scale=8.0)                         # not a multiple of 32: a partial warp with dead lanes
    rays[:, 1] = np.abs(rays[:, 1])                                   # origins above the floor
    rays[::50, 3 + (np.arange(len(rays[::50])) % 3)] = 0.0            # axis-parallel components (1/d clamped in the slabs)
    rays[::77, :3] *= 1e4                                             # origins far outside the scene bound: no culling
    ids = check_against_linear_scan(hs, orc, benchmark_scene, rays, oracle_every=13)
    assert (ids >= 0).mean() > 0.3 and (ids < 0).any()                # hits and misses both occur


def test_fused_scan_benchmark_scene_primary_rays(hs, orc, benchmark_scene):
    scene = benchmark_scene.with_resolution(160, 90)
    rays, xs, ys = primary_rays(orc, scene)
    ids_f, t_f = hs.trace_fused(scene, rays)
    ids_o, t_o = orc.first_hit(scene)
    assert np.array_equal(ids_f, ids_o[ys, xs])
    assert np.array_equal(u32(t_f), u32(t_o[ys, xs]))


@pytest.mark.parametrize("n,cap", [(40, 8), (300, 16), (700, 24), (900, 32)])
def test_fused_scan_cluster_sizes(hs, orc, n, cap):
    """Bigger scenes keep <= 32 top-level entries by growing the clusters: 16 / 24 / 32 members take the multi-step
    member stage (trace_fused<false>)."""
    scene = ss.config4(n, 96, 54)
    info = hs.fused_info(scene)
    assert info["fused_ok"] == 1 and info["fused_cap"] == cap and info["fused_top"] <= 32
    rng = np.random.default_rng(n)
    rays = random_rays(rng, 1601, scale=60.0)
    rays[:, 1] = np.abs(rays[:, 1]) * 0.5
    rays[::50, 3 + (np.arange(len(rays[::50])) % 3)] = 0.0
    rays[::77, :3] *= 1e4
    check_against_linear_scan(hs, orc, scene, rays, oracle_every=9)
    prim, _, _ = primary_rays(orc, scene, step=2)
    check_against_linear_scan(hs, orc, scene, prim, oracle_every=11)


def test_fused_scan_direct_entries(hs, orc):
    """Large primitives stay alone at the top level and skip the member stage (spheres first, then cubes): their
    survivors go straight to the exact-test lists."""
    import copy
    base = ss.config4(80, 96, 54)
    s = copy.copy(base)
    big_kind = np.array([0, 0, 1], np.uint32)                          # two big spheres, one big cube (+ the ground cube)
    big_geom = np.array([[40, 30, 40, 25], [-60, 20, 10, 18], [10, 15, -70, 30]], np.float32)
    s.kind = np.concatenate([base.kind, big_kind]); s.geom = np.concatenate([base.geom, big_geom])
    s.material = np.concatenate([base.material, base.material[1:4]])
    info = hs.fused_info(s)
    assert info["fused_ok"] == 1 and info["fused_direct"] >= 3 and info["fused_ns_direct"] >= 2, info
    rng = np.random.default_rng(5)
    rays = random_rays(rng, 2401, scale=60.0)
    rays[:, 1] = np.abs(rays[:, 1]) * 0.5
    rays[::50, 3 + (np.arange(len(rays[::50])) % 3)] = 0.0
    ids = check_against_linear_scan(hs, orc, s, rays, oracle_every=9)
    n0 = len(base.kind)
    assert np.isin(ids, [n0, n0 + 1, n0 + 2]).any()                    # the big primitives do win for some rays
    prim, _, _ = primary_rays(orc, s, step=2)
    check_against_linear_scan(hs, orc, s, prim, oracle_every=11)


def test_fused_scan_ties_and_tiny_scenes(hs, orc, default_scene):
    """Equal t from different objects (integer-grid centres, rays through grid points) and boxes nested in boxes: the
    lowest original index wins whatever cluster or list the candidates sit in; a 3-object scene has a single top-level
    entry."""
    import copy
    assert hs.fused_info(default_scene)["fused_top"] == 1
    rays, _, _ = primary_rays(orc, default_scene.with_resolution(64, 36))
    check_against_linear_scan(hs, orc, default_scene.with_resolution(64, 36), rays, oracle_every=7)
    rng = np.random.default_rng(3)
    kind, geom = [], []
    for i in range(240):
        c = rng.integers(-4, 5, 3).astype(np.float32)
        k = int(rng.integers(0, 2))
        kind.append(k); geom.append([c[0], c[1], c[2] + 10, [0.5, 1.0][k] * float(rng.choice([1.0, 1.0, 2.0, 8.0]))])
    s = copy.copy(default_scene)
    s.kind = np.asarray(kind, np.uint32); s.geom = np.asarray(geom, np.float32)
    s.material = np.tile(default_scene.material[1], (240, 1))
    assert hs.fused_info(s)["fused_ok"] == 1
    n = 8_000
    d = np.concatenate([rng.integers(-6, 7, (n, 2)) / np.float32(8.0), np.ones((n, 1))], 1).astype(np.float32)
    rays = np.concatenate([np.zeros((n, 3), np.float32), d], 1)
    ids = check_against_linear_scan(hs, orc, s, rays, oracle_every=101)
    _, t_l, _ = hs.trace(s, rays, use_cull=False)
    assert (ids >= 0).mean() > 0.5


def test_direct_entries_by_ballot_keep_the_winner(hs, orc, benchmark_scene):
    """The single-primitive top entries (the floor of benchmark.rscn; several, spheres among them, in the synthetic
    scene) reach the survivor lists by one ballot each, and the per-ray sphere margin comes from the widened approximate
    form (exact square roots here: the emulator has no sqrt.approx): same winners as the exact linear scan."""
    import copy
    rng = np.random.default_rng(17)
    rays = random_rays(rng, 3_001, scale=8.0)
    rays[:, 1] = np.abs(rays[:, 1])
    rays[::50, 3 + (np.arange(len(rays[::50])) % 3)] = 0.0
    rays[::77, :3] *= 1e4
    ids_v, t_v = hs.trace_fused(benchmark_scene, rays)
    ids_l, t_l, _ = hs.trace(benchmark_scene, rays, use_cull=False)
    assert np.array_equal(ids_v, ids_l) and np.array_equal(u32(t_v), u32(t_l))
    base = ss.config4(80, 96, 54)                                       # several direct entries, spheres among them
    s = copy.copy(base)
    s.kind = np.concatenate([base.kind, np.array([0, 0, 1], np.uint32)])
    s.geom = np.concatenate([base.geom, np.array([[40, 30, 40, 25], [-60, 20, 10, 18], [10, 15, -70, 30]], np.float32)])
    s.material = np.concatenate([base.material, base.material[1:4]])
    rays = random_rays(rng, 2_401, scale=60.0)
    rays[:, 1] = np.abs(rays[:, 1]) * 0.5
    ids_v, t_v = hs.trace_fused(s, rays)
    ids_l, t_l, _ = hs.trace(s, rays, use_cull=False)
    assert np.array_equal(ids_v, ids_l) and np.array_equal(u32(t_v), u32(t_l))
    scene = ss.config4(700, 96, 54)                                     # 24-member clusters
    rays = random_rays(rng, 1_601, scale=60.0)
    ids_v, t_v = hs.trace_fused(scene, rays)
    ids_l, t_l, _ = hs.trace(scene, rays, use_cull=False)
    assert np.array_equal(ids_v, ids_l) and np.array_equal(u32(t_v), u32(t_l))


# ---- the warp-cooperative cluster scan (AUTO's fallback when a scene has more than 32 fused top-level entries) -----------
@pytest.mark.parametrize("name", ["benchmark", "n1000"])
def test_cooperative_cluster_scan_under_emulator(hs, orc, benchmark_scene, name):
    scene = benchmark_scene if name == "benchmark" else ss.config4(1000, 96, 54)
    if name == "n1000":
        assert hs.fused_info(scene)["fused_ok"] == 0                    # this is the scene size where AUTO runs this scan
    rng = np.random.default_rng(23)
    rays = random_rays(rng, 2_403, scale=8.0 if name == "benchmark" else 60.0)
    rays[:, 1] = np.abs(rays[:, 1])
    rays[::50, 3 + (np.arange(len(rays[::50])) % 3)] = 0.0
    rays[::77, :3] *= 1e4
    ids_e, t_e = hs.trace_coop_emu(scene, rays)
    ids_l, t_l, _ = hs.trace(scene, rays, use_cull=False)
    assert np.array_equal(ids_e, ids_l) and np.array_equal(u32(t_e), u32(t_l))
    prim, _, _ = primary_rays(orc, scene.with_resolution(96, 54), step=2)
    ids_e, t_e = hs.trace_coop_emu(scene, prim)
    ids_l, t_l, _ = hs.trace(scene, prim, use_cull=False)
    assert np.array_equal(ids_e, ids_l) and np.array_equal(u32(t_e), u32(t_l))


# ---- the warp-cooperative hierarchy (config 4's search) -------------------------------------------------------------------
@pytest.mark.parametrize("n", [300, 6_000])
def test_cooperative_hierarchy_under_emulator(hs, orc, n):
    """trace_bvh2 (rdr_bvh2.cuh): per-ray near-first stacks served 8 rays per round by groups of 4 lanes, ballot-popcount
    emission, pruning by the best exact hit -- the same winners as the exact linear scan."""
    scene = ss.config4(n, 96, 54)
    rng = np.random.default_rng(n)
    rays = random_rays(rng, 1_603, scale=90.0)
    rays[:, 1] = np.abs(rays[:, 1]) * 0.4
    rays[::50, 3 + (np.arange(len(rays[::50])) % 3)] = 0.0
    rays[::77, :3] *= 1e4
    prim, _, _ = primary_rays(orc, scene, step=2)
    for batch in (rays, prim):
        ids_e, t_e = hs.trace_bvh2_emu(scene, batch)
        for i in range(len(batch)):                                      # the oracle's linear scan (cpu.rs:344-352) for every ray
            idx, t = orc.trace(scene, batch[i, :3], batch[i, 3:])
            assert idx == ids_e[i], i
            if idx >= 0:
                assert u32(np.float32(t)) == u32(t_e[i])
    assert (hs.trace_bvh2_emu(scene, prim)[0] >= 0).mean() > 0.5


def test_cooperative_hierarchy_ties(hs, orc, default_scene):
    import copy
    rng = np.random.default_rng(3)
    kind, geom = [], []
    for i in range(400):
        c = rng.integers(-4, 5, 3).astype(np.float32)
        k = int(rng.integers(0, 2))
        kind.append(k); geom.append([c[0], c[1], c[2] + 10, [0.5, 1.0][k] * float(rng.choice([1.0, 1.0, 2.0, 8.0]))])
    s = copy.copy(default_scene)
    s.kind = np.asarray(kind, np.uint32); s.geom = np.asarray(geom, np.float32)
    s.material = np.tile(default_scene.material[1], (400, 1))
    n = 6_000
    d = np.concatenate([rng.integers(-6, 7, (n, 2)) / np.float32(8.0), np.ones((n, 1))], 1).astype(np.float32)
    rays = np.concatenate([np.zeros((n, 3), np.float32), d], 1)
    ids_e, t_e = hs.trace_bvh2_emu(s, rays)
    ids_l, t_l, _ = hs.trace(s, rays, use_cull=False)
    assert np.array_equal(ids_e, ids_l) and np.array_equal(u32(t_e), u32(t_l))


# ---- the render kernel's sample loop (rdr_loop_body.inc: the text render_kernel compiles) under the emulator ----------
@pytest.mark.parametrize("cold", [True, False])
def test_render_loop_bit_exact(hs, orc, benchmark_scene, cold):
    """One emulated CTA of 4 warps sharing the atomic pixel counter runs the kernel's loop body around the fused scan:
    persistent lanes, warp lock-step, sample refill from the cached primary hit, cold / parked lane state in
    shared-memory columns (cold=True, the shipped 896-thread variant's policy) or in registers."""
    scene = benchmark_scene.with_resolution(64, 36)
    want = orc.render(scene, 7, 0, 3, 12, n_threads=2)
    for order in (None, [3, 1, 0, 2], [0, 0, 0, 1, 2, 3]):               # different interleavings of the warps
        got = hs.render_fused_emu(scene, 7, 0, 3, 12, cold=cold, order=order)
        assert np.array_equal(u32(got), u32(want))


def test_render_loop_progressive_stripes_and_edges(hs, orc, default_scene, benchmark_scene):
    scene = benchmark_scene.with_resolution(48, 27)
    want = orc.render(scene, 5, 0, 5, 12, n_threads=2)
    acc = hs.render_fused_emu(scene, 5, 0, 2, 12)                         # two launches on one accumulator
    acc = hs.render_fused_emu(scene, 5, 2, 3, 12, accum=acc)
    assert np.array_equal(u32(acc), u32(want))
    total = np.zeros_like(want)                                          # three row-stripe shards, 5 rows per stripe
    for index in range(3):
        total += hs.render_fused_emu(scene, 5, 0, 5, 12, stripes=(5, index, 3))
    assert np.array_equal(u32(total), u32(want))
    for bounces in (0, 1):                                               # `for _ in 0..0`, and paths cut after one trace
        want_b = orc.render(scene, 5, 0, 4, bounces, n_threads=2)
        assert np.array_equal(u32(hs.render_fused_emu(scene, 5, 0, 4, bounces)), u32(want_b))
    small = default_scene.with_resolution(40, 22)                        # primary rays that miss: pixels finished by the sky
    assert np.array_equal(u32(hs.render_fused_emu(small, 9, 0, 4, 12)), u32(orc.render(small, 9, 0, 4, 12, n_threads=2)))
    tiny = benchmark_scene.with_resolution(5, 3)                         # fewer pixels than lanes: most lanes never get work
    assert np.array_equal(u32(hs.render_fused_emu(tiny, 9, 0, 3, 12)), u32(orc.render(tiny, 9, 0, 3, 12, n_threads=1)))

"""Adversarial checks of the conservative cull margins (raydar_b200/csrc/rdr_core.cuh): whenever the
reference-ordered exact test hits (and could beat the pruning bound), the FMA test must have kept the
primitive.  Run on the CPU through tests/hostsim, which executes the same __host__ __device__ code as the
kernels with identical rounding (explicit fma/mul/add)."""
import numpy as np
import pytest


def make_rays(rng, n, scale, centers, sizes, mode):
    o = rng.uniform(-scale, scale, (n, 3)).astype(np.float32)
    dlen = rng.uniform(0.02, 2.0, (n, 1)).astype(np.float32)
    if mode == "grazing":
        # aim at a point ON the primitive's silhouette (distance ~ size from the centre, perpendicular-ish)
        v = rng.normal(size=(n, 3)).astype(np.float32)
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        jitter = (1.0 + rng.normal(size=(n, 1)) * 1e-4).astype(np.float32)
        tgt = centers + v * sizes[:, None] * jitter
    elif mode == "through":
        tgt = centers + rng.normal(size=(n, 3)).astype(np.float32) * sizes[:, None] * np.float32(0.4)
    else:
        tgt = rng.uniform(-scale, scale, (n, 3)).astype(np.float32)
    d = (tgt - o)
    d = d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20) * dlen
    if mode == "surface":           # origins sitting 1e-4 off the primitive surface, like bounce rays (cpu.rs:323)
        v = rng.normal(size=(n, 3)).astype(np.float32)
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        o = (centers + v * sizes[:, None] * np.float32(1.0001)).astype(np.float32)
        d = rng.normal(size=(n, 3)).astype(np.float32) * dlen
    return np.concatenate([o, d.astype(np.float32)], 1).astype(np.float32)


@pytest.mark.parametrize("scale", [1.0, 10.0, 200.0, 5000.0])
@pytest.mark.parametrize("mode", ["grazing", "through", "random", "surface"])
def test_sphere_cull_never_drops_a_hit(hs, scale, mode):
    rng = np.random.default_rng(int(scale) + len(mode))
    n = 400_000
    c = rng.uniform(-scale, scale, (n, 3)).astype(np.float32)
    r = (rng.uniform(0.01, 1.0, n) * min(scale, 20.0) * rng.choice([0.02, 0.2, 1.0], n)).astype(np.float32)
    rays = make_rays(rng, n, scale, c, r, mode)
    prims = np.concatenate([c, r[:, None]], 1).astype(np.float32)
    q_max = float((2.0 * (c.astype(np.float64) ** 2).sum(1) + r.astype(np.float64) ** 2).max())
    origin_bound = float(np.abs(rays[:, :3]).max() * 1.001 + 1e-3)
    hit, t = hs.exact(True, rays, prims)
    may, deg = hs.sphere_cull(rays, prims, q_max, origin_bound)
    if mode != "random":
        assert hit.sum() > n // 50
    assert deg.sum() < 10
    dropped = (hit == 1) & (may == 0)
    assert dropped.sum() == 0, f"{dropped.sum()} exact hits rejected by the cull"
    # the cull must still be useful: it rejects most true misses in the random regime
    if mode == "random" and scale <= 200:
        assert ((may == 1) & (hit == 0)).sum() < 0.2 * n


@pytest.mark.parametrize("scale", [1.0, 10.0, 200.0, 5000.0])
@pytest.mark.parametrize("mode", ["grazing", "through", "random", "surface"])
def test_cube_cull_never_drops_a_hit(hs, scale, mode):
    rng = np.random.default_rng(1000 + int(scale) + len(mode))
    n = 400_000
    c = rng.uniform(-scale, scale, (n, 3)).astype(np.float32)
    side = (rng.uniform(0.02, 2.0, n) * min(scale, 20.0) * rng.choice([0.02, 0.2, 1.0, 8.0], n)).astype(np.float32)
    half = side * np.float32(0.5)
    rays = make_rays(rng, n, scale, c, half * (np.float32(1.7) if mode == "grazing" else np.float32(1.0)), mode)
    if mode == "grazing":           # slide along a face plane: origin exactly in the plane of one face
        k = rng.integers(0, 3, n)
        rays[np.arange(n), k] = (c[np.arange(n), k] + half * rng.choice([-1, 1], n)).astype(np.float32)
        rays[np.arange(n), 3 + k] *= np.float32(1e-3)
    prims = np.concatenate([c, side[:, None]], 1).astype(np.float32)
    obj_bound = float((np.abs(c).max(1) + side).max())
    origin_bound = float(max(obj_bound, np.abs(rays[:, :3]).max()) * 1.001 + 1e-3)
    pad = np.float32(2.0 ** -18) * np.float32(origin_bound + obj_bound)
    hit, t = hs.exact(False, rays, prims)
    inf = np.full(n, np.inf, np.float32)
    may, deg = hs.cube_cull(rays, prims, float(pad), origin_bound, inf)
    if mode != "random":
        assert hit.sum() > n // 50
    dropped = (hit == 1) & (may == 0)
    assert dropped.sum() == 0, f"{dropped.sum()} exact hits rejected by the cull"
    # pruning bound: a cube whose exact t equals or beats `best` must be kept (ties go to the lower index)
    best = np.where(hit == 1, t, inf).astype(np.float32)
    may2, _ = hs.cube_cull(rays, prims, float(pad), origin_bound, best)
    assert ((hit == 1) & (may2 == 0)).sum() == 0
    # ... and one ulp below the exact t it may be dropped, but a bound far below must prune most
    far_below = np.where(hit == 1, t * np.float32(0.5) - np.float32(1.0), inf).astype(np.float32)
    may3, _ = hs.cube_cull(rays, prims, float(pad), origin_bound, far_below)
    positive_t = (hit == 1) & (t > 4 * pad / np.maximum(np.linalg.norm(rays[:, 3:], axis=1), 1e-9))
    if mode == "through":
        assert (may3[positive_t] == 0).mean() > 0.5


def test_scene_constants(hs, benchmark_scene):
    q_max, origin_bound, pad = hs.scene_consts(benchmark_scene)
    g = benchmark_scene.geom
    assert origin_bound >= np.abs(benchmark_scene.cam_pos).max()
    assert origin_bound >= (np.abs(g[:, :3]).max(1) + g[:, 3]).max()
    assert 0 < pad < 0.01
    sph = g[benchmark_scene.kind == 0]
    assert q_max >= (2 * (sph[:, :3] ** 2).sum(1) + sph[:, 3] ** 2).max() * 0.999

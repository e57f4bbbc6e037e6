"""The 8-wide hierarchy (raydar_b200/csrc/rdr_bvh.h, trace_bvh in rdr_trace.cuh) must return exactly the winner of
the reference's linear scan (cpu.rs:344-352), including the first-minimum tie-break, on scenes far larger than the
scan can hold.  CPU checks through tests/hostsim; the GPU twin is tests/test_gpu_parity.py::test_bvh_*."""
import numpy as np
import pytest

import synth_scenes as ss


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_cluster_shape(hs, benchmark_scene):
    n_top, nbytes = hs.cluster_info(benchmark_scene)
    assert n_top == 24                                   # the floor cube on its own + 182 primitives in 23 clusters of <= 8
    n_top5, _ = hs.cluster_info(ss.config5(64, 36))
    assert 65 <= n_top5 <= 70


def test_hierarchy_shape(hs):
    scene = ss.config4(5000, 64, 36)
    n_nodes, mode, nbytes = hs.bvh_info(scene)
    assert mode == 1
    assert 5001 / 8 <= n_nodes <= 5001                      # 8-wide: between N/8 and N nodes
    n_nodes2, n_root, ok2 = hs.bvh2_info(scene)               # pair-packed twin: <= 32 root entries, 16 quads per node
    assert ok2 == 1 and 4 <= n_root <= 32 and 5001 / 8 / 1.5 <= n_nodes2 <= 5001
    head = 256 * n_nodes + 64 * 5001
    raw = head + (-head) % 128 + 256 * n_nodes2             # the cooperative hierarchy's nodes start on a 128-byte line
    assert nbytes == raw + (-raw) % 16


@pytest.mark.parametrize("n", [300, 20_000])
def test_config4_first_hit_matches_linear_scan(hs, orc, n):
    """BASELINE config 4 (random Spheres/Cubes over a 200 x 200 field + ground cube), reduced object count and
    resolution so that the O(N) oracle finishes in seconds."""
    scene = ss.config4(n, 256, 144)
    ids_o, t_o = orc.first_hit(scene)
    ids, ts, st = hs.first_hit(scene, hs.BVH)
    assert np.array_equal(ids, ids_o)
    assert np.array_equal(u32(ts), u32(t_o))
    assert len(np.unique(ids_o)) > min(n, 2000) // 4
    assert st.degenerate == 0
    assert (st.sphere_exact + st.cube_exact) / st.traces < 0.01 * n + 12      # a tiny fraction of the N exact tests
    ids2, ts2, st2 = hs.first_hit(scene, hs.BVH2)             # the pair-packed hierarchy holds the same winner
    assert np.array_equal(ids2, ids_o) and np.array_equal(u32(ts2), u32(t_o))
    assert (st2.sphere_exact + st2.cube_exact) / st2.traces < 0.01 * n + 12


def test_config4_secondary_rays(hs, orc):
    scene = ss.config4(3000, 96, 54)
    want = orc.render(scene, 11, 0, 3, 12, n_threads=orc.max_threads())
    got, st = hs.render(scene, 11, 0, 3, 12, use_cull=hs.BVH)
    assert np.array_equal(u32(got), u32(want))


def test_config5_glass_metal_lattice(hs, orc):
    """BASELINE config 5: 8x8x8 glass/metal lattice in a closed box, 32 bounces."""
    scene = ss.config5(160, 90)
    assert scene.n_objects == 513
    ids_o, t_o = orc.first_hit(scene)
    for mode in (True, hs.BVH, hs.CLUSTER, hs.BVH2):
        ids, ts, _ = hs.first_hit(scene, mode)
        assert np.array_equal(ids, ids_o) and np.array_equal(u32(ts), u32(t_o))
    small = scene.with_resolution(64, 36)
    want, stats = orc.render(small, 5, 0, 2, 32, n_threads=orc.max_threads(), want_stats=True)
    assert stats.trace_calls / stats.samples > 25            # closed box: paths rarely end before the bounce limit
    for mode in (True, hs.BVH, hs.CLUSTER):
        got, _ = hs.render(small, 5, 0, 2, 32, use_cull=mode)
        assert np.array_equal(u32(got), u32(want))


def test_coincident_and_nested_primitives(hs, orc, default_scene):
    """Ties (equal t from different objects) and boxes that contain other boxes: the lowest original index wins."""
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
    n = 30_000
    d = np.concatenate([rng.integers(-6, 7, (n, 2)) / np.float32(8.0), np.ones((n, 1))], 1).astype(np.float32)
    rays = np.concatenate([np.zeros((n, 3), np.float32), d], 1)
    a_ids, a_t, _ = hs.trace(s, rays, False)
    b_ids, b_t, _ = hs.trace(s, rays, hs.BVH)
    assert np.array_equal(a_ids, b_ids) and np.array_equal(u32(a_t), u32(b_t))
    c_ids, c_t, _ = hs.trace(s, rays, hs.CLUSTER)
    assert np.array_equal(a_ids, c_ids) and np.array_equal(u32(a_t), u32(c_t))
    e_ids, e_t, _ = hs.trace(s, rays, hs.BVH2)
    assert np.array_equal(a_ids, e_ids) and np.array_equal(u32(a_t), u32(e_t))
    for i in range(0, n, 101):
        idx, t = orc.trace(s, rays[i, :3], rays[i, 3:])
        assert idx == b_ids[i]


def test_rscn_round_trip(rb, orc, tmp_path):
    """The synthetic scenes can be saved as .rscn and read back by both loaders with identical arrays."""
    scene = ss.config5(320, 180)
    path = str(tmp_path / "config5.rscn")
    ss.write_rscn(scene, path)
    py = orc.load_rscn(path)
    assert np.array_equal(u32(py.geom), u32(scene.geom)) and np.array_equal(u32(py.inv_view), u32(scene.inv_view))
    f = rb.Scene.load(path).flat()
    assert f.n_objects == 513 and np.array_equal(u32(np.array(f.inv_proj[:])), u32(scene.inv_proj))
    # and the C++ update_matrices agrees with the numpy restatement used by the generator
    sc = rb.Scene.load(path).set_resolution(320, 180)
    assert np.array_equal(u32(sc.matrices()[2]), u32(scene.inv_view)) and np.array_equal(u32(sc.matrices()[3]), u32(scene.inv_proj))


def test_cluster_refinement_keeps_shape_and_enters_fewer_boxes(hs, orc, benchmark_scene):
    """refine_clusters (rdr_bvh.h): a surface-area local search on the fused clustering.  Every object stays in exactly
    one cluster, the number of clusters and the <= 8 bound are kept, no cluster falls below two members -- and rays of
    real paths enter clearly fewer cluster boxes (each entered box is a member-stage task of the fused scan)."""
    import copy
    other = copy.copy(benchmark_scene)                                   # a different geometry in between: no stale cache
    other.geom = benchmark_scene.geom.copy(); other.geom[1:, 0] += 0.25
    plain = hs.fused_clusters(benchmark_scene, variant=("norefine", ("RDR_CLUSTER_REFINE=0",)))   # compile-time switch
    hs.fused_clusters(other)
    refined = hs.fused_clusters(benchmark_scene)
    assert np.array_equal(hs.fused_clusters(benchmark_scene), refined)  # deterministic, and the cached copy is the same
    n = benchmark_scene.n_objects
    for cl in (plain, refined):
        assert cl.shape == (n,) and cl.min() == 0 and len(np.unique(cl)) == cl.max() + 1
    assert plain.max() == refined.max()
    sizes = np.bincount(refined)
    assert sizes.max() <= 8 and np.array_equal(np.sort(sizes)[:1], [1]) and (np.sort(sizes)[1:] >= 2).all()   # the floor alone
    assert (plain != refined).any()

    g = benchmark_scene.geom.astype(np.float64); half = np.where(benchmark_scene.kind == 1, g[:, 3] * 0.5, g[:, 3])
    lo, hi = g[:, :3] - half[:, None], g[:, :3] + half[:, None]
    rng = np.random.default_rng(1)
    scene = benchmark_scene.with_resolution(1920, 1080)
    rays = []
    for _ in range(400):
        x, y = int(rng.integers(0, 1920)), int(rng.integers(0, 1080))
        o, d = orc.camera_ray(scene, x, y)
        rays.append(np.concatenate([o, d]))
        steps = orc.trace_path(scene, x, y, int(rng.integers(0, 1000)), 5, 12)
        steps = steps[0] if isinstance(steps, tuple) else steps
        rays += [np.array(list(st.origin) + list(st.direction)) for st in steps if st.object >= 0]
    rays = np.array(rays, np.float64)

    def entered(cl):
        ids = [c for c in np.unique(cl) if (cl == c).sum() > 1]
        blo = np.array([lo[cl == c].min(0) for c in ids]); bhi = np.array([hi[cl == c].max(0) for c in ids])
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = (blo[None] - rays[:, None, :3]) / rays[:, None, 3:]; t2 = (bhi[None] - rays[:, None, :3]) / rays[:, None, 3:]
        tn = np.nanmax(np.minimum(t1, t2), axis=2); tf = np.nanmin(np.maximum(t1, t2), axis=2)
        return (tf >= np.maximum(tn, 0)).sum(1).mean()

    a, b = entered(plain), entered(refined)
    assert b < 0.85 * a, (a, b)

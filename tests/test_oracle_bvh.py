"""The oracle's own BVH (oracle/raydar_oracle.c: orc_trace_bvh; NOT part of the reference, whose trace_ray is a linear
scan, cpu.rs:344-352) must return the linear scan's winner bit for bit -- it is what bench.py times as the CPU path at
100k objects.  Also orc_render_region against the corresponding crop of orc_render."""
import numpy as np
import pytest

import synth_scenes as ss


def u32(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _rays(scene, rng, n):
    """primary rays, rays from object surfaces in random directions, axis-parallel and grazing rays"""
    rays = []
    g = scene.geom
    for _ in range(n):
        k = rng.integers(0, 4)
        i = rng.integers(0, len(g))
        if k == 0:
            o = scene.cam_pos + rng.normal(0, 0.01, 3)
            d = g[rng.integers(0, len(g)), :3] - o + rng.normal(0, 0.3, 3)
        elif k == 1:
            o = g[i, :3] + rng.normal(0, 1, 3) * g[i, 3]
            d = rng.normal(0, 1, 3)
        elif k == 2:
            o = g[i, :3] + rng.normal(0, 2, 3)
            d = np.zeros(3); d[rng.integers(0, 3)] = rng.choice([-1.0, 1.0])           # zero components: linear-scan path
        else:
            j = rng.integers(0, len(g))
            o = g[i, :3] + np.array([0, g[i, 3] * 0.5, 0])
            d = g[j, :3] + np.array([0, g[j, 3] * (1 if scene.kind[j] == 0 else 0.5), 0]) - o   # grazes the top of j
        rays.append((o.astype(np.float32), d.astype(np.float32)))
    return rays


@pytest.mark.parametrize("which", ["benchmark", "config5", "config4_5k"])
def test_bvh_winner_equals_linear_scan(orc, benchmark_scene, which):
    scene = {"benchmark": lambda: benchmark_scene, "config5": lambda: ss.config5(64, 36),
             "config4_5k": lambda: ss.config4(5000, 64, 36)}[which]()
    bvh = orc.build_bvh(scene)
    rng = np.random.default_rng(5)
    n_hit = 0
    for o, d in _rays(scene, rng, 3000):
        want_i, want_t = orc.trace(scene, o, d)
        got_i, got_t, tests = bvh.trace(o, d)
        assert got_i == want_i, (o, d)
        if want_i >= 0:
            n_hit += 1
            assert u32(np.float32(got_t)) == u32(np.float32(want_t))
    assert n_hit > 500


def test_bvh_ties_lowest_index_wins(orc):
    """coincident primitives: min_by_key keeps the FIRST minimum (cpu.rs:349)"""
    base = ss.config5(32, 18)
    kind = np.concatenate([base.kind[:40], base.kind[:40], base.kind[:40]])
    geom = np.concatenate([base.geom[:40], base.geom[:40], base.geom[:40]])
    mat = np.concatenate([base.material[:40]] * 3)
    scene = orc.Scene(32, 18, base.inv_proj, base.inv_view, base.cam_pos, base.world_kind, base.world_a, base.world_b, kind, geom, mat)
    bvh = orc.build_bvh(scene)
    rng = np.random.default_rng(9)
    for o, d in _rays(scene, rng, 1500):
        want_i, want_t = orc.trace(scene, o, d)
        got_i, got_t, _ = bvh.trace(o, d)
        assert got_i == want_i and (want_i < 0 or want_i < 40 or scene.kind[want_i] == scene.kind[want_i % 40])


def test_render_region_and_bvh_render(orc, benchmark_scene):
    scene = benchmark_scene.with_resolution(96, 54)
    full = orc.render(scene, 3, 0, 3, 12, n_threads=2)
    reg, st, _ = orc.render_region(scene, 3, 0, 3, 12, 20, 10, 40, 30, n_threads=2, want_stats=True)
    assert np.array_equal(u32(reg), u32(full[10:40, 20:60]))
    assert st.samples == 40 * 30 * 3
    bvh = orc.build_bvh(scene)
    reg_b, st_b, _ = orc.render_region(scene, 3, 0, 3, 12, 20, 10, 40, 30, n_threads=2, want_stats=True, bvh=bvh)
    assert np.array_equal(u32(reg_b), u32(reg))
    assert st_b.trace_calls == st.trace_calls and st_b.primitive_tests < st.primitive_tests / 4
    c4 = ss.config4(3000, 48, 27)
    want = orc.render(c4, 1, 0, 2, 12, n_threads=2)
    got, _, _ = orc.render_region(c4, 1, 0, 2, 12, 0, 0, 48, 27, n_threads=2, bvh=orc.build_bvh(c4))
    assert np.array_equal(u32(got), u32(want))

"""Generates tests/golden/oracle_golden.npz from the C oracle (oracle/raydar_oracle.c).

The reference (Rust, non-deterministic RNG, no tests) offers no golden vectors, so these are ORACLE outputs:
they pin the oracle against drift and give the GPU tests fixed vectors that travel to the box.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SEED = 0x5EED0001
PATH_PIXELS = [(0, 0), (10, 7), (100, 50), (213, 119), (150, 80), (57, 101), (199, 3), (120, 60)]


def digest(a: np.ndarray) -> np.ndarray:
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8).copy()


def compute(orc, default_scene, benchmark_scene) -> dict:
    out = {}
    for name, scene in (("default", default_scene), ("benchmark1080", benchmark_scene.with_resolution(1920, 1080))):
        ids, ts = orc.first_hit(scene)
        out[f"{name}_first_hit_ids_sha256"] = digest(ids)
        out[f"{name}_first_hit_t_sha256"] = digest(ts)
        out[f"{name}_first_hit_ids_sub"] = ids[::16, ::16].copy()
        out[f"{name}_first_hit_t_sub"] = ts[::16, ::16].copy()
    for name, scene, spp in (("default", default_scene.with_resolution(214, 120), 8),
                             ("benchmark", benchmark_scene.with_resolution(160, 90), 4)):
        acc = orc.render(scene, SEED, 0, spp, 12, n_threads=orc.max_threads())
        out[f"{name}_accum"] = acc
        out[f"{name}_rgba8"] = orc.resolve(acc, spp)
        rows = []
        for (x, y) in PATH_PIXELS:
            x, y = x % scene.width, y % scene.height
            for sample in (0, 1, 17):
                steps, rgba = orc.trace_path(scene, x, y, sample, SEED, 12)
                for b, s in enumerate(steps):
                    rows.append([x, y, sample, b, s.object, s.lobe, s.front_face, s.t, *s.position, *s.normal,
                                 *s.origin, *s.direction, *s.attenuation, *s.light])
        rows = np.array(rows, np.float64)
        out[f"{name}_paths_int"] = rows[:, :7].astype(np.int32)
        out[f"{name}_paths_f32"] = rows[:, 7:].astype(np.float32)
    return out


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    from oracle import orc
    d = orc.load_rscn(os.path.join(ROOT, "scenes", "default.rscn"))
    b = orc.load_rscn(os.path.join(ROOT, "scenes", "benchmark.rscn"))
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **compute(orc, d, b))
    print("wrote", os.path.join(HERE, "oracle_golden.npz"))

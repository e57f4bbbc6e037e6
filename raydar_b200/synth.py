"""Seeded synthetic scenes of BASELINE.json configs 4 and 5, as plain numpy arrays and as .rscn files.

config4(n): n random Spheres/Cubes with mixed materials over a 200 x 200 field + one ground cube (BVH stress).
config5():  8x8x8 jittered lattice of glass/metal primitives inside a closed diffuse box (divergence stress).

The generator is numpy-only.  `load(cfg)` writes the scene in the serde_json shape of `Scene` (scene/mod.rs:13-18) and
reads it back through the product's own loader (rdr_scene_load_rscn), whose set_resolution recomputes the four camera
matrices (Camera::update_matrices, camera.rs:210-231); tests/synth_scenes.py wraps the same arrays for the oracle."""
from __future__ import annotations

import json
import os
import tempfile

import numpy as np

SPHERE, CUBE = 0, 1
MASK = (1 << 64) - 1
SKY_TOP, SKY_BOTTOM = (0.53, 0.8, 0.92), (1.0, 1.0, 1.0)


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & MASK

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
        return z ^ (z >> 31)

    def u(self, lo=0.0, hi=1.0):
        return lo + (hi - lo) * ((self.next() >> 11) * (1.0 / (1 << 53)))


def _material(rng):
    #        albedo(3) rough metal emission(3) estr trans ior
    albedo = [rng.u(0.2, 0.95) for _ in range(3)]
    p = rng.u()
    if p < 0.40:   return albedo + [rng.u(0.6, 1.0), 0.0, 0, 0, 0, 0.0, 0.0, 1.5]                       # diffuse
    if p < 0.65:   return albedo + [rng.u(0.0, 0.5), 1.0, 0, 0, 0, 0.0, 0.0, 1.5]                       # metal
    if p < 0.85:   return albedo + [rng.u(0.0, 0.1), 0.0, 0, 0, 0, 0.0, 1.0, 1.5]                       # glass
    if p < 0.95:   return albedo + [rng.u(0.1, 0.5), 0.0, 0, 0, 0, 0.0, 0.0, 1.5]                       # glossy dielectric
    return albedo + [0.5, 0.0, rng.u(0.5, 1.0), rng.u(0.5, 1.0), rng.u(0.5, 1.0), rng.u(5.0, 30.0), 0.0, 1.5]   # emissive


def _pack(kind, geom, mat, camera):
    return {"kind": np.asarray(kind, np.uint32), "geom": np.asarray(geom, np.float32), "material": np.asarray(mat, np.float32),
            "camera": camera}


def config4(n=100_000, width=1920, height=1080, seed=0x5EED0001):
    rng = SplitMix64(seed)
    kind, geom, mat = [CUBE], [[0.0, -1000.0, 0.0, 2000.0]], [[0.5, 0.5, 0.5, 0.9, 0.0, 0, 0, 0, 0.0, 0.0, 1.5]]
    for _ in range(n):
        k = SPHERE if rng.u() < 0.5 else CUBE
        c = [rng.u(-100, 100), rng.u(0.5, 40.0), rng.u(-100, 100)]
        size = rng.u(0.2, 1.0) if k == SPHERE else rng.u(0.4, 2.0)
        kind.append(k); geom.append(c + [size]); mat.append(_material(rng))
    cam = dict(position=[-120.0, 60.0, -120.0], target=[0.0, 5.0, 0.0], up=[0.0, 1.0, 0.0], width=width, height=height,
               fov=40.0, near=0.01, far=1000.0)
    return _pack(kind, geom, mat, cam)


def config5(width=1920, height=1080, seed=0x5EED0002):
    rng = SplitMix64(seed)
    kind, geom, mat = [CUBE], [[0.0, 0.0, 0.0, 60.0]], [[0.7, 0.7, 0.7, 0.9, 0.0, 0, 0, 0, 0.0, 0.0, 1.5]]
    for i in range(8):
        for j in range(8):
            for k in range(8):
                c = [(i - 3.5) * 2.5 + rng.u(-0.4, 0.4), (j - 3.5) * 2.5 + rng.u(-0.4, 0.4), (k - 3.5) * 2.5 + rng.u(-0.4, 0.4)]
                p = rng.u()
                albedo = [rng.u(0.6, 0.98) for _ in range(3)]
                if p < 0.60:
                    kind.append(SPHERE); geom.append(c + [rng.u(0.5, 0.9)])
                    mat.append(albedo + [rng.u(0.0, 0.05), 0.0, 0, 0, 0, 0.0, 1.0, rng.u(1.3, 1.8)])            # glass sphere
                elif p < 0.95:
                    sph = rng.u() < 0.5
                    kind.append(SPHERE if sph else CUBE); geom.append(c + [rng.u(0.5, 0.9) if sph else rng.u(0.8, 1.5)])
                    mat.append(albedo + [rng.u(0.0, 0.3), 1.0, 0, 0, 0, 0.0, 0.0, 1.5])                          # metal
                else:
                    kind.append(SPHERE); geom.append(c + [rng.u(0.4, 0.7)])
                    mat.append(albedo + [0.5, 0.0, rng.u(0.5, 1.0), rng.u(0.5, 1.0), rng.u(0.5, 1.0), rng.u(5.0, 30.0), 0.0, 1.5])
    cam = dict(position=[-24.0, 6.0, -26.0], target=[0.0, 0.0, 0.0], up=[0.0, 1.0, 0.0], width=width, height=height,
               fov=50.0, near=0.01, far=1000.0)
    return _pack(kind, geom, mat, cam)


def write_rscn(cfg, path, matrices=None):
    """serde_json shape of `Scene` (compact).  matrices: dict view/proj/inv_view/inv_proj of 16 column-major floats each;
    None writes identities (the loader's set_resolution recomputes them from position / target / up / projection)."""
    v = lambda a: {"x": float(a[0]), "y": float(a[1]), "z": float(a[2])}
    m4 = lambda m: {c: {r: float(m[ci * 4 + ri]) for ri, r in enumerate("xyzw")} for ci, c in enumerate("xyzw")}
    ident = np.eye(4, dtype=np.float32).reshape(16)
    mats = matrices or {k: ident for k in ("view", "proj", "inv_view", "inv_proj")}
    cam = cfg["camera"]
    doc = {"camera": {"position": v(cam["position"]), "target": v(cam["target"]), "up": v(cam["up"]),
                      "resolution_x": int(cam["width"]), "resolution_y": int(cam["height"]),
                      "projection": {"Perspective": {"fov": float(cam["fov"])}},
                      "near_clip": float(cam["near"]), "far_clip": float(cam["far"]),
                      "view_matrix": m4(mats["view"]), "proj_matrix": m4(mats["proj"]),
                      "inverse_view_matrix": m4(mats["inv_view"]), "inverse_proj_matrix": m4(mats["inv_proj"])},
           "world": {"SkyColor": {"top_color": v(SKY_TOP), "bottom_color": v(SKY_BOTTOM)}},
           "objects": []}
    for k, g, m in zip(cfg["kind"], cfg["geom"], cfg["material"]):
        geo = {"Sphere": {"center": v(g), "radius": float(g[3])}} if k == SPHERE else {"Cube": {"center": v(g), "side_length": float(g[3])}}
        doc["objects"].append({"geometry": geo, "material": {
            "albedo": v(m[0:3]), "roughness": float(m[3]), "metallic": float(m[4]), "emission_color": v(m[5:8]),
            "emission_strength": float(m[8]), "transmission": float(m[9]), "ior": float(m[10])}})
    with open(path, "w") as f:
        json.dump(doc, f)


def load(cfg, directory=None, name="synth.rscn"):
    """The scene through the product's loader, camera matrices recomputed by its update_matrices."""
    import raydar_b200 as rb
    own = directory is None
    directory = directory or tempfile.mkdtemp(prefix="raydar_synth_")
    path = os.path.join(directory, name)
    write_rscn(cfg, path)
    try:
        scene = rb.Scene.load(path)
    finally:
        if own:
            os.remove(path); os.rmdir(directory)
    return scene.set_resolution(int(cfg["camera"]["width"]), int(cfg["camera"]["height"]))

"""Seeded synthetic scenes of BASELINE.json configs 4 and 5.  TEST / BENCH INPUTS.

config4(n): n random Spheres/Cubes with mixed materials over a 200 x 200 field + one ground cube (BVH stress).
config5():  8x8x8 jittered lattice of glass/metal primitives inside a closed diffuse box (divergence stress).
Both return an oracle.orc.Scene (flat arrays + camera matrices from the numpy restatement of update_matrices) and
can be written as .rscn (serde_json shape of scene/mod.rs:13-18) so that the C++ loader reads the same scene."""
import json

import numpy as np

import np_restatement as npr
from oracle import orc

MASK = (1 << 64) - 1


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


def _camera(position, target, up, width, height, fov, near=0.01, far=1000.0):
    view = npr.look_at_lh(position, target, up)
    proj = npr.perspective(fov, np.float32(width) / np.float32(height), near, far)
    return dict(position=np.asarray(position, np.float32), target=np.asarray(target, np.float32), up=np.asarray(up, np.float32),
                width=width, height=height, fov=fov, near=near, far=far, view=view, proj=proj,
                inv_view=npr.invert4(view), inv_proj=npr.invert4(proj))


def _scene(cam, kind, geom, mat):
    extra = {"position": cam["position"], "target": cam["target"], "up": cam["up"],
             "projection": {"Perspective": {"fov": cam["fov"]}}, "near_clip": cam["near"], "far_clip": cam["far"],
             "view": cam["view"], "proj": cam["proj"]}
    return orc.Scene(cam["width"], cam["height"], cam["inv_proj"], cam["inv_view"], cam["position"], orc.WORLD_SKY,
                     np.array([0.53, 0.8, 0.92], np.float32), np.array([1.0, 1.0, 1.0], np.float32),
                     np.asarray(kind, np.uint32), np.asarray(geom, np.float32), np.asarray(mat, np.float32), extra)


def _material(rng, cls=None):
    #        albedo(3) rough metal emission(3) estr trans ior
    albedo = [rng.u(0.2, 0.95) for _ in range(3)]
    p = rng.u() if cls is None else cls
    if p < 0.40:   return albedo + [rng.u(0.6, 1.0), 0.0, 0, 0, 0, 0.0, 0.0, 1.5]                       # diffuse
    if p < 0.65:   return albedo + [rng.u(0.0, 0.5), 1.0, 0, 0, 0, 0.0, 0.0, 1.5]                       # metal
    if p < 0.85:   return albedo + [rng.u(0.0, 0.1), 0.0, 0, 0, 0, 0.0, 1.0, 1.5]                       # glass
    if p < 0.95:   return albedo + [rng.u(0.1, 0.5), 0.0, 0, 0, 0, 0.0, 0.0, 1.5]                       # glossy dielectric
    return albedo + [0.5, 0.0, rng.u(0.5, 1.0), rng.u(0.5, 1.0), rng.u(0.5, 1.0), rng.u(5.0, 30.0), 0.0, 1.5]   # emissive


def config4(n=100_000, width=1920, height=1080, seed=0x5EED0001):
    rng = SplitMix64(seed)
    kind, geom, mat = [orc.CUBE], [[0.0, -1000.0, 0.0, 2000.0]], [[0.5, 0.5, 0.5, 0.9, 0.0, 0, 0, 0, 0.0, 0.0, 1.5]]
    for _ in range(n):
        k = orc.SPHERE if rng.u() < 0.5 else orc.CUBE
        c = [rng.u(-100, 100), rng.u(0.5, 40.0), rng.u(-100, 100)]
        size = rng.u(0.2, 1.0) if k == orc.SPHERE else rng.u(0.4, 2.0)
        kind.append(k); geom.append(c + [size]); mat.append(_material(rng))
    cam = _camera([-120.0, 60.0, -120.0], [0.0, 5.0, 0.0], [0.0, 1.0, 0.0], width, height, 40.0)
    return _scene(cam, kind, geom, mat)


def config5(width=1920, height=1080, seed=0x5EED0002):
    rng = SplitMix64(seed)
    kind, geom, mat = [orc.CUBE], [[0.0, 0.0, 0.0, 60.0]], [[0.7, 0.7, 0.7, 0.9, 0.0, 0, 0, 0, 0.0, 0.0, 1.5]]
    for i in range(8):
        for j in range(8):
            for k in range(8):
                c = [(i - 3.5) * 2.5 + rng.u(-0.4, 0.4), (j - 3.5) * 2.5 + rng.u(-0.4, 0.4), (k - 3.5) * 2.5 + rng.u(-0.4, 0.4)]
                p = rng.u()
                albedo = [rng.u(0.6, 0.98) for _ in range(3)]
                if p < 0.60:
                    kind.append(orc.SPHERE); geom.append(c + [rng.u(0.5, 0.9)])
                    mat.append(albedo + [rng.u(0.0, 0.05), 0.0, 0, 0, 0, 0.0, 1.0, rng.u(1.3, 1.8)])            # glass sphere
                elif p < 0.95:
                    sph = rng.u() < 0.5
                    kind.append(orc.SPHERE if sph else orc.CUBE); geom.append(c + [rng.u(0.5, 0.9) if sph else rng.u(0.8, 1.5)])
                    mat.append(albedo + [rng.u(0.0, 0.3), 1.0, 0, 0, 0, 0.0, 0.0, 1.5])                          # metal
                else:
                    kind.append(orc.SPHERE); geom.append(c + [rng.u(0.4, 0.7)])
                    mat.append(albedo + [0.5, 0.0, rng.u(0.5, 1.0), rng.u(0.5, 1.0), rng.u(0.5, 1.0), rng.u(5.0, 30.0), 0.0, 1.5])
    cam = _camera([-24.0, 6.0, -26.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0], width, height, 50.0)
    return _scene(cam, kind, geom, mat)


def write_rscn(scene, path):
    """serde_json shape of `Scene` (compact, not pretty-printed)."""
    v = lambda a: {"x": float(a[0]), "y": float(a[1]), "z": float(a[2])}
    m4 = lambda m: {c: {r: float(m[ci * 4 + ri]) for ri, r in enumerate("xyzw")} for ci, c in enumerate("xyzw")}
    e = scene.extra
    doc = {"camera": {"position": v(e["position"]), "target": v(e["target"]), "up": v(e["up"]),
                      "resolution_x": scene.width, "resolution_y": scene.height, "projection": e["projection"],
                      "near_clip": float(e["near_clip"]), "far_clip": float(e["far_clip"]),
                      "view_matrix": m4(e["view"]), "proj_matrix": m4(e["proj"]),
                      "inverse_view_matrix": m4(scene.inv_view), "inverse_proj_matrix": m4(scene.inv_proj)},
           "world": {"SkyColor": {"top_color": v(scene.world_a), "bottom_color": v(scene.world_b)}},
           "objects": []}
    for k, g, m in zip(scene.kind, scene.geom, scene.material):
        geo = {"Sphere": {"center": v(g), "radius": float(g[3])}} if k == orc.SPHERE else {"Cube": {"center": v(g), "side_length": float(g[3])}}
        doc["objects"].append({"geometry": geo, "material": {
            "albedo": v(m[0:3]), "roughness": float(m[3]), "metallic": float(m[4]), "emission_color": v(m[5:8]),
            "emission_strength": float(m[8]), "transmission": float(m[9]), "ior": float(m[10])}})
    with open(path, "w") as f:
        json.dump(doc, f)

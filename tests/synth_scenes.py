"""Seeded synthetic scenes of BASELINE.json configs 4 and 5.  TEST / BENCH INPUTS.

config4(n): n random Spheres/Cubes with mixed materials over a 200 x 200 field + one ground cube (BVH stress).
config5():  8x8x8 jittered lattice of glass/metal primitives inside a closed diffuse box (divergence stress).
Both return an oracle.orc.Scene (flat arrays + camera matrices from the numpy restatement of update_matrices) and
can be written as .rscn (serde_json shape of scene/mod.rs:13-18) so that the C++ loader reads the same scene."""
import json

import numpy as np

import np_restatement as npr
from oracle import orc

from raydar_b200 import synth as _synth


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


def _wrap(cfg):
    cam = cfg["camera"]
    c = _camera(cam["position"], cam["target"], cam["up"], cam["width"], cam["height"], cam["fov"], cam["near"], cam["far"])
    return _scene(c, cfg["kind"], cfg["geom"], cfg["material"])


def config4(n=100_000, width=1920, height=1080, seed=0x5EED0001):
    return _wrap(_synth.config4(n, width, height, seed))


def config5(width=1920, height=1080, seed=0x5EED0002):
    return _wrap(_synth.config5(width, height, seed))


def write_rscn(scene, path):
    """serde_json shape of `Scene` with the matrices of the numpy restatement of update_matrices."""
    e = scene.extra
    cfg = {"kind": scene.kind, "geom": scene.geom, "material": scene.material,
           "camera": dict(position=e["position"], target=e["target"], up=e["up"], width=scene.width, height=scene.height,
                          fov=e["projection"]["Perspective"]["fov"], near=e["near_clip"], far=e["far_clip"])}
    _synth.write_rscn(cfg, path, {"view": e["view"], "proj": e["proj"], "inv_view": scene.inv_view, "inv_proj": scene.inv_proj})

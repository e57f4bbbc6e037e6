"""ctypes binding of the parity oracle (oracle/raydar_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (raydar_b200) never does.

Also holds an independent Python reader for the reference's .rscn scene format
(serde_json of `Scene`, /root/reference/src/scene/mod.rs:13-18) so that the C++ loader in the
product can be cross-checked against it.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libraydar_oracle.so")

SPHERE, CUBE = 0, 1
WORLD_SKY, WORLD_SOLID, WORLD_TRANSPARENT = 0, 1, 2
LOBE_MISS, LOBE_DIFFUSE, LOBE_SPECULAR, LOBE_REFRACT = 0, 1, 2, 3
MAT_STRIDE = 11


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "raydar_oracle.c")
    hdr = os.path.join(_HERE, "raydar_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


class _Scene(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32),
        ("inv_proj", C.c_float * 16), ("inv_view", C.c_float * 16),
        ("cam_pos", C.c_float * 3),
        ("world_kind", C.c_uint32),
        ("world_a", C.c_float * 3), ("world_b", C.c_float * 3),
        ("n_objects", C.c_uint32),
        ("kind", C.POINTER(C.c_uint32)),
        ("geom", C.POINTER(C.c_float)),
        ("material", C.POINTER(C.c_float)),
    ]


class PathStep(C.Structure):
    _fields_ = [
        ("object", C.c_int32), ("lobe", C.c_uint32), ("front_face", C.c_uint32), ("t", C.c_float),
        ("position", C.c_float * 3), ("normal", C.c_float * 3),
        ("origin", C.c_float * 3), ("direction", C.c_float * 3),
        ("attenuation", C.c_float * 3), ("light", C.c_float * 3),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("samples", C.c_uint64), ("trace_calls", C.c_uint64), ("primitive_tests", C.c_uint64),
        ("alive_at_bounce", C.c_uint64 * 64), ("lobe_count", C.c_uint64 * 4), ("exhausted", C.c_uint64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp, u32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
        sp = C.POINTER(_Scene)
        L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.orc_rng_block.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u32p]
        L.orc_u01.argtypes = [C.c_uint32]; L.orc_u01.restype = C.c_float
        L.orc_range_pm1.argtypes = [C.c_uint32]; L.orc_range_pm1.restype = C.c_float
        L.orc_random_in_unit_sphere.argtypes = [u32p, fp]
        L.orc_camera_ray.argtypes = [sp, C.c_uint32, C.c_uint32, fp, fp]
        L.orc_hit_sphere.argtypes = [fp, fp, fp, C.c_float, fp]; L.orc_hit_sphere.restype = C.c_int
        L.orc_hit_cube.argtypes = [fp, fp, fp, C.c_float, fp]; L.orc_hit_cube.restype = C.c_int
        L.orc_trace.argtypes = [sp, fp, fp, fp]; L.orc_trace.restype = C.c_int
        L.orc_closest_hit.argtypes = [sp, C.c_int, fp, fp, C.c_float, fp, fp, u32p]
        L.orc_reflect.argtypes = [fp, fp, fp]
        L.orc_refract.argtypes = [fp, fp, C.c_float, fp]
        L.orc_can_refract.argtypes = [fp, fp, C.c_float]; L.orc_can_refract.restype = C.c_int
        L.orc_world_sample.argtypes = [sp, fp, fp]
        L.orc_hit_sphere_batch.argtypes = [C.c_uint32, fp, fp, fp, i32p]
        L.orc_hit_cube_batch.argtypes = [C.c_uint32, fp, fp, fp, i32p]
        L.orc_first_hit.argtypes = [sp, i32p, fp, C.c_int]
        L.orc_trace_path.argtypes = [sp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32,
                                     C.POINTER(PathStep), fp]
        L.orc_trace_path.restype = C.c_uint32
        L.orc_render.argtypes = [sp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                 fp, C.c_int, C.POINTER(Stats)]
        L.orc_bvh_build.argtypes = [sp]; L.orc_bvh_build.restype = C.c_void_p
        L.orc_bvh_free.argtypes = [C.c_void_p]; L.orc_bvh_free.restype = None
        L.orc_trace_bvh.argtypes = [sp, C.c_void_p, fp, fp, fp, C.POINTER(C.c_uint64)]; L.orc_trace_bvh.restype = C.c_int
        L.orc_render_region.argtypes = [sp, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                        C.c_uint32, C.c_uint32, fp, C.c_int, C.POINTER(Stats)]
        L.orc_resolve.argtypes = [fp, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint8)]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


@dataclass
class Scene:
    """Flat scene: the arrays both the oracle and the CUDA C-ABI consume."""
    width: int
    height: int
    inv_proj: np.ndarray          # 16 f32, column-major as stored in .rscn
    inv_view: np.ndarray
    cam_pos: np.ndarray           # 3 f32
    world_kind: int
    world_a: np.ndarray
    world_b: np.ndarray
    kind: np.ndarray              # n u32
    geom: np.ndarray              # n x 4 f32
    material: np.ndarray          # n x 11 f32
    extra: dict = field(default_factory=dict)

    @property
    def n_objects(self) -> int:
        return int(self.kind.shape[0])

    def with_resolution(self, width: int, height: int) -> "Scene":
        """Same stored matrices at another resolution (valid when the aspect ratio is unchanged,
        as for benchmark.rscn 3840x2160 -> 1920x1080)."""
        import copy
        s = copy.copy(self)
        s.width, s.height = int(width), int(height)
        return s

    def c_struct(self) -> _Scene:
        self.kind = np.ascontiguousarray(self.kind, dtype=np.uint32)
        self.geom = np.ascontiguousarray(self.geom, dtype=np.float32).reshape(-1, 4)
        self.material = np.ascontiguousarray(self.material, dtype=np.float32).reshape(-1, MAT_STRIDE)
        s = _Scene()
        s.width, s.height = self.width, self.height
        s.inv_proj = (C.c_float * 16)(*self.inv_proj.astype(np.float32).tolist())
        s.inv_view = (C.c_float * 16)(*self.inv_view.astype(np.float32).tolist())
        s.cam_pos = _f3(self.cam_pos)
        s.world_kind = self.world_kind
        s.world_a = _f3(self.world_a)
        s.world_b = _f3(self.world_b)
        s.n_objects = self.n_objects
        s.kind = self.kind.ctypes.data_as(C.POINTER(C.c_uint32))
        s.geom = _fp(self.geom)
        s.material = _fp(self.material)
        return s


def _vec3(d):
    return np.array([d["x"], d["y"], d["z"]], dtype=np.float32)


def _mat4_cols(d):
    # cgmath Matrix4 serialises as {x: col0, y: col1, z: col2, w: col3}, each column {x,y,z,w}
    return np.array([[d[c][r] for r in "xyzw"] for c in "xyzw"], dtype=np.float32).reshape(16)


def scene_from_json(doc: dict) -> Scene:
    cam = doc["camera"]
    world = doc["world"]
    if world == "Transparent":
        wk, wa, wb = WORLD_TRANSPARENT, np.zeros(3, np.float32), np.zeros(3, np.float32)
    elif "SkyColor" in world:
        wk = WORLD_SKY
        wa, wb = _vec3(world["SkyColor"]["top_color"]), _vec3(world["SkyColor"]["bottom_color"])
    else:
        wk, wa, wb = WORLD_SOLID, _vec3(world["SolidColor"]), np.zeros(3, np.float32)
    n = len(doc["objects"])
    kind = np.zeros(n, np.uint32)
    geom = np.zeros((n, 4), np.float32)
    mat = np.zeros((n, MAT_STRIDE), np.float32)
    for i, o in enumerate(doc["objects"]):
        g = o["geometry"]
        if "Sphere" in g:
            kind[i] = SPHERE
            geom[i, :3] = _vec3(g["Sphere"]["center"]); geom[i, 3] = g["Sphere"]["radius"]
        else:
            kind[i] = CUBE
            geom[i, :3] = _vec3(g["Cube"]["center"]); geom[i, 3] = g["Cube"]["side_length"]
        m = o["material"]
        mat[i, 0:3] = _vec3(m["albedo"]); mat[i, 3] = m["roughness"]; mat[i, 4] = m["metallic"]
        mat[i, 5:8] = _vec3(m["emission_color"]); mat[i, 8] = m["emission_strength"]
        mat[i, 9] = m["transmission"]; mat[i, 10] = m["ior"]
    extra = {
        "position": _vec3(cam["position"]), "target": _vec3(cam["target"]), "up": _vec3(cam["up"]),
        "projection": cam["projection"], "near_clip": cam["near_clip"], "far_clip": cam["far_clip"],
        "view": _mat4_cols(cam["view_matrix"]), "proj": _mat4_cols(cam["proj_matrix"]),
    }
    return Scene(int(cam["resolution_x"]), int(cam["resolution_y"]),
                 _mat4_cols(cam["inverse_proj_matrix"]), _mat4_cols(cam["inverse_view_matrix"]),
                 _vec3(cam["position"]), wk, wa, wb, kind, geom, mat, extra)


def load_rscn(path: str) -> Scene:
    with open(path) as f:
        return scene_from_json(json.load(f))


# ---- oracle calls --------------------------------------------------------------------------------

def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)


def rng_block(seed, pixel, sample, bounce, block):
    o = (C.c_uint32 * 4)()
    lib().orc_rng_block(seed, pixel, sample, bounce, block, o)
    return list(o)


def camera_ray(scene: Scene, x: int, y: int):
    s = scene.c_struct(); o = (C.c_float * 3)(); d = (C.c_float * 3)()
    lib().orc_camera_ray(C.byref(s), x, y, o, d)
    return np.array(o, np.float32), np.array(d, np.float32)


def hit_sphere_batch(rays: np.ndarray, spheres: np.ndarray):
    rays = np.ascontiguousarray(rays, np.float32); spheres = np.ascontiguousarray(spheres, np.float32)
    n = rays.shape[0]; t = np.zeros(n, np.float32); h = np.zeros(n, np.int32)
    lib().orc_hit_sphere_batch(n, _fp(rays), _fp(spheres), _fp(t), h.ctypes.data_as(C.POINTER(C.c_int32)))
    return h, t


def hit_cube_batch(rays: np.ndarray, cubes: np.ndarray):
    rays = np.ascontiguousarray(rays, np.float32); cubes = np.ascontiguousarray(cubes, np.float32)
    n = rays.shape[0]; t = np.zeros(n, np.float32); h = np.zeros(n, np.int32)
    lib().orc_hit_cube_batch(n, _fp(rays), _fp(cubes), _fp(t), h.ctypes.data_as(C.POINTER(C.c_int32)))
    return h, t


def trace(scene: Scene, o, d):
    s = scene.c_struct(); t = C.c_float(0)
    idx = lib().orc_trace(C.byref(s), _f3(o), _f3(d), C.byref(t))
    return idx, np.float32(t.value)


def closest_hit(scene: Scene, obj: int, o, d, t):
    s = scene.c_struct(); p = (C.c_float * 3)(); n = (C.c_float * 3)(); ff = C.c_uint32(0)
    lib().orc_closest_hit(C.byref(s), obj, _f3(o), _f3(d), float(t), p, n, C.byref(ff))
    return np.array(p, np.float32), np.array(n, np.float32), int(ff.value)


def first_hit(scene: Scene, n_threads: int = 0):
    s = scene.c_struct()
    ids = np.zeros(scene.width * scene.height, np.int32); ts = np.zeros(scene.width * scene.height, np.float32)
    lib().orc_first_hit(C.byref(s), ids.ctypes.data_as(C.POINTER(C.c_int32)), _fp(ts),
                        n_threads or lib().orc_max_threads())
    return ids.reshape(scene.height, scene.width), ts.reshape(scene.height, scene.width)


def trace_path(scene: Scene, x: int, y: int, sample: int, seed: int, max_bounces: int):
    s = scene.c_struct(); steps = (PathStep * max(1, max_bounces))(); rgba = (C.c_float * 4)()
    n = lib().orc_trace_path(C.byref(s), x, y, sample, seed, max_bounces, steps, rgba)
    return [steps[i] for i in range(n)], np.array(rgba, np.float32)


def render(scene: Scene, seed: int, sample_begin: int, sample_end: int, max_bounces: int,
           accum: np.ndarray | None = None, n_threads: int = 1, row_begin: int = 0, row_end: int | None = None,
           want_stats: bool = False):
    s = scene.c_struct()
    if accum is None:
        accum = np.zeros((scene.height, scene.width, 4), np.float32)
    st = Stats()
    lib().orc_render(C.byref(s), seed, sample_begin, sample_end, max_bounces, row_begin,
                     scene.height if row_end is None else row_end, _fp(accum), n_threads,
                     C.byref(st) if want_stats else None)
    return (accum, st) if want_stats else accum


class Bvh:
    """The oracle's own BVH over a scene (not in the reference; see raydar_oracle.h).  Keeps the scene arrays alive."""

    def __init__(self, scene: Scene):
        self.scene = scene
        self._s = scene.c_struct()
        self._h = lib().orc_bvh_build(C.byref(self._s))

    def trace(self, o, d):
        t = C.c_float(0); tests = C.c_uint64(0)
        idx = lib().orc_trace_bvh(C.byref(self._s), self._h, _f3(o), _f3(d), C.byref(t), C.byref(tests))
        return idx, np.float32(t.value), int(tests.value)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_bvh_free(self._h)
            self._h = None


def build_bvh(scene: Scene) -> Bvh:
    return Bvh(scene)


def render_region(scene: Scene, seed: int, sample_begin: int, sample_end: int, max_bounces: int,
                  x0: int, y0: int, w: int, h: int, n_threads: int = 1, want_stats: bool = False, bvh: Bvh | None = None):
    """orc.render for a pixel rectangle; returns (accum[h, w, 4], stats, seconds)."""
    import time
    s = scene.c_struct()
    accum = np.zeros((h, w, 4), np.float32)
    st = Stats()
    t0 = time.perf_counter()
    lib().orc_render_region(C.byref(s), bvh._h if bvh is not None else None, seed, sample_begin, sample_end, max_bounces,
                            x0, y0, w, h, _fp(accum), n_threads, C.byref(st) if want_stats else None)
    return accum, st, time.perf_counter() - t0


def resolve(accum: np.ndarray, sample_count: int) -> np.ndarray:
    accum = np.ascontiguousarray(accum, np.float32)
    out = np.zeros(accum.shape, np.uint8)
    lib().orc_resolve(_fp(accum), accum.size // 4, sample_count, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def max_threads() -> int:
    return lib().orc_max_threads()

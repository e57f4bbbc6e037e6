"""raydar_b200 -- Python (ctypes) binding of libraydar_cuda.so, the B200 CUDA backend for Raydar's
per-pixel path-tracing sample loop.

The binding mirrors the reference's `Renderer` trait (/root/reference/src/renderer/mod.rs:25-35) the way
a Rust `impl Renderer for CudaRenderer` would sit on the same C ABI (include/raydar_cuda.h):

    r = Renderer(RendererConfig(max_sample_count=64, max_bounces=12))      # CpuRenderer::new(config)
    image = r.render_frame(scene)                                           # -> HxWx4 uint8 (RgbaImage)
    r.new_frame(scene); img = r.render_sample(scene)                        # -> image or None

There is no CPU fallback: the library must be built (`python -m raydar_b200.build`) and a CUDA device
must be present for any compute call; otherwise RaydarError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RAYDAR_CUDA_LIB: another build of the same library (kernel experiments: scripts/gpu_variants.sh); still no fallback
LIB_PATH = os.environ.get("RAYDAR_CUDA_LIB") or os.path.join(_HERE, "libraydar_cuda.so")

SPHERE, CUBE = 0, 1
WORLD_SKY, WORLD_SOLID, WORLD_TRANSPARENT = 0, 1, 2
ACCEL_AUTO, ACCEL_BRUTE, ACCEL_BVH, ACCEL_CLUSTER, ACCEL_COOP, ACCEL_FUSED, ACCEL_BVH_COOP = 0, 1, 2, 3, 4, 5, 6
PARTITION_SAMPLES, PARTITION_STRIPES = 0, 1
KAT_REFLECT, KAT_REFRACT, KAT_CAN_REFRACT, KAT_WORLD_SAMPLE, KAT_CLOSEST_HIT, KAT_QUANTISE, KAT_RAND_FLOATS = range(7)
COMBINE_AUTO, COMBINE_PEER, COMBINE_NCCL = 0, 1, 2
IPC_HANDLE_BYTES = 128
MAT_STRIDE = 11
OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_IO, ERR_PARSE, ERR_NCCL, ERR_NOMEM = range(8)

# every symbol include/raydar_cuda.h declares
EXPORTS = [
    "rdr_create", "rdr_destroy", "rdr_last_error", "rdr_new_frame", "rdr_render_sample", "rdr_render_frame",
    "rdr_profiler", "rdr_sample_count", "rdr_max_sample_count", "rdr_max_bounces", "rdr_set_max_sample_count",
    "rdr_set_max_bounces", "rdr_reset_frame", "rdr_set_seed", "rdr_set_sample_offset", "rdr_set_row_stripes", "rdr_set_partition", "rdr_set_accel", "rdr_render_samples",
    "rdr_resolve", "rdr_read_accum", "rdr_accum_device_ptr", "rdr_stream", "rdr_synchronize", "rdr_launch_count",
    "rdr_scene_device_bytes",
    "rdr_create_multi", "rdr_first_hit", "rdr_trace_path", "rdr_kat_hit_sphere", "rdr_kat_hit_cube",
    "rdr_kat_trace", "rdr_kat_camera_rays", "rdr_kat_rng", "rdr_scene_load_rscn", "rdr_scene_default",
    "rdr_scene_set_resolution", "rdr_scene_override_resolution", "rdr_scene_flat", "rdr_scene_free",
    "rdr_write_png", "rdr_version",
    "rdr_set_combine", "rdr_combine_in_use", "rdr_alloc_host_image", "rdr_free_host_image",
    "rdr_kat_vec", "rdr_finish_frame", "rdr_write_accum", "rdr_ipc_export", "rdr_peer_attach", "rdr_peer_combine", "rdr_peer_detach", "rdr_read_image",
]


class RaydarError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"raydar_cuda error {status}: {message}")
        self.status = status
        self.message = message


class RdrConfig(C.Structure):
    _fields_ = [("max_sample_count", C.c_uint32), ("max_bounces", C.c_uint32)]


class RdrSceneFlat(C.Structure):
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


class RdrProfiler(C.Structure):
    _fields_ = [
        ("frame_ns", C.c_uint64), ("sample_ns", C.c_uint64), ("prepare_ns", C.c_uint64), ("render_ns", C.c_uint64),
        ("has_frame", C.c_uint32), ("has_sample", C.c_uint32), ("has_prepare", C.c_uint32), ("has_render", C.c_uint32),
        ("device_render_ms", C.c_double),
    ]


class RdrPathStep(C.Structure):
    _fields_ = [
        ("object", C.c_int32), ("lobe", C.c_uint32), ("front_face", C.c_uint32), ("t", C.c_float),
        ("position", C.c_float * 3), ("normal", C.c_float * 3),
        ("origin", C.c_float * 3), ("direction", C.c_float * 3),
        ("attenuation", C.c_float * 3), ("light", C.c_float * 3),
    ]


_lib = None


def load_library():
    """Loads libraydar_cuda.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RaydarError(ERR_CUDA, f"{LIB_PATH} is missing: build it with `python -m raydar_b200.build` "
                                    "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u8p, u32p, i32p, fp = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_float)
    sfp = C.POINTER(RdrSceneFlat)
    L.rdr_version.restype = C.c_char_p
    L.rdr_last_error.argtypes = [vp]; L.rdr_last_error.restype = C.c_char_p
    L.rdr_create.argtypes = [C.POINTER(RdrConfig), C.c_int, C.POINTER(vp)]
    L.rdr_create_multi.argtypes = [C.POINTER(RdrConfig), C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]
    L.rdr_destroy.argtypes = [vp]; L.rdr_destroy.restype = None
    L.rdr_new_frame.argtypes = [vp, sfp]
    L.rdr_render_sample.argtypes = [vp, u8p, C.POINTER(C.c_int)]
    L.rdr_render_frame.argtypes = [vp, sfp, u8p]
    L.rdr_finish_frame.argtypes = [vp, u8p]
    L.rdr_profiler.argtypes = [vp, C.POINTER(RdrProfiler)]
    for name in ("rdr_sample_count", "rdr_max_sample_count", "rdr_max_bounces"):
        getattr(L, name).argtypes = [vp]; getattr(L, name).restype = C.c_uint32
    L.rdr_set_max_sample_count.argtypes = [vp, C.c_uint32]
    L.rdr_set_max_bounces.argtypes = [vp, C.c_uint32]
    L.rdr_set_seed.argtypes = [vp, C.c_uint64]
    L.rdr_set_sample_offset.argtypes = [vp, C.c_uint32]
    L.rdr_set_accel.argtypes = [vp, C.c_int]
    L.rdr_set_row_stripes.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32]
    L.rdr_set_partition.argtypes = [vp, C.c_int, C.c_uint32]
    L.rdr_debug_set_cull.argtypes = [vp, C.c_int]
    L.rdr_set_combine.argtypes = [vp, C.c_int]
    L.rdr_combine_in_use.argtypes = [vp]
    L.rdr_alloc_host_image.argtypes = [C.c_size_t, C.POINTER(u8p)]
    L.rdr_free_host_image.argtypes = [u8p]; L.rdr_free_host_image.restype = None
    L.rdr_ipc_export.argtypes = [vp, C.c_void_p]
    L.rdr_peer_attach.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_void_p]
    L.rdr_peer_combine.argtypes = [vp, C.c_uint32]
    L.rdr_peer_detach.argtypes = [vp]
    L.rdr_read_image.argtypes = [vp, u8p]
    L.rdr_render_samples.argtypes = [vp, C.c_uint32]
    L.rdr_reset_frame.argtypes = [vp]
    L.rdr_resolve.argtypes = [vp, C.c_uint32, u8p]
    L.rdr_read_accum.argtypes = [vp, fp]
    L.rdr_write_accum.argtypes = [vp, fp, C.c_uint32]
    L.rdr_accum_device_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.rdr_stream.argtypes = [vp, C.POINTER(vp)]
    L.rdr_synchronize.argtypes = [vp]
    L.rdr_launch_count.argtypes = [vp]; L.rdr_launch_count.restype = C.c_uint64
    L.rdr_scene_device_bytes.argtypes = [vp]; L.rdr_scene_device_bytes.restype = C.c_uint64
    L.rdr_first_hit.argtypes = [vp, i32p, fp]
    L.rdr_trace_path.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(RdrPathStep), C.c_uint32, u32p, fp]
    L.rdr_kat_hit_sphere.argtypes = [vp, C.c_uint32, fp, fp, fp, i32p]
    L.rdr_kat_hit_cube.argtypes = [vp, C.c_uint32, fp, fp, fp, i32p]
    L.rdr_kat_trace.argtypes = [vp, C.c_uint32, fp, i32p, fp]
    L.rdr_kat_camera_rays.argtypes = [vp, C.c_uint32, u32p, fp]
    L.rdr_kat_vec.argtypes = [vp, C.c_int, C.c_uint32, fp, fp]
    L.rdr_kat_rng.argtypes = [vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u32p]
    L.rdr_scene_load_rscn.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.rdr_scene_default.argtypes = [C.POINTER(vp)]
    L.rdr_scene_set_resolution.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.rdr_scene_override_resolution.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.rdr_scene_flat.argtypes = [vp, sfp]
    L.rdr_scene_free.argtypes = [vp]; L.rdr_scene_free.restype = None
    L.rdr_debug_scene_matrices.argtypes = [vp, fp]
    L.rdr_write_png.argtypes = [C.c_char_p, u8p, C.c_uint32, C.c_uint32]
    _lib = L
    return L


def _check(status: int, handle=None):
    if status != OK:
        msg = load_library().rdr_last_error(handle)
        raise RaydarError(status, msg.decode() if msg else "")


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def flat_from_arrays(width, height, inv_proj, inv_view, cam_pos, world_kind, world_a, world_b, kind, geom, material):
    """Builds an RdrSceneFlat over numpy arrays (kept alive on the returned struct)."""
    kind = np.ascontiguousarray(kind, np.uint32)
    geom = np.ascontiguousarray(geom, np.float32).reshape(-1, 4)
    material = np.ascontiguousarray(material, np.float32).reshape(-1, MAT_STRIDE)
    f = RdrSceneFlat()
    f.width, f.height = int(width), int(height)
    f.inv_proj = (C.c_float * 16)(*np.asarray(inv_proj, np.float32).reshape(16).tolist())
    f.inv_view = (C.c_float * 16)(*np.asarray(inv_view, np.float32).reshape(16).tolist())
    f.cam_pos = (C.c_float * 3)(*np.asarray(cam_pos, np.float32).tolist())
    f.world_kind = int(world_kind)
    f.world_a = (C.c_float * 3)(*np.asarray(world_a, np.float32).tolist())
    f.world_b = (C.c_float * 3)(*np.asarray(world_b, np.float32).tolist())
    f.n_objects = int(kind.shape[0])
    f.kind = kind.ctypes.data_as(C.POINTER(C.c_uint32))
    f.geom = _fp(geom)
    f.material = _fp(material)
    f._keep = (kind, geom, material)
    return f


class Scene:
    """Host-side scene (scene/mod.rs:13-18) owned by the C++ loader behind the C ABI."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def load(cls, path: str) -> "Scene":
        h = C.c_void_p()
        _check(load_library().rdr_scene_load_rscn(os.fsencode(path), C.byref(h)))
        return cls(h)

    @classmethod
    def default(cls) -> "Scene":
        h = C.c_void_p()
        _check(load_library().rdr_scene_default(C.byref(h)))
        return cls(h)

    def set_resolution(self, width: int, height: int) -> "Scene":
        """Camera::set_resolution_x/y: recomputes the four matrices (camera.rs:141-157, 210-231)."""
        _check(load_library().rdr_scene_set_resolution(self._h, width, height))
        return self

    def override_resolution(self, width: int, height: int) -> "Scene":
        """Changes the resolution but keeps the stored matrices (same aspect ratio)."""
        _check(load_library().rdr_scene_override_resolution(self._h, width, height))
        return self

    def flat(self) -> RdrSceneFlat:
        f = RdrSceneFlat()
        _check(load_library().rdr_scene_flat(self._h, C.byref(f)))
        f._keep = self
        return f

    def matrices(self) -> np.ndarray:
        out = np.zeros(64, np.float32)
        _check(load_library().rdr_debug_scene_matrices(self._h, _fp(out)))
        return out.reshape(4, 16)

    @property
    def width(self): return self.flat().width

    @property
    def height(self): return self.flat().height

    @property
    def n_objects(self): return self.flat().n_objects

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.rdr_scene_free(self._h)
            self._h = None


@dataclass
class RendererConfig:
    """RendererConfig, renderer/mod.rs:11-23."""
    max_sample_count: int = 1024
    max_bounces: int = 12


def _as_flat(scene) -> RdrSceneFlat:
    if isinstance(scene, RdrSceneFlat):
        return scene
    if isinstance(scene, Scene):
        return scene.flat()
    if hasattr(scene, "inv_proj") and hasattr(scene, "geom"):      # duck-typed flat scene (e.g. oracle.orc.Scene)
        return flat_from_arrays(scene.width, scene.height, scene.inv_proj, scene.inv_view, scene.cam_pos,
                                scene.world_kind, scene.world_a, scene.world_b, scene.kind, scene.geom, scene.material)
    raise TypeError(f"cannot use {type(scene)!r} as a scene")


class Renderer:
    """`impl Renderer for CudaRenderer` over the C ABI (renderer/mod.rs:25-35)."""

    def __init__(self, config: RendererConfig | None = None, device: int = 0, devices: list[int] | None = None):
        L = load_library()
        config = config or RendererConfig()
        cfg = RdrConfig(config.max_sample_count, config.max_bounces)
        self._h = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            _check(L.rdr_create_multi(C.byref(cfg), len(devices), arr, C.byref(self._h)))
        else:
            _check(L.rdr_create(C.byref(cfg), device, C.byref(self._h)))
        self._L = L
        self._shape = None

    # -- trait methods -------------------------------------------------------------------------
    def new_frame(self, scene) -> None:
        f = _as_flat(scene)
        _check(self._L.rdr_new_frame(self._h, C.byref(f)), self._h)
        self._shape = (f.height, f.width)

    def render_sample(self, scene=None, out=None):
        """Returns the resolved image after one more sample, or None once sample_count >= max_sample_count.
        Like the reference (cpu.rs:142-158) the scene argument is only used for its resolution: the frame
        was snapshotted by new_frame.  out: optional (H, W, 4) uint8 array to receive the image (an interactive
        caller reuses one buffer, ideally pinned, instead of allocating 8 MB per call)."""
        if self._shape is None:
            if scene is None:
                raise RaydarError(ERR_INVALID, "render_sample before new_frame")
            self.new_frame(scene)
        img = out if out is not None else np.empty((*self._shape, 4), np.uint8)
        if img.shape != (*self._shape, 4) or img.dtype != np.uint8 or not img.flags["C_CONTIGUOUS"]:
            raise RaydarError(ERR_INVALID, "out must be a C-contiguous (H, W, 4) uint8 array")
        produced = C.c_int(0)
        _check(self._L.rdr_render_sample(self._h, img.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(produced)), self._h)
        return img if produced.value else None

    def render_frame(self, scene, out=None) -> np.ndarray:
        """out: optional (H, W, 4) uint8 array to receive the image, e.g. a pinned one from alloc_host_image()."""
        f = _as_flat(scene)
        img = out if out is not None else np.empty((f.height, f.width, 4), np.uint8)
        if img.shape != (f.height, f.width, 4) or img.dtype != np.uint8 or not img.flags["C_CONTIGUOUS"]:
            raise RaydarError(ERR_INVALID, "out must be a C-contiguous (H, W, 4) uint8 array")
        _check(self._L.rdr_render_frame(self._h, C.byref(f), img.ctypes.data_as(C.POINTER(C.c_uint8))), self._h)
        self._shape = (f.height, f.width)
        return img

    def finish_frame(self, out=None) -> np.ndarray:
        """Every sample the current frame has left, then the image (render_frame without its new_frame)."""
        img = out if out is not None else np.empty((*self._shape, 4), np.uint8)
        _check(self._L.rdr_finish_frame(self._h, img.ctypes.data_as(C.POINTER(C.c_uint8))), self._h)
        return img

    def profiler(self) -> RdrProfiler:
        p = RdrProfiler()
        _check(self._L.rdr_profiler(self._h, C.byref(p)), self._h)
        return p

    def sample_count(self) -> int: return self._L.rdr_sample_count(self._h)
    def max_sample_count(self) -> int: return self._L.rdr_max_sample_count(self._h)
    def max_bounces(self) -> int: return self._L.rdr_max_bounces(self._h)
    def set_max_sample_count(self, count: int) -> None: _check(self._L.rdr_set_max_sample_count(self._h, count), self._h)
    def set_max_bounces(self, bounces: int) -> None: _check(self._L.rdr_set_max_bounces(self._h, bounces), self._h)

    # -- extensions ------------------------------------------------------------------------------
    def set_seed(self, seed: int) -> None: _check(self._L.rdr_set_seed(self._h, seed), self._h)
    def set_sample_offset(self, first: int) -> None: _check(self._L.rdr_set_sample_offset(self._h, first), self._h)
    def set_accel(self, accel: int) -> None: _check(self._L.rdr_set_accel(self._h, accel), self._h)

    def set_row_stripes(self, stripe_rows: int, index: int, count: int) -> None:
        """Render only the row stripes s with s % count == index (image-tile sharding; count <= 1: whole image)."""
        _check(self._L.rdr_set_row_stripes(self._h, stripe_rows, index, count), self._h)

    def set_partition(self, partition: int, stripe_rows: int = 0) -> None:
        """Multi-GPU handle only: PARTITION_SAMPLES (sample ranges) or PARTITION_STRIPES (round-robin row stripes)."""
        _check(self._L.rdr_set_partition(self._h, partition, stripe_rows), self._h)

    def set_combine(self, combine: int) -> None:
        """Multi-GPU handle only: COMBINE_PEER (fused reduce + resolve over NVLink peer memory), COMBINE_NCCL, COMBINE_AUTO."""
        _check(self._L.rdr_set_combine(self._h, combine), self._h)

    def combine_in_use(self) -> int: return self._L.rdr_combine_in_use(self._h)

    # one process per GPU: the fused combine over CUDA IPC (include/raydar_cuda.h)
    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        _check(self._L.rdr_ipc_export(self._h, buf), self._h)
        return buf.raw

    def peer_attach(self, rank: int, world: int, handles: bytes) -> None:
        assert len(handles) == world * IPC_HANDLE_BYTES
        _check(self._L.rdr_peer_attach(self._h, rank, world, handles), self._h)

    def peer_combine(self, divisor: int) -> None: _check(self._L.rdr_peer_combine(self._h, divisor), self._h)
    def peer_detach(self) -> None: _check(self._L.rdr_peer_detach(self._h), self._h)

    def read_image(self, out=None) -> np.ndarray:
        img = out if out is not None else np.empty((*self._shape, 4), np.uint8)
        _check(self._L.rdr_read_image(self._h, img.ctypes.data_as(C.POINTER(C.c_uint8))), self._h)
        return img

    def debug_set_cull(self, enabled: bool) -> None: _check(self._L.rdr_debug_set_cull(self._h, int(enabled)), self._h)
    def reset_frame(self) -> None: _check(self._L.rdr_reset_frame(self._h), self._h)
    def render_samples(self, n: int) -> None: _check(self._L.rdr_render_samples(self._h, n), self._h)
    def synchronize(self) -> None: _check(self._L.rdr_synchronize(self._h), self._h)
    def launch_count(self) -> int: return self._L.rdr_launch_count(self._h)
    def scene_device_bytes(self) -> int: return self._L.rdr_scene_device_bytes(self._h)

    def resolve(self, divisor: int = 0, out=None) -> np.ndarray:
        img = out if out is not None else np.empty((*self._shape, 4), np.uint8)
        _check(self._L.rdr_resolve(self._h, divisor, img.ctypes.data_as(C.POINTER(C.c_uint8))), self._h)
        return img

    def read_accum(self) -> np.ndarray:
        acc = np.empty((*self._shape, 4), np.float32)
        _check(self._L.rdr_read_accum(self._h, _fp(acc)), self._h)
        return acc

    def write_accum(self, accum: np.ndarray, sample_count: int) -> None:
        """Restores the frame's accumulation state (resume a progressive render from a saved accumulator)."""
        accum = np.ascontiguousarray(accum, np.float32)
        assert accum.shape == (*self._shape, 4)
        _check(self._L.rdr_write_accum(self._h, _fp(accum), sample_count), self._h)

    def accum_device_ptr(self) -> tuple[int, int]:
        p = C.c_void_p(); n = C.c_size_t()
        _check(self._L.rdr_accum_device_ptr(self._h, C.byref(p), C.byref(n)), self._h)
        return int(p.value), int(n.value)

    def stream(self) -> int:
        p = C.c_void_p()
        _check(self._L.rdr_stream(self._h, C.byref(p)), self._h)
        return int(p.value or 0)

    # -- parity / debug --------------------------------------------------------------------------
    def first_hit(self):
        ids = np.empty(self._shape, np.int32); t = np.empty(self._shape, np.float32)
        _check(self._L.rdr_first_hit(self._h, ids.ctypes.data_as(C.POINTER(C.c_int32)), _fp(t)), self._h)
        return ids, t

    def trace_path(self, x: int, y: int, sample: int, capacity: int = 128):
        steps = (RdrPathStep * capacity)(); n = C.c_uint32(0); rgba = np.zeros(4, np.float32)
        _check(self._L.rdr_trace_path(self._h, x, y, sample, steps, capacity, C.byref(n), _fp(rgba)), self._h)
        return [steps[i] for i in range(n.value)], rgba

    def kat_hit(self, sphere: bool, rays: np.ndarray, prims: np.ndarray):
        rays = np.ascontiguousarray(rays, np.float32); prims = np.ascontiguousarray(prims, np.float32)
        n = rays.shape[0]; t = np.zeros(n, np.float32); hit = np.zeros(n, np.int32)
        fn = self._L.rdr_kat_hit_sphere if sphere else self._L.rdr_kat_hit_cube
        _check(fn(self._h, n, _fp(rays), _fp(prims), _fp(t), hit.ctypes.data_as(C.POINTER(C.c_int32))), self._h)
        return hit, t

    def kat_trace(self, rays: np.ndarray):
        rays = np.ascontiguousarray(rays, np.float32)
        n = rays.shape[0]; t = np.zeros(n, np.float32); ids = np.zeros(n, np.int32)
        _check(self._L.rdr_kat_trace(self._h, n, _fp(rays), ids.ctypes.data_as(C.POINTER(C.c_int32)), _fp(t)), self._h)
        return ids, t

    def kat_camera_rays(self, xy: np.ndarray) -> np.ndarray:
        xy = np.ascontiguousarray(xy, np.uint32)
        n = xy.shape[0]; rays = np.zeros((n, 6), np.float32)
        _check(self._L.rdr_kat_camera_rays(self._h, n, xy.ctypes.data_as(C.POINTER(C.c_uint32)), _fp(rays)), self._h)
        return rays

    def kat_vec(self, op: int, records: np.ndarray) -> np.ndarray:
        """Device execution of a shading helper (KAT_*): records (n, 12) f32 -> (n, 8) f32."""
        records = np.ascontiguousarray(records, np.float32).reshape(-1, 12)
        out = np.zeros((records.shape[0], 8), np.float32)
        _check(self._L.rdr_kat_vec(self._h, op, records.shape[0], _fp(records), _fp(out)), self._h)
        return out

    def kat_rng(self, seed: int, pixel: int, sample: int, bounce: int, block: int):
        out = (C.c_uint32 * 4)()
        _check(self._L.rdr_kat_rng(self._h, seed, pixel, sample, bounce, block, out), self._h)
        return list(out)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.rdr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostImage:
    """A pinned (H, W, 4) uint8 image from rdr_alloc_host_image: the GPUs write it directly.  Keep the object alive while
    `.array` is in use; close() (or garbage collection) frees it."""

    def __init__(self, height: int, width: int):
        self._L = load_library()
        self._p = C.POINTER(C.c_uint8)()
        _check(self._L.rdr_alloc_host_image(height * width * 4, C.byref(self._p)))
        self.array = np.ctypeslib.as_array(self._p, shape=(height, width, 4))

    def close(self) -> None:
        if getattr(self, "_p", None):
            self.array = None
            self._L.rdr_free_host_image(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def write_png(path: str, rgba8: np.ndarray) -> None:
    rgba8 = np.ascontiguousarray(rgba8, np.uint8)
    h, w = rgba8.shape[:2]
    _check(load_library().rdr_write_png(os.fsencode(path), rgba8.ctypes.data_as(C.POINTER(C.c_uint8)), w, h))


def version() -> str:
    return load_library().rdr_version().decode()

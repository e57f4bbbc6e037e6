"""ctypes binding of tests/hostsim/libhostsim.so: the product's __host__ __device__ per-lane code
(raydar_b200/csrc/rdr_core.cuh, rdr_trace.cuh, rdr_pack.h) compiled for the CPU.  TEST INFRASTRUCTURE:
lets the `-m "not gpu"` suite check the logic the CUDA kernels execute against the oracle."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import raydar_b200 as rb

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SRC = os.path.join(_HERE, "hostsim", "hostsim.cpp")
_LIB = os.path.join(_HERE, "hostsim", "libhostsim.so")


class TraceStats(C.Structure):
    _fields_ = [("traces", C.c_uint64), ("sphere_exact", C.c_uint64), ("cube_exact", C.c_uint64), ("degenerate", C.c_uint64),
                ("nodes_visited", C.c_uint64), ("entries_hit", C.c_uint64)]


def build(variant: str = "", defines=()):
    """libhostsim.so, or libhostsim_<variant>.so compiled with extra -D switches (the prepared kernel variants of
    rdr_fused.cuh, so that their logic is checked on the CPU before they ever run on a GPU)."""
    csrc = os.path.join(_ROOT, "raydar_b200", "csrc")
    out = _LIB if not variant else _LIB.replace("libhostsim.so", f"libhostsim_{variant}.so")
    deps = [_SRC, os.path.join(_HERE, "hostsim", "warp_emu.h")] + [os.path.join(csrc, f) for f in os.listdir(csrc)] + \
           [os.path.join(_ROOT, "include", "raydar_cuda.h")]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    cmd = ["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-O2", "-std=c++17", "-ffp-contract=off",
           "-fno-fast-math", "-mfma", "-fopenmp", "-shared", "-fPIC", "-Wno-unknown-pragmas",
           "-I", os.path.join(_ROOT, "include"), "-I", csrc, "-I", os.path.join(_HERE, "hostsim"),
           *[f"-D{d}" for d in defines], _SRC, "-o", out]
    subprocess.run(cmd, check=True, capture_output=True)
    return out


def trace_fused_variant(variant, defines, scene, rays):
    """trace_fused of a build with extra -D switches (only this entry point of the variant library is used)."""
    L = C.CDLL(build(variant, defines))
    L.hs_trace_fused.argtypes = [C.POINTER(rb.RdrSceneFlat), C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_float)]
    f = rb._as_flat(scene)
    rays = np.ascontiguousarray(rays, np.float32)
    n = rays.shape[0]; ids = np.zeros(n, np.int32); ts = np.zeros(n, np.float32)
    _ok(L.hs_trace_fused(C.byref(f), n, _fp(rays), _ip(ids), _fp(ts)))
    return ids, ts


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        fp, i32p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)
        sfp, stp = C.POINTER(rb.RdrSceneFlat), C.POINTER(TraceStats)
        L.hs_first_hit.argtypes = [sfp, C.c_int, i32p, fp, stp]
        L.hs_trace.argtypes = [sfp, C.c_int, C.c_uint32, fp, i32p, fp, stp]
        L.hs_render.argtypes = [sfp, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, fp, stp]
        L.hs_trace_path.argtypes = [sfp, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                    C.POINTER(rb.RdrPathStep), C.c_uint32, u32p, fp]
        L.hs_resolve.argtypes = [fp, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint8)]
        L.hs_sphere_cull_batch.argtypes = [C.c_uint32, fp, fp, C.c_float, C.c_float, i32p, i32p]
        L.hs_cube_cull_batch.argtypes = [C.c_uint32, fp, fp, C.c_float, C.c_float, fp, i32p, i32p]
        L.hs_exact_batch.argtypes = [C.c_int, C.c_uint32, fp, fp, fp, i32p]
        L.hs_rng_block.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u32p]
        L.hs_scene_consts.argtypes = [sfp, fp, fp, fp]
        L.hs_bvh_info.argtypes = [sfp, u32p, u32p, u32p]
        L.hs_pack_mode.argtypes = [sfp, C.c_int, u32p, u32p]
        L.hs_bvh2_info.argtypes = [sfp, u32p, u32p, u32p]
        L.hs_bvh2_warp_sim.argtypes = [sfp, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
        L.hs_bvh2_perray_sim.argtypes = [sfp, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
        L.hs_cluster_info.argtypes = [sfp, u32p, u32p]
        L.hs_trace_fused.argtypes = [sfp, C.c_uint32, fp, i32p, fp]
        L.hs_fused_info.argtypes = [sfp, u32p]
        L.hs_stripe_pixels.argtypes = [C.c_uint32] * 5 + [u32p]
        L.hs_stripe_pixels.restype = C.c_uint32
        L.hs_render_stripes.argtypes = [sfp, C.c_uint64] + [C.c_uint32] * 6 + [fp]
        _lib = L
    return _lib


def _fp(a): return a.ctypes.data_as(C.POINTER(C.c_float))
def _ip(a): return a.ctypes.data_as(C.POINTER(C.c_int32))


def _ok(st):
    if st != 0:
        raise RuntimeError(f"hostsim status {st}")


def first_hit(scene, use_cull=True):
    f = rb._as_flat(scene)
    ids = np.zeros((f.height, f.width), np.int32); ts = np.zeros((f.height, f.width), np.float32)
    st = TraceStats()
    _ok(lib().hs_first_hit(C.byref(f), int(use_cull), _ip(ids), _fp(ts), C.byref(st)))
    return ids, ts, st


def trace(scene, rays, use_cull=True):
    f = rb._as_flat(scene)
    rays = np.ascontiguousarray(rays, np.float32)
    n = rays.shape[0]; ids = np.zeros(n, np.int32); ts = np.zeros(n, np.float32); st = TraceStats()
    _ok(lib().hs_trace(C.byref(f), int(use_cull), n, _fp(rays), _ip(ids), _fp(ts), C.byref(st)))
    return ids, ts, st


def render(scene, seed, sample_begin, n_samples, max_bounces, accum=None, use_cull=True):
    f = rb._as_flat(scene)
    if accum is None:
        accum = np.zeros((f.height, f.width, 4), np.float32)
    st = TraceStats()
    _ok(lib().hs_render(C.byref(f), int(use_cull), seed, sample_begin, n_samples, max_bounces, _fp(accum), C.byref(st)))
    return accum, st


def trace_fused(scene, rays):
    """Nearest hits by the product's warp-cooperative fused scan, run on the CPU under the warp emulator."""
    f = rb._as_flat(scene)
    rays = np.ascontiguousarray(rays, np.float32)
    n = rays.shape[0]; ids = np.zeros(n, np.int32); ts = np.zeros(n, np.float32)
    _ok(lib().hs_trace_fused(C.byref(f), n, _fp(rays), _ip(ids), _fp(ts)))
    return ids, ts


def render_fused_emu(scene, seed, sample_begin, n_samples, max_bounces, cold=True, stripes=(0, 0, 1), order=None,
                     chunk_samples=0, prior_samples=0, accum=None, variant=None):
    """The render kernel's sample loop + fused scan under the warp emulator (one CTA of 4 warps sharing the pixel counter).
    variant = (name, defines): a hostsim build with extra -D switches (e.g. RDR_CHUNKED=1)."""
    L = lib() if variant is None else C.CDLL(build(*variant))
    L.hs_render_fused_emu.argtypes = [C.POINTER(rb.RdrSceneFlat), C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                      C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32), C.c_uint32, C.c_uint32,
                                      C.c_uint32, C.POINTER(C.c_float)]
    f = rb._as_flat(scene)
    if accum is None:
        accum = np.zeros((f.height, f.width, 4), np.float32)
    ord_arr = np.ascontiguousarray(order if order is not None else [], np.int32)
    _ok(L.hs_render_fused_emu(C.byref(f), seed, sample_begin, n_samples, max_bounces, int(cold), stripes[0], stripes[1], stripes[2],
                              ord_arr.ctypes.data_as(C.POINTER(C.c_int32)), len(ord_arr), chunk_samples, prior_samples, _fp(accum)))
    return accum


def trace_bvh2_emu(scene, rays):
    """Nearest hits by the product's warp-cooperative hierarchy traversal (trace_bvh2), run under the warp emulator."""
    L = lib()
    L.hs_trace_bvh2.argtypes = [C.POINTER(rb.RdrSceneFlat), C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_float)]
    f = rb._as_flat(scene)
    rays = np.ascontiguousarray(rays, np.float32)
    n = rays.shape[0]; ids = np.zeros(n, np.int32); ts = np.zeros(n, np.float32)
    _ok(L.hs_trace_bvh2(C.byref(f), n, _fp(rays), _ip(ids), _fp(ts)))
    return ids, ts


def trace_coop_emu(scene, rays):
    """Nearest hits by the product's warp-cooperative cluster scan (trace_cluster_coop), run under the warp emulator."""
    L = lib()
    L.hs_trace_coop.argtypes = [C.POINTER(rb.RdrSceneFlat), C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_float)]
    f = rb._as_flat(scene)
    rays = np.ascontiguousarray(rays, np.float32)
    n = rays.shape[0]; ids = np.zeros(n, np.int32); ts = np.zeros(n, np.float32)
    _ok(L.hs_trace_coop(C.byref(f), n, _fp(rays), _ip(ids), _fp(ts)))
    return ids, ts


def fused_clusters(scene, variant=None):
    """Top-level entry of the fused clustering that holds each original object.  variant = (name, defines): a hostsim
    build with extra -D switches (e.g. RDR_CLUSTER_REFINE=0: the clustering before the surface-area refinement)."""
    L = C.CDLL(build(*variant)) if variant else lib()
    L.hs_fused_clusters.argtypes = [C.POINTER(rb.RdrSceneFlat), C.POINTER(C.c_int32)]
    f = rb._as_flat(scene)
    cl = np.zeros(f.n_objects, np.int32)
    _ok(L.hs_fused_clusters(C.byref(f), _ip(cl)))
    return cl


def fused_info(scene):
    f = rb._as_flat(scene)
    out = (C.c_uint32 * 7)()
    _ok(lib().hs_fused_info(C.byref(f), out))
    return dict(zip(("fused_ok", "fused_top", "fused_cap", "fused_direct", "fused_ns_direct", "fused_stage_bytes", "blob_bytes"), out))


def stripe_pixels(width, height, rows, index, count):
    """Pixel indices one row-stripe shard renders, in hand-out order (the product's stripe_pixel())."""
    L = lib()
    args = (width, height, rows, index, count)
    n = L.hs_stripe_pixels(*args, None)
    out = np.zeros(n, np.uint32)
    L.hs_stripe_pixels(*args, out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out


def render_stripes(scene, seed, sample_begin, n_samples, max_bounces, rows, index, count, accum=None):
    f = rb._as_flat(scene)
    if accum is None:
        accum = np.zeros((f.height, f.width, 4), np.float32)
    _ok(lib().hs_render_stripes(C.byref(f), seed, sample_begin, n_samples, max_bounces, rows, index, count, _fp(accum)))
    return accum


def trace_path(scene, x, y, sample, seed, max_bounces, use_cull=True, capacity=128):
    f = rb._as_flat(scene)
    steps = (rb.RdrPathStep * capacity)(); n = C.c_uint32(0); rgba = np.zeros(4, np.float32)
    _ok(lib().hs_trace_path(C.byref(f), int(use_cull), seed, x, y, sample, max_bounces, steps, capacity, C.byref(n), _fp(rgba)))
    return [steps[i] for i in range(n.value)], rgba


def resolve(accum, divisor):
    accum = np.ascontiguousarray(accum, np.float32)
    out = np.zeros(accum.shape, np.uint8)
    lib().hs_resolve(_fp(accum), accum.size // 4, divisor, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def sphere_cull(rays, spheres, q_max, origin_bound):
    rays = np.ascontiguousarray(rays, np.float32); spheres = np.ascontiguousarray(spheres, np.float32)
    n = rays.shape[0]; may = np.zeros(n, np.int32); deg = np.zeros(n, np.int32)
    lib().hs_sphere_cull_batch(n, _fp(rays), _fp(spheres), q_max, origin_bound, _ip(may), _ip(deg))
    return may, deg


def cube_cull(rays, cubes, pad, origin_bound, best):
    rays = np.ascontiguousarray(rays, np.float32); cubes = np.ascontiguousarray(cubes, np.float32)
    best = np.ascontiguousarray(best, np.float32)
    n = rays.shape[0]; may = np.zeros(n, np.int32); deg = np.zeros(n, np.int32)
    lib().hs_cube_cull_batch(n, _fp(rays), _fp(cubes), pad, origin_bound, _fp(best), _ip(may), _ip(deg))
    return may, deg


def exact(sphere, rays, prims):
    rays = np.ascontiguousarray(rays, np.float32); prims = np.ascontiguousarray(prims, np.float32)
    n = rays.shape[0]; t = np.zeros(n, np.float32); hit = np.zeros(n, np.int32)
    lib().hs_exact_batch(int(sphere), n, _fp(rays), _fp(prims), _fp(t), _ip(hit))
    return hit, t


def rng_block(seed, pixel, sample, bounce, block):
    out = (C.c_uint32 * 4)()
    lib().hs_rng_block(seed, pixel, sample, bounce, block, out)
    return list(out)


BVH = 2      # pass as use_cull to select the hierarchy (1/True = scan + cull, 0/False = scan, exact test on everything)


CLUSTER = 3  # two-level cluster scan
BVH2 = 4     # scalar walk of the pair-packed hierarchy of the warp-cooperative traversal (builder check)


def cluster_info(scene):
    f = rb._as_flat(scene)
    n = C.c_uint32(); nbytes = C.c_uint32()
    _ok(lib().hs_cluster_info(C.byref(f), C.byref(n), C.byref(nbytes)))
    return n.value, nbytes.value


def bvh_info(scene):
    f = rb._as_flat(scene)
    n = C.c_uint32(); mode = C.c_uint32(); nbytes = C.c_uint32()
    _ok(lib().hs_bvh_info(C.byref(f), C.byref(n), C.byref(mode), C.byref(nbytes)))
    return n.value, mode.value, nbytes.value


def bvh2_info(scene):
    f = rb._as_flat(scene)
    n = C.c_uint32(); root = C.c_uint32(); ok = C.c_uint32()
    _ok(lib().hs_bvh2_info(C.byref(f), C.byref(n), C.byref(root), C.byref(ok)))
    return n.value, root.value, ok.value


def bvh2_warp_sim(scene, flush_at=32, order_children=0):
    """(node visits per ray, exact tests per ray, rounds per warp, tasks per round) of the cooperative stack discipline."""
    f = rb._as_flat(scene)
    out = (C.c_double * 4)()
    _ok(lib().hs_bvh2_warp_sim(C.byref(f), flush_at, order_children, out))
    return tuple(out)


def bvh2_perray_sim(scene, flush_at=32, per_round=8):
    """(node visits per ray, exact tests per ray, rounds per warp, tasks per round, deepest per-ray stack) with one
    near-first stack per ray and `per_round` rays served per round."""
    f = rb._as_flat(scene)
    out = (C.c_double * 5)()
    _ok(lib().hs_bvh2_perray_sim(C.byref(f), flush_at, per_round, out))
    return tuple(out)


def scene_consts(scene):
    f = rb._as_flat(scene)
    q = C.c_float(); ob = C.c_float(); pad = C.c_float()
    _ok(lib().hs_scene_consts(C.byref(f), C.byref(q), C.byref(ob), C.byref(pad)))
    return q.value, ob.value, pad.value


def pack_mode(scene, accel):
    """(status, layout mode, scan top entries) of rdr_new_frame's packing decision for `accel` (host only)."""
    f = rb._as_flat(scene)
    mode = C.c_uint32(0); n_top = C.c_uint32(0)
    st = lib().hs_pack_mode(C.byref(f), accel, C.byref(mode), C.byref(n_top))
    return st, int(mode.value), int(n_top.value)

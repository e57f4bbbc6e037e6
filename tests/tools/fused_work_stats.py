#!/usr/bin/env python
"""Work counters of the fused scan on the CPU: runs the product's trace_fused under the warp emulator (an RDR_EMU_STATS
build of tests/hostsim) on warps of rays taken from real paths of benchmark.rscn -- 32 neighbouring pixels per warp, every
lane always holding the next ray of its own path, as in the kernel's sample refill -- and prints tasks, member rounds and
exact rounds per warp-trace with and without the surface-area refinement of the clustering (RDR_CLUSTER_REFINE).
TEST INFRASTRUCTURE (uses the oracle for the paths).  Usage: OMP_NUM_THREADS=1 python tests/tools/fused_work_stats.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
os.environ.setdefault("OMP_NUM_THREADS", "1")          # the counters are per thread

from oracle import orc                                # noqa: E402
import hostsim_py as hs                               # noqa: E402
import raydar_b200 as rb                              # noqa: E402

scene = orc.load_rscn(os.path.join(os.path.dirname(os.path.dirname(HERE)), "scenes", "benchmark.rscn")).with_resolution(1920, 1080)
rng = np.random.default_rng(4)
warps = []
for _ in range(150):
    x0, y0 = int(rng.integers(0, 1920 - 32)), int(rng.integers(0, 1080))
    paths = []
    for i in range(32):
        steps = orc.trace_path(scene, x0 + i, y0, int(rng.integers(0, 1000)), 5, 12)
        steps = steps[0] if isinstance(steps, tuple) else steps
        o, d = orc.camera_ray(scene, x0 + i, y0)
        paths.append([np.concatenate([o, d])] + [np.array(list(st.origin) + list(st.direction)) for st in steps if st.object >= 0])
    for it in range(6):
        warps.append([paths[i][it % len(paths[i])] for i in range(32)])
rays = np.array(warps, np.float32).reshape(-1, 6)

f = rb._as_flat(scene)
out = (C.c_ulonglong * 7)()
for env in ("0", "1"):                     # compile-time switch: one hostsim build per setting
    L = C.CDLL(hs.build(f"stats_refine{env}", ("RDR_EMU_STATS=1", f"RDR_CLUSTER_REFINE={env}")))
    L.hs_trace_fused.argtypes = [C.POINTER(rb.RdrSceneFlat), C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_float)]
    L.hs_fused_emu_stats(out, 1)
    ids = np.zeros(len(rays), np.int32); ts = np.zeros(len(rays), np.float32)
    L.hs_trace_fused(C.byref(f), len(rays), rays.ctypes.data_as(C.POINTER(C.c_float)), ids.ctypes.data_as(C.POINTER(C.c_int32)),
                     ts.ctypes.data_as(C.POINTER(C.c_float)))
    L.hs_fused_emu_stats(out, 1)
    n = out[0]
    print(f"RDR_CLUSTER_REFINE={env}: {n} warp-traces; per warp-trace: tasks {out[1] / n:.1f}, member rounds {out[2] / n:.2f}, "
          f"sphere rounds {out[3] / n:.2f} ({out[5] / n:.1f} tests), cube rounds {out[4] / n:.2f} ({out[6] / n:.1f} tests)")

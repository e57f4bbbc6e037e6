#!/usr/bin/env python
"""bench.py -- Msamples/s of the path-tracing sample loop on scenes/benchmark.rscn (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's CUDA backend
    python bench.py --impl reference [--steps K] [--warmup W]           # the CPU path (oracle port) on host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...               # one rank per GPU

A "step" is one full frame of the hot path: W x H pixels x spp samples, 12 bounces, into a fresh float
accumulator, then (N > 1) one NCCL reduce of the per-GPU accumulators and the resolve to RGBA8.
N GPUs shard by sample range with a fixed per-GPU load (weak scaling): rank g renders global samples
[g*spp, (g+1)*spp), so a step produces an image with N*spp samples per pixel.

`value` = samples of all ranks / max-over-ranks time with the scene resident in HBM.
`e2e`   = the same through the reference-facing call (rdr_render_frame: borrowed HOST scene in, HOST RGBA8 image
          out; scene packing, H2D, kernels, resolve and D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Msamples/s on benchmark.rscn 1080p, 12 bounces"
UNIT = "Msamples/s"

# algorithmic flops per sample, SURVEY.md 8(d): F = F_raygen + B * (Ns*17 + Nc*22 + F_shade)
F_SPHERE, F_CUBE, F_SHADE, F_RAYGEN = 17.0, 22.0, 200.0, 60.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default=os.path.join(ROOT, "scenes", "benchmark.rscn"))
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=1024)
    ap.add_argument("--bounces", type=int, default=12)
    ap.add_argument("--seed", type=int, default=0x5EED)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--progressive", action="store_true",
                    help="time the interactive path instead: render_sample() = one 1-spp launch + resolve + image read-back per call")
    ap.add_argument("--partition", default="samples", choices=["samples", "stripes"],
                    help="N > 1: shard by sample range (default) or by round-robin 16-row stripes (bit-identical to 1 GPU); "
                         "both keep the per-GPU load fixed: the frame has N*spp samples per pixel")
    ap.add_argument("--accel", default="auto", choices=["auto", "brute", "bvh", "cluster", "coop", "fused", "bvh2"], help="nearest-hit search of the CUDA backend")
    return ap.parse_args()


def workload_name(a):
    return f"{os.path.basename(a.scene)} {a.width}x{a.height}, {a.spp} spp, {a.bounces} bounces"


def flops_per_sample(n_spheres, n_cubes, traces_per_sample):
    return F_RAYGEN + traces_per_sample * (n_spheres * F_SPHERE + n_cubes * F_CUBE + F_SHADE)


# ---- CPU arm: the oracle port of the reference's CPU backend ---------------------------------------------
def host_cores():
    """Cores this process may run on.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to every rank, and
    the reference arm is meant to use all the host threads it can (the oracle passes the count in a num_threads clause)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_run(a, target_seconds, threads=None):
    """Times the oracle on a bounded sample of the workload: the full-resolution frame at a reduced spp
    (samples/s does not depend on spp).  Returns (Msamples/s, description, traces_per_sample, cores)."""
    from oracle import orc
    scene = orc.load_rscn(a.scene).with_resolution(a.width, a.height)
    threads = threads or host_cores()
    t0 = time.perf_counter()
    _, st = orc.render(scene, a.seed, 0, 1, a.bounces, n_threads=threads, want_stats=True)
    t1 = time.perf_counter() - t0
    spp = int(max(1, min(64, round(target_seconds / max(t1, 1e-3)))))
    t0 = time.perf_counter()
    _, st = orc.render(scene, a.seed, 0, spp, a.bounces, n_threads=threads, want_stats=True)
    dt = time.perf_counter() - t0
    samples = a.width * a.height * spp
    return samples / dt / 1e6, f"{a.width}x{a.height} x {spp} spp ({samples / 1e6:.1f} Msamples, {dt:.1f} s)", \
        st.trace_calls / max(1, st.samples), threads, scene


def cpu_single_thread(a):
    from oracle import orc
    scene = orc.load_rscn(a.scene).with_resolution(max(1, a.width // 2), max(1, a.height // 2))
    t0 = time.perf_counter()
    orc.render(scene, a.seed, 0, 1, a.bounces, n_threads=1)
    dt = time.perf_counter() - t0
    n = scene.width * scene.height
    return {"value": n / dt / 1e6, "unit": UNIT, "cores": 1, "sample": f"{scene.width}x{scene.height} x 1 spp ({dt:.1f} s)"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import orc
    scene = orc.load_rscn(a.scene).with_resolution(a.width, a.height)
    threads = host_cores()
    # size a step (a bounded sample of the frame: full resolution, reduced spp) to a few seconds
    t0 = time.perf_counter(); orc.render(scene, a.seed, 0, 1, a.bounces, n_threads=threads); t1 = time.perf_counter() - t0
    spp = int(max(1, min(16, round(4.0 / max(t1, 1e-3)))))
    for _ in range(a.warmup):
        orc.render(scene, a.seed, 0, 1, a.bounces, n_threads=threads)
    t0 = time.perf_counter()
    for k in range(a.steps):
        orc.render(scene, a.seed, k * spp, (k + 1) * spp, a.bounces, n_threads=threads)
    dt = time.perf_counter() - t0
    samples = a.width * a.height * spp * a.steps
    value = samples / dt / 1e6
    sample = f"{a.width}x{a.height} x {spp} spp per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "step": sample,
                   "note": "CPU path = C port of the reference's Rust CPU backend (oracle/), OpenMP over rows; "
                           "the reference itself is single-threaded and cannot be built here (no cargo)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---- clocks ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows, self.proc, self.idx = [], None, device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


# ---- GPU arm -----------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist
    import raydar_b200 as rb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA backend has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world

    scene = rb.Scene.load(a.scene).override_resolution(a.width, a.height)
    flat = scene.flat()
    n_pixels = a.width * a.height
    from raydar_b200 import dist as rdist
    stripes = a.partition == "stripes" and world > 1
    # samples: rank g renders spp samples of every pixel; stripes: rank g renders all N*spp samples of 1/N of the rows
    rank_spp = a.spp * world if stripes else a.spp
    r = rb.Renderer(rb.RendererConfig(rank_spp, a.bounces), device=local)
    r.set_seed(a.seed)
    r.set_accel({"auto": rb.ACCEL_AUTO, "brute": rb.ACCEL_BRUTE, "bvh": rb.ACCEL_BVH, "cluster": rb.ACCEL_CLUSTER, "coop": rb.ACCEL_COOP, "fused": rb.ACCEL_FUSED, "bvh2": rb.ACCEL_BVH_COOP}[a.accel])
    if stripes:
        r.set_row_stripes(rdist.STRIPE_ROWS, rank, world)
    else:
        r.set_sample_offset(rank * a.spp)                # weak scaling: every rank renders spp samples of its own range
    r.new_frame(flat)
    ptr, nbytes = r.accum_device_ptr()

    class _Arr:                                          # wrap the renderer's device accumulator as a torch tensor
        __cuda_array_interface__ = {"shape": (n_pixels * 4,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    accum_t = torch.as_tensor(_Arr(), device=f"cuda:{local}")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")    # > 126 MB L2
    img = np.empty((a.height, a.width, 4), np.uint8)
    device_ms = []

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        flush.zero_()                                    # L2 flush between timed iterations
        torch.cuda.synchronize()
        r.reset_frame()
        r.render_samples(rank_spp)                       # returns after the kernel's stop event
        device_ms.append(r.profiler().device_render_ms)
        if world > 1:
            dist.reduce(accum_t, dst=0, op=dist.ReduceOp.SUM)     # NCCL over NVLink, per-GPU accumulators -> rank 0
            torch.cuda.synchronize()
        if rank == 0:
            r.resolve(a.spp * world)                     # resolve kernel (+ image read-back; 8 MB, not in the kernel time)

    def step_e2e():
        flush.zero_()
        torch.cuda.synchronize()
        if world == 1:
            return r.render_frame(flat)                  # host scene in -> host RGBA8 out
        r.new_frame(flat)
        r.render_samples(rank_spp)
        dist.reduce(accum_t, dst=0, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
        return r.resolve(a.spp * world) if rank == 0 else None

    if a.progressive:
        run_progressive(a, r, flat, rank)
        if world > 1:
            dist.destroy_process_group()
        r.close()
        return

    for _ in range(a.warmup):
        step_resident()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = r.launch_count()
    device_ms.clear()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_resident()
    barrier()
    dt = time.perf_counter() - t0
    launches = r.launch_count() - launches0
    kernel_ms = float(np.mean(device_ms)) if device_ms else float("nan")

    for _ in range(1):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    barrier()
    dt_e2e = time.perf_counter() - t0
    clock_info = clocks.stop() if rank == 0 else None

    t = torch.tensor([dt, dt_e2e, kernel_ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt, dt_e2e, kernel_ms = [float(v) for v in t.tolist()]

    if rank == 0:
        samples_per_step = n_pixels * a.spp * world
        value = samples_per_step * a.steps / dt / 1e6
        e2e = samples_per_step * a.steps / dt_e2e / 1e6
        peaks, peak_src = measured_peaks()
        sm_count = torch.cuda.get_device_properties(local).multi_processor_count
        fp32_peak = sm_count * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12      # TFLOP/s
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "parallelism": (f"row-stripes({rdist.STRIPE_ROWS}) x{world}" if stripes else f"sample-range x{world}"), "accel": a.accel,
                       "l2": "256 MiB buffer written between timed iterations (L2 flush)",
                       "step": "fresh accumulator + spp samples/pixel in one kernel launch"
                               + (" + NCCL reduce to rank 0" if world > 1 else "") + " + resolve to RGBA8"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(r.scene_device_bytes()),
                    "d2h_bytes_per_step": n_pixels * 4, "ms_per_step": dt_e2e / a.steps * 1e3},
            "gpu_launches": int(launches),
            "clocks": clock_info,
        }
        cpu = None
        # mean trace_ray calls per sample, counted by the oracle; re-counted below when the CPU leg runs
        traces_per_sample = {("benchmark.rscn", 12): 3.1028}.get((os.path.basename(a.scene), a.bounces))
        if not a.no_cpu_baseline and world == 1:
            v, sample, traces_per_sample, cores, _ = cpu_run(a, a.cpu_seconds)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
            # the reference's CPU backend is single-threaded (cpu.rs:198): the same port on ONE core, on a quarter-size frame
            cpu["single_thread"] = cpu_single_thread(a)
        n_sph = int((np.ctypeslib.as_array(flat.kind, (flat.n_objects,)) == rb.SPHERE).sum())
        n_cub = flat.n_objects - n_sph
        # scenes whose trace count is unknown (no CPU leg in this run) and scenes traversed through the BVH have no
        # brute-force-scan roofline: only the raw Msamples/s is reported for them
        have_roofline = traces_per_sample is not None and flat.n_objects <= 1024
        f_sample = flops_per_sample(n_sph, n_cub, traces_per_sample) if have_roofline else None
        achieved = (n_pixels * a.spp) / (kernel_ms * 1e-3) * f_sample / 1e12 if have_roofline else None
        traffic, traffic_src = ncu_traffic()
        headline = os.path.basename(a.scene) == "benchmark.rscn" and a.bounces == 12
        out["roofline"] = {
            "bound": "fp32", "kernel": "render_kernel (fused two-level scan)" if a.accel in ("auto", "fused") else f"render_kernel ({a.accel})",
            "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": achieved / fp32_peak if have_roofline else None, "traffic": traffic if headline and (a.width, a.height) == (1920, 1080) else None,
            "traffic_source": traffic_src.get("source"),
            # the same capture's utilisation figures for the dominant kernel (north-star: FP32 pipe, warp execution efficiency, issue slots)
            "ncu": {k: traffic_src.get(k) for k in ("fp32_pipe_active_pct", "fma_pipe_inst_pct", "alu_pipe_inst_pct", "issue_slots_busy_pct",
                                                   "warp_execution_efficiency_lanes", "achieved_occupancy_pct")} if headline else None,
            "kernel_ms": kernel_ms, "flops_per_sample": f_sample, "traces_per_sample": traces_per_sample,
            "peak_source": f"{sm_count} SMs x 128 lanes x 2 flop x sm_max_mhz from {peak_src}",
            "note": "FP32-pipe bound (no dense contraction, HBM traffic is 32 B/pixel/launch); achieved = algorithmic "
                    "flops of the reference's brute-force scan per launch / CUDA-event kernel time",
            "hbm_algorithmic_bytes_per_launch": n_pixels * 32 + int(r.scene_device_bytes()),
        }
        if cpu:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    r.close()


def run_progressive(a, r, flat, rank):
    """The editor's loop (raydar_editor.rs:95-115): new_frame, then render_sample() until it returns None; every call
    is one 1-spp kernel launch, the resolve kernel and the RGBA8 read-back.  Reports the mean latency per call."""
    import torch
    n_calls = min(a.spp, 256)
    r.set_max_sample_count(n_calls)
    img = torch.empty((a.height, a.width, 4), dtype=torch.uint8, pin_memory=True).numpy()     # one pinned image buffer, reused
    lat = []
    for rep in range(a.warmup + a.steps):
        r.new_frame(flat)
        t0 = time.perf_counter()
        calls = 0
        while r.render_sample(flat, out=img) is not None:
            calls += 1
        dt = time.perf_counter() - t0
        if rep >= a.warmup:
            lat.append(dt / max(1, calls))
    if rank == 0:
        ms = float(np.mean(lat)) * 1e3
        print(json.dumps({"metric": "ms per render_sample() call (1 spp + resolve + read-back)", "value": ms, "unit": "ms",
                          "higher_is_better": False, "calls_per_frame": n_calls, "steps": a.steps, "warmup": a.warmup,
                          "Msamples/s": a.width * a.height / (ms * 1e-3) / 1e6,
                          "config": {"workload": workload_name(a).replace(f"{a.spp} spp", "1 spp per call"), "accel": a.accel}}))


def ncu_traffic():
    """DRAM bytes of one render-kernel launch from the committed `ncu --set full` capture (profiles/ncu_traffic.json,
    written by scripts/ncu_summary.py).  Per launch the kernel touches every pixel's accumulator once each way, so the
    figure does not depend on spp."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, {}
    with open(p) as f:
        d = json.load(f)
    return d.get("dram_bytes_per_launch"), d


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)

#!/usr/bin/env python
"""bench.py -- Msamples/s of the path-tracing sample loop on scenes/benchmark.rscn (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's CUDA backend
    python bench.py --impl reference [--steps K] [--warmup W]           # the CPU path (oracle port) on host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...               # one rank per GPU
    python bench.py --config {2,3,4,5}                                  # another BASELINE.json config as the workload

A "step" is one full frame of the hot path: W x H pixels x spp samples, 12 bounces, into a fresh float accumulator,
combined over the GPUs and resolved to RGBA8.

N = 1: `value` = samples / device-timed step with the scene resident in HBM; `e2e` = the same through the reference-facing
call rdr_render_frame (borrowed HOST scene in, pinned HOST RGBA8 image out; scene packing, H2D, kernels, resolve and
D2H inside the timed region).

N > 1 (STRONG scaling: the frame's spp are split N ways, sample ranges): `value` = one rank per GPU, each rendering its
share, then the fused reduce + resolve kernel over the ranks' accumulators mapped through CUDA IPC (NVLink peer reads;
falls back to one NCCL reduce onto rank 0 if IPC is unavailable).  `e2e` = rank 0 calls rdr_render_frame on an
rdr_create_multi handle over all N devices (what a Rust host would call; the other ranks idle at a host barrier).
`weak` (secondary) keeps the per-GPU load fixed instead.  `parity_check`: a small frame on the multi-GPU handle against
one GPU (sample ranges: alpha exact, colour within one RGBA8 level; row stripes: bit-identical).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Msamples/s on benchmark.rscn 1080p, 12 bounces"
UNIT = "Msamples/s"

# algorithmic flops per sample, SURVEY.md 8(d): F = F_raygen + B * (Ns*17 + Nc*22 + F_shade)
F_SPHERE, F_CUBE, F_SHADE, F_RAYGEN = 17.0, 22.0, 200.0, 60.0
# mean trace_ray calls per sample counted by the oracle (re-counted whenever the CPU leg runs)
TRACES_PER_SAMPLE = {("benchmark.rscn", 12): 3.1028}

ACCELS = ["auto", "brute", "bvh", "cluster", "coop", "fused", "bvh2"]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 2 = benchmark.rscn 1080p 1024 spp (headline), 3 = 4K 4096 spp, "
                         "4 = 100k synthetic objects 256 spp, 5 = glass/metal lattice 1024 spp 32 bounces")
    ap.add_argument("--scene", default=None, help="a .rscn file instead of the config's scene")
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--spp", type=int, default=None)
    ap.add_argument("--bounces", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0x5EED)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="N = 1 headline run: skip the short runs of configs 3, 4 and 5")
    ap.add_argument("--progressive", action="store_true",
                    help="time the interactive path instead: render_sample() = one 1-spp launch + resolve + image read-back per call")
    ap.add_argument("--partition", default="samples", choices=["samples", "stripes"],
                    help="N > 1: shard by sample range (default) or by round-robin 16-row stripes (bit-identical to 1 GPU)")
    ap.add_argument("--accel", default="auto", choices=ACCELS, help="nearest-hit search of the CUDA backend")
    a = ap.parse_args()
    cfg = {2: ("benchmark.rscn", 1920, 1080, 1024, 12), 3: ("benchmark.rscn", 3840, 2160, 4096, 12),
           4: ("config4", 1920, 1080, 256, 12), 5: ("config5", 1920, 1080, 1024, 32)}[a.config]
    a.scene_name = os.path.basename(a.scene) if a.scene else (cfg[0] if cfg[0].endswith(".rscn") else cfg[0] + ".rscn (synthetic, seeded)")
    a.synthetic = None if (a.scene or cfg[0].endswith(".rscn")) else cfg[0]
    if a.scene is None and a.synthetic is None:
        a.scene = os.path.join(ROOT, "scenes", cfg[0])
    a.width = a.width or cfg[1]; a.height = a.height or cfg[2]; a.spp = a.spp or cfg[3]; a.bounces = a.bounces or cfg[4]
    return a


def workload_name(a, spp=None):
    return f"{a.scene_name} {a.width}x{a.height}, {spp or a.spp} spp, {a.bounces} bounces"


def flops_per_sample(n_spheres, n_cubes, traces_per_sample):
    return F_RAYGEN + traces_per_sample * (n_spheres * F_SPHERE + n_cubes * F_CUBE + F_SHADE)


_SYNTH_CACHE = {}


def synth_cfg(a):
    """The seeded synthetic scene's arrays (generated once per process: 100k objects take a few seconds of Python)."""
    from raydar_b200 import synth
    key = (a.synthetic, a.width, a.height)
    if key not in _SYNTH_CACHE:
        _SYNTH_CACHE[key] = synth.config4(100_000, a.width, a.height) if a.synthetic == "config4" else synth.config5(a.width, a.height)
    return _SYNTH_CACHE[key]


def load_scene(a):
    """The workload's scene through the product's loader (rdr_scene_load_rscn)."""
    import raydar_b200 as rb
    if a.synthetic:
        from raydar_b200 import synth
        return synth.load(synth_cfg(a))
    return rb.Scene.load(a.scene).override_resolution(a.width, a.height)


# ---- CPU arm: the oracle port of the reference's CPU backend ---------------------------------------------
def host_cores():
    """Cores this process may run on.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to every rank, and
    the reference arm is meant to use all the host threads it can (the oracle passes the count in a num_threads clause)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def oracle_scene(a):
    """The same scene in the oracle's container (the CPU legs only)."""
    from oracle import orc
    if a.synthetic:
        key = ("oracle", a.synthetic, a.width, a.height)
        if key not in _SYNTH_CACHE:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import synth_scenes as ss
            _SYNTH_CACHE[key] = ss._wrap(synth_cfg(a))
        return _SYNTH_CACHE[key]
    return orc.load_rscn(a.scene).with_resolution(a.width, a.height)


def cpu_run(a, target_seconds, threads=None):
    """Times the oracle on a bounded sample of the workload.  The reference's trace_ray is a linear scan over all
    objects (cpu.rs:344-352), so the sample is the full frame at a reduced spp when that fits the time target, and a
    centred crop of the frame at 1 spp otherwise (100k objects).  Returns a cpu_baseline dict + traces per sample."""
    from oracle import orc
    scene = oracle_scene(a)
    threads = threads or host_cores()
    # probe: a small centred crop at 1 spp
    probe = orc.render_region(scene, a.seed, 0, 1, a.bounces, *centre_crop(a, 64, 36), n_threads=threads, want_stats=True)
    per_sample = probe[2] / max(1, 64 * 36)
    budget = max(1.0, target_seconds / max(per_sample, 1e-9))                  # samples the time target buys
    if budget >= a.width * a.height:
        spp = int(max(1, min(64, budget // (a.width * a.height))))
        t0 = time.perf_counter()
        _, st = orc.render(scene, a.seed, 0, spp, a.bounces, n_threads=threads, want_stats=True)
        dt = time.perf_counter() - t0
        samples = a.width * a.height * spp
        desc = f"{a.width}x{a.height} x {spp} spp ({samples / 1e6:.1f} Msamples, {dt:.1f} s)"
    else:
        cw = int(max(16, min(a.width, (budget * a.width / a.height) ** 0.5)))
        ch = int(max(9, min(a.height, cw * a.height // a.width)))
        x0, y0, cw, ch = centre_crop(a, cw, ch)
        _, st, dt = orc.render_region(scene, a.seed, 0, 1, a.bounces, x0, y0, cw, ch, n_threads=threads, want_stats=True)
        samples = cw * ch
        desc = f"centred {cw}x{ch} crop of the {a.width}x{a.height} frame x 1 spp ({samples / 1e6:.3f} Msamples, {dt:.1f} s)"
    return {"value": samples / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc,
            "algorithm": "linear scan over all objects per trace_ray, as the reference (cpu.rs:344-352)"}, \
        st.trace_calls / max(1, st.samples)


def centre_crop(a, cw, ch):
    cw, ch = min(cw, a.width), min(ch, a.height)
    return (a.width - cw) // 2, (a.height - ch) // 2, cw, ch


def cpu_single_thread(a):
    from oracle import orc
    scene = oracle_scene(a)
    x0, y0, cw, ch = centre_crop(a, a.width // 2, a.height // 2)
    _, st, dt = orc.render_region(scene, a.seed, 0, 1, a.bounces, x0, y0, cw, ch, n_threads=1, want_stats=True)
    return {"value": cw * ch / dt / 1e6, "unit": UNIT, "cores": 1, "sample": f"centred {cw}x{ch} crop x 1 spp ({dt:.1f} s)"}


def cpu_bvh_run(a, target_seconds):
    """A faster CPU path the reference does not have: the oracle's own BVH (verified against its linear scan in
    tests/test_oracle_bvh.py) -- reported beside the reference-faithful linear scan for scenes where that scan is hopeless."""
    from oracle import orc
    scene = oracle_scene(a)
    threads = host_cores()
    t0 = time.perf_counter(); bvh = orc.build_bvh(scene); t_build = time.perf_counter() - t0
    x0, y0, cw, ch = centre_crop(a, 480, 270)
    _, st, dt = orc.render_region(scene, a.seed, 0, 1, a.bounces, x0, y0, cw, ch, n_threads=threads, want_stats=True, bvh=bvh)
    spp = int(max(1, min(16, round(target_seconds / max(dt, 1e-3)))))
    if spp > 1:
        _, st, dt = orc.render_region(scene, a.seed, 0, spp, a.bounces, x0, y0, cw, ch, n_threads=threads, want_stats=True, bvh=bvh)
    return {"value": cw * ch * spp / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port + oracle BVH (not in the reference)",
            "sample": f"centred {cw}x{ch} crop x {spp} spp ({dt:.1f} s; BVH build {t_build:.1f} s)"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import orc
    scene = oracle_scene(a)
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
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if a.gpus > 1 else "weak",
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


def ncu_profile():
    """Figures of the committed `ncu --set full` capture of the render kernel (profiles/ncu_traffic.json, written by
    scripts/ncu_summary.py from the capture named in its `source`): DRAM bytes per launch and the utilisation metrics."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return {}
    with open(p) as f:
        return json.load(f)


# ---- GPU arm: shared pieces ----------------------------------------------------------------------------------
def accel_id(rb, name):
    return {"auto": rb.ACCEL_AUTO, "brute": rb.ACCEL_BRUTE, "bvh": rb.ACCEL_BVH, "cluster": rb.ACCEL_CLUSTER, "coop": rb.ACCEL_COOP,
            "fused": rb.ACCEL_FUSED, "bvh2": rb.ACCEL_BVH_COOP}[name]


def scene_counts(rb, flat):
    n_sph = int((np.ctypeslib.as_array(flat.kind, (flat.n_objects,)) == rb.SPHERE).sum())
    return n_sph, flat.n_objects - n_sph


def roofline(a, rb, flat, torch, local, kernel_ms, samples_per_launch, traces_per_sample, traces_source, clocks=None):
    peaks, peak_src = measured_peaks()
    sm_count = torch.cuda.get_device_properties(local).multi_processor_count
    fp32_peak = sm_count * 128 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12      # TFLOP/s
    n_sph, n_cub = scene_counts(rb, flat)
    have = traces_per_sample is not None
    f_sample = flops_per_sample(n_sph, n_cub, traces_per_sample) if have else None
    achieved = samples_per_launch / (kernel_ms * 1e-3) * f_sample / 1e12 if have else None
    prof = ncu_profile()
    headline = a.config == 2 and a.scene_name == "benchmark.rscn" and (a.width, a.height, a.bounces) == (1920, 1080, 12)
    out = {
        "bound": "fp32", "kernel": "render_kernel (" + ("warp-cooperative hierarchy" if flat.n_objects > 1024 else "fused two-level scan") + ")"
                                   if a.accel == "auto" else f"render_kernel ({a.accel})",
        "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak if have else None,
        "frac_is": "an EFFECTIVE rate: algorithmic flops of the reference's brute-force scan (every object, every trace) "
                   "per second / FP32 peak.  The kernel culls most of those tests and reuses the primary hit, so it can "
                   "exceed 1; pipe utilisation is ncu.fp32_pipe_active_pct",
        "traffic": prof.get("dram_bytes_per_launch") if headline else None,
        "traffic_source": prof.get("source") if headline else None,
        "ncu": ({k: prof.get(k) for k in ("fp32_pipe_active_pct", "fma_pipe_inst_pct", "alu_pipe_inst_pct", "issue_slots_busy_pct",
                                          "warp_execution_efficiency_lanes", "achieved_occupancy_pct", "executed_fp32_flops_per_sample",
                                          "executed_over_algorithmic_flops")} if headline else None),
        "kernel_ms": kernel_ms, "flops_per_sample": f_sample, "traces_per_sample": traces_per_sample,
        "traces_per_sample_source": traces_source,
        "peak_source": f"{sm_count} SMs x 128 lanes x 2 flop x sm_max_mhz from {peak_src} (nominal at the measured clock; the file has no FP32 figure)",
        "note": "FP32-pipe bound (no dense contraction, HBM traffic is 52 B/pixel/launch); achieved = algorithmic "
                "flops of the reference's brute-force scan per launch / CUDA-event kernel time",
        # accumulator read + write (16 B each) and the primary-table record (20 B) per pixel per launch
        "hbm_algorithmic_bytes_per_launch": a.width * a.height * 52,
    }
    if clocks and clocks.get("sm_mhz"):
        # against the max-clock peak (above) and the peak at the SM clock sampled under load during the timed region
        out["peak_at_clock_under_load"] = sm_count * 128 * 2 * float(clocks["sm_mhz"]) * 1e6 / 1e12
        out["frac_at_clock_under_load"] = achieved / out["peak_at_clock_under_load"] if have else None
    if headline and prof.get("executed_fp32_flops_per_sample"):
        ex = prof["executed_fp32_flops_per_sample"] * samples_per_launch / (kernel_ms * 1e-3) / 1e12
        out["executed"] = {"tflops": ex, "frac_of_fp32_peak": ex / fp32_peak,
                           "note": "FP32 flops the kernel EXECUTES (ncu smsp__sass_thread_inst_executed_op_{fadd,fmul,ffma}_pred_on, FFMA = 2) "
                                   "per sample of the committed capture x this run's samples/s: the machine-level fraction of FP32 peak"}
    return out


def flush_buffers(torch, devices):
    return [torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{d}") for d in devices]      # > 126 MB L2 each


def time_progressive(a, r, flat, torch, n_calls=64, reps=3):
    """The editor's loop (raydar_editor.rs:95-115): new_frame, then render_sample() until None; every call is one 1-spp
    launch per device, the resolve / combine kernel and the RGBA8 image in pinned host memory.  Mean ms per call."""
    import raydar_b200 as rb
    old = r.max_sample_count()
    r.set_max_sample_count(n_calls)
    img = rb.HostImage(a.height, a.width)
    lat = []
    for rep in range(reps + 1):
        r.new_frame(flat)
        t0 = time.perf_counter()
        calls = 0
        while r.render_sample(flat, out=img.array) is not None:
            calls += 1
        dt = time.perf_counter() - t0
        if rep >= 1:
            lat.append(dt / max(1, calls))
    r.set_max_sample_count(old)
    img.close()
    return float(np.mean(lat)) * 1e3


# ---- GPU arm, one GPU -------------------------------------------------------------------------------------------
def bench_single(a, rb, torch, local, steps, warmup, want_e2e=True):
    """Device-timed and end-to-end throughput of one workload on one GPU."""
    scene = load_scene(a)
    flat = scene.flat()
    n_pixels = a.width * a.height
    r = rb.Renderer(rb.RendererConfig(a.spp, a.bounces), device=local)
    r.set_seed(a.seed)
    r.set_accel(accel_id(rb, a.accel))
    r.new_frame(flat)
    flush = flush_buffers(torch, [local])[0]
    img = rb.HostImage(a.height, a.width)
    device_ms = []

    def step_resident():
        flush.zero_()                                    # L2 flush between timed iterations
        torch.cuda.synchronize()
        r.reset_frame()
        r.render_samples(a.spp)                          # returns after the kernel's stop event
        device_ms.append(r.profiler().device_render_ms)
        r.resolve(a.spp, out=img.array)                  # resolve kernel + RGBA8 read-back

    def step_e2e():
        flush.zero_()
        torch.cuda.synchronize()
        return r.render_frame(flat, out=img.array)       # host scene in -> host RGBA8 out

    for _ in range(warmup):
        step_resident()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = r.launch_count()
    device_ms.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_resident()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    launches = r.launch_count() - launches0
    kernel_ms = float(np.mean(device_ms))
    dt_e2e = None
    if want_e2e:
        step_e2e()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            step_e2e()
        torch.cuda.synchronize()
        dt_e2e = time.perf_counter() - t0
    clock_info = clocks.stop()
    res = {"value": n_pixels * a.spp * steps / dt / 1e6, "ms_per_step": dt / steps * 1e3, "kernel_ms": kernel_ms,
           "e2e": n_pixels * a.spp * steps / dt_e2e / 1e6 if dt_e2e else None, "e2e_ms": dt_e2e / steps * 1e3 if dt_e2e else None,
           "launches": int(launches), "clocks": clock_info, "h2d": int(r.scene_device_bytes()), "flat": flat, "scene": scene,
           "renderer": r}
    img.close()
    return res


def run_single(a, rb, torch, local):
    res = bench_single(a, rb, torch, local, a.steps, a.warmup)
    r, flat = res["renderer"], res["flat"]
    if a.progressive:
        ms = time_progressive(a, r, flat, torch, n_calls=min(a.spp, 256), reps=a.steps)
        print(json.dumps({"metric": "ms per render_sample() call (1 spp + resolve + read-back)", "value": ms, "unit": "ms",
                          "higher_is_better": False, "steps": a.steps, "warmup": a.warmup,
                          "Msamples/s": a.width * a.height / (ms * 1e-3) / 1e6,
                          "config": {"workload": workload_name(a).replace(f"{a.spp} spp", "1 spp per call"), "accel": a.accel}}))
        r.close()
        return
    n_pixels = a.width * a.height
    out = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "parallelism": "1 GPU", "accel": a.accel,
                   "l2": "256 MiB buffer written between timed iterations (L2 flush)",
                   "step": "fresh accumulator + spp samples/pixel in one kernel launch + resolve to RGBA8 + image read-back"},
        "e2e": {"value": res["e2e"], "unit": UNIT, "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": n_pixels * 4,
                "ms_per_step": res["e2e_ms"], "call": "rdr_render_frame (host scene in, pinned host RGBA8 out)"},
        "gpu_launches": res["launches"], "clocks": res["clocks"],
    }
    traces = TRACES_PER_SAMPLE.get((a.scene_name, a.bounces))
    traces_src = "oracle count, DESIGN.md 5" if traces is not None else None
    cpu = None
    if not a.no_cpu_baseline:
        cpu, traces = cpu_run(a, a.cpu_seconds)
        traces_src = "counted by the oracle in this run's cpu_baseline leg"
        # the reference's CPU backend is single-threaded (cpu.rs:198): the same port on ONE core
        cpu["single_thread"] = cpu_single_thread(a) if a.config in (2, 3) else None
        if a.config == 4:
            cpu["with_oracle_bvh"] = cpu_bvh_run(a, 5.0)
    out["roofline"] = roofline(a, rb, flat, torch, local, res["kernel_ms"], n_pixels * a.spp, traces, traces_src, res["clocks"])
    if cpu:
        out["cpu_baseline"] = cpu
    out["render_sample_ms"] = time_progressive(a, r, flat, torch)
    r.close()
    default_headline = (a.config == 2 and a.scene_name == "benchmark.rscn" and a.accel == "auto" and
                        (a.width, a.height, a.spp, a.bounces) == (1920, 1080, 1024, 12))
    if default_headline and not a.no_other_configs:
        out["other_configs"] = other_configs(a, rb, torch, local)
    print(json.dumps(out))


def other_configs(a, rb, torch, local):
    """Short runs of BASELINE.json configs 3, 4 and 5 on this GPU, appended to the headline line: device-timed value, e2e,
    clocks, and the CPU path on a bounded sample beside each (samples/s does not depend on spp, so steps use a reduced spp)."""
    res = {}
    for cfg, spp, label in ((3, 512, "benchmark.rscn 3840x2160, 12 bounces; 512 spp per step = one GPU's share of the 4096 spp on 8"),
                            (4, 256, "100k synthetic objects 1920x1080, 256 spp, 12 bounces (hierarchy in L2)"),
                            (5, 256, "glass/metal lattice (513 objects) 1920x1080, 32 bounces; 256 spp per step")):
        try:
            b = argparse.Namespace(**vars(a))
            b.config = cfg
            name, w, h, _, bounces = {3: ("benchmark.rscn", 3840, 2160, 0, 12), 4: ("config4", 1920, 1080, 0, 12), 5: ("config5", 1920, 1080, 0, 32)}[cfg]
            b.width, b.height, b.spp, b.bounces = w, h, spp, bounces
            b.synthetic = None if name.endswith(".rscn") else name
            b.scene = os.path.join(ROOT, "scenes", name) if b.synthetic is None else None
            b.scene_name = name if b.synthetic is None else name + ".rscn (synthetic, seeded)"
            t0 = time.perf_counter()
            r1 = bench_single(b, rb, torch, local, steps=2, warmup=1)
            entry = {"workload": label, "value": r1["value"], "unit": UNIT, "kernel_ms": r1["kernel_ms"], "e2e": r1["e2e"],
                     "clocks": r1["clocks"], "n_objects": int(r1["flat"].n_objects), "steps": 2, "warmup": 1}
            r1["renderer"].close()
            if not a.no_cpu_baseline:
                cpu, traces = cpu_run(b, 6.0)
                entry["cpu_baseline"] = cpu
                entry["traces_per_sample"] = traces
                if cfg == 4:
                    entry["cpu_baseline"]["with_oracle_bvh"] = cpu_bvh_run(b, 4.0)
                elif r1["flat"].n_objects <= 1024:
                    rf = roofline(b, rb, r1["flat"], torch, local, r1["kernel_ms"], w * h * spp, traces, "oracle, this run")
                    entry["roofline_frac_effective"] = rf["frac"]
            entry["wall_s"] = time.perf_counter() - t0
            res[f"config{cfg}"] = entry
        except Exception as e:                                   # a secondary measurement must not lose the headline line
            res[f"config{cfg}"] = {"error": f"{type(e).__name__}: {e}"}
    return res


# ---- GPU arm, N GPUs (one rank per GPU) -------------------------------------------------------------------------
def run_multi(a, rb, torch, dist, world, rank, local):
    from raydar_b200 import dist as rdist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    host_group = dist.new_group(backend="gloo")              # host-side barriers: idle ranks must not spin on their GPU
    scene = load_scene(a)
    flat = scene.flat()
    n_pixels = a.width * a.height
    stripes = a.partition == "stripes"
    flush = flush_buffers(torch, [local])[0]
    token = torch.zeros(1, device=f"cuda:{local}")

    def gpu_barrier():                                       # every rank's GPU work so far is finished
        torch.cuda.synchronize()
        dist.all_reduce(token)
        torch.cuda.synchronize()

    def make_renderer(total_spp, weak):
        if stripes:
            rank_spp, first = (total_spp * world if weak else total_spp), 0
        elif weak:
            first, rank_spp = rdist.weak_sample_range(rank, total_spp)
        else:
            first, rank_spp = rdist.strong_sample_range(rank, world, total_spp)
        r = rb.Renderer(rb.RendererConfig(rank_spp, a.bounces), device=local)
        r.set_seed(a.seed)
        r.set_accel(accel_id(rb, a.accel))
        if stripes:
            r.set_row_stripes(rdist.STRIPE_ROWS, rank, world)
        else:
            r.set_sample_offset(first)
        r.new_frame(flat)
        return r, rank_spp

    def attach(r):
        """CUDA IPC: every rank maps the other ranks' accumulators and rank 0's device image.  All ranks agree on the outcome."""
        ok = 1
        try:
            handles = [None] * world
            dist.all_gather_object(handles, r.ipc_export(), group=host_group)
            r.peer_attach(rank, world, b"".join(handles))
        except Exception as e:
            ok = 0
            if rank == 0:
                print(f"bench.py: CUDA IPC attach failed ({e}); falling back to the NCCL reduce", file=sys.stderr)
        flag = torch.tensor([ok], device=f"cuda:{local}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            r.peer_detach()
            return False
        return True

    def timed(r, rank_spp, total_divisor, steps, warmup, ipc):
        accum_t = None if ipc else rdist.device_accum_tensor(r, n_pixels, local)
        device_ms = []

        def step():
            flush.zero_()                                    # L2 flush between timed iterations
            gpu_barrier()                                    # the previous step's combine has read every accumulator
            r.reset_frame()
            r.render_samples(rank_spp)                       # returns after the kernel's stop event
            device_ms.append(r.profiler().device_render_ms)
            if ipc:
                gpu_barrier()                                # every rank has rendered: the accumulators are final
                r.peer_combine(total_divisor)                # fused reduce + resolve of this rank's pixel slice over NVLink
            else:
                dist.reduce(accum_t, dst=0, op=dist.ReduceOp.SUM)
                torch.cuda.synchronize()
                if rank == 0:
                    r.resolve(total_divisor)

        for _ in range(warmup):
            step()
        launches0 = r.launch_count()
        device_ms.clear()
        gpu_barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        gpu_barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt, float(np.mean(device_ms))], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), r.launch_count() - launches0

    # ---- strong scaling: the frame's spp split over the ranks -------------------------------------------------------
    r, rank_spp = make_renderer(a.spp, weak=False)
    ipc = (not stripes) and attach(r)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    dt, kernel_ms, launches = timed(r, rank_spp, a.spp, a.steps, a.warmup, ipc)
    image_rank0 = r.read_image() if (rank == 0 and ipc) else None
    r.peer_detach()
    r.close()

    # ---- weak scaling (secondary): every rank renders the full spp of its own sample range ---------------------------
    rw, rank_spp_w = make_renderer(a.spp, weak=True)
    ipc_w = (not stripes) and attach(rw)
    dt_w, kernel_ms_w, _ = timed(rw, rank_spp_w, a.spp * world, max(2, a.steps // 2), 1, ipc_w)
    rw.peer_detach()
    rw.close()

    # ---- e2e: rank 0 drives all N devices through the product's multi-GPU handle; the other ranks idle on the host --
    e2e = None
    dist.barrier(group=host_group)
    if rank == 0:
        try:
            e2e = multi_handle_leg(a, rb, torch, flat, world, stripes, image_rank0)
        except Exception as exc:                              # the other ranks wait at the barrier below: never leave them there
            e2e = {"value": None, "ms_per_step": None, "combine": None, "parity_check": False, "render_sample_ms": None, "h2d": 0,
                   "parity_detail": {"error": f"{type(exc).__name__}: {exc}"}}
    dist.barrier(group=host_group)
    clock_info = clocks.stop() if rank == 0 else None

    if rank == 0:
        samples_per_step = n_pixels * a.spp
        out = {
            "metric": METRIC, "value": samples_per_step * a.steps / dt / 1e6, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a),
                       "parallelism": (f"row-stripes({rdist.STRIPE_ROWS}) x{world}" if stripes else f"sample-range x{world}") + f": {a.spp} spp split over {world} GPUs",
                       "accel": a.accel, "l2": "256 MiB buffer written between timed iterations (L2 flush)",
                       "combine": "fused reduce + resolve kernel over CUDA-IPC-mapped peer accumulators (NVLink), RGBA8 slices into rank 0's device image"
                                  if ipc else "one NCCL reduce(sum, f32) onto rank 0 + resolve there",
                       "step": "fresh accumulators + each rank's share of the spp in one kernel launch + combine"},
            "e2e": {"value": e2e["value"], "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"] * world, "d2h_bytes_per_step": n_pixels * 4,
                    "ms_per_step": e2e["ms_per_step"], "call": "rdr_render_frame on an rdr_create_multi handle over all devices, from rank 0 "
                                                               "(host scene in, pinned host RGBA8 out)", "combine": e2e["combine"]},
            "gpu_launches": int(launches) * world, "clocks": clock_info,
            "weak": {"value": n_pixels * a.spp * world * max(2, a.steps // 2) / dt_w / 1e6, "unit": UNIT, "spp_per_gpu": a.spp,
                     "kernel_ms": kernel_ms_w, "note": "fixed per-GPU load: the frame has N x spp samples per pixel"},
            "parity_check": e2e["parity_check"], "parity_detail": e2e["parity_detail"],
            "multi_render_sample_ms": e2e["render_sample_ms"],
        }
        traces = TRACES_PER_SAMPLE.get((a.scene_name, a.bounces))
        out["roofline"] = roofline(a, rb, flat, torch, local, kernel_ms, n_pixels * rank_spp, traces,
                                   "oracle count, DESIGN.md 5 (the CPU leg runs at N = 1 only)" if traces is not None else None, clock_info)
        print(json.dumps(out))
    dist.destroy_process_group()


def multi_handle_leg(a, rb, torch, flat, world, stripes, image_rank0):
    """Rank 0 only: rdr_render_frame through the single-process multi-GPU handle, the in-bench parity assertion and the
    progressive latency on that handle."""
    devices = list(range(world))
    n_pixels = a.width * a.height
    m = rb.Renderer(rb.RendererConfig(a.spp, a.bounces), devices=devices)
    m.set_seed(a.seed)
    m.set_accel(accel_id(rb, a.accel))
    if stripes:
        m.set_partition(rb.PARTITION_STRIPES, 16)
    flush = flush_buffers(torch, devices)
    img = rb.HostImage(a.height, a.width)

    def sync_all():
        for d in devices:
            torch.cuda.synchronize(d)

    def step():
        for f in flush:
            f.zero_()
        sync_all()
        m.render_frame(flat, out=img.array)

    step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    sync_all()
    dt = time.perf_counter() - t0
    combine = {rb.COMBINE_PEER: "peer (fused reduce + resolve over NVLink, RGBA8 slices written into the pinned host image)",
               rb.COMBINE_NCCL: "nccl (ncclReduce onto device 0 + resolve + copy)"}.get(m.combine_in_use(), "single device")
    detail = {}
    ok = True
    # the torchrun arm and the multi-GPU handle render the same frame with the same summation order: same image
    if image_rank0 is not None and not stripes:
        detail["torchrun_arm_image_equals_multi_handle_image"] = bool(np.array_equal(image_rank0, img.array))
        ok &= detail["torchrun_arm_image_equals_multi_handle_image"]
    # small frame: multi-GPU handle against one GPU
    small = flat_with_resolution(a, 640, 360)
    one = rb.Renderer(rb.RendererConfig(16, a.bounces), device=0); one.set_seed(77)
    img1 = one.render_frame(small); acc1 = one.read_accum()
    m.set_max_sample_count(16); m.set_seed(77)
    m.set_partition(rb.PARTITION_SAMPLES)
    imgs = m.render_frame(small); accs = m.read_accum()
    detail["samples_alpha_exact"] = bool(np.array_equal(accs[..., 3], acc1[..., 3]))
    detail["samples_max_level_diff"] = int(np.abs(imgs.astype(int) - img1.astype(int)).max())
    detail["samples_allclose"] = bool(np.allclose(accs, acc1, rtol=1e-5, atol=1e-5))
    m.set_partition(rb.PARTITION_STRIPES, 16)
    imgt = m.render_frame(small); acct = m.read_accum()
    detail["stripes_bit_identical"] = bool(np.array_equal(acct.view(np.uint32), acc1.view(np.uint32)) and np.array_equal(imgt, img1))
    ok &= detail["samples_alpha_exact"] and detail["samples_max_level_diff"] <= 1 and detail["samples_allclose"] and detail["stripes_bit_identical"]
    one.close()
    # progressive latency on the multi handle (row stripes: every device renders the sample for its own rows)
    m.set_seed(a.seed)
    ms = time_progressive(a, m, flat, torch)
    h2d = m.scene_device_bytes()
    m.close()
    img.close()
    return {"value": n_pixels * a.spp * a.steps / dt / 1e6, "ms_per_step": dt / a.steps * 1e3, "combine": combine,
            "parity_check": bool(ok), "parity_detail": detail, "render_sample_ms": ms, "h2d": int(h2d)}


def flat_with_resolution(a, w, h):
    import raydar_b200 as rb
    if a.synthetic:
        from raydar_b200 import synth
        cfg = dict(synth_cfg(a)); cfg["camera"] = dict(cfg["camera"], width=w, height=h)
        s = synth.load(cfg)
    else:
        s = rb.Scene.load(a.scene).override_resolution(w, h)
    return s.flat()


def run_b200(a):
    import torch
    import raydar_b200 as rb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA backend has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        run_multi(a, rb, torch, dist, world, rank, local)
    else:
        run_single(a, rb, torch, local)


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)

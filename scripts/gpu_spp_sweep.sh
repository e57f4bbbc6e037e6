#!/bin/bash
# kernel time against spp and resolution: separates the per-launch, per-pixel and per-sample costs of render_kernel
TAG=${1:-sweep}; OUT=gpurun_out; mkdir -p $OUT
python - <<'PY' 2>&1 | tee $OUT/spp_sweep_$TAG.txt
import numpy as np, raydar_b200 as rb
for (w, h) in ((1920, 1080), (3840, 2160), (960, 540)):
    scene = rb.Scene.load("scenes/benchmark.rscn").override_resolution(w, h); flat = scene.flat()
    for spp in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512):
        if w * h * spp > 3840 * 2160 * 64: continue
        r = rb.Renderer(rb.RendererConfig(spp, 12)); r.new_frame(flat)
        ms = []
        for rep in range(4):
            r.reset_frame(); r.render_samples(spp); ms.append(r.profiler().device_render_ms)
        print(f"{w}x{h} spp {spp:4d} kernel_ms {min(ms[1:]):9.4f}  ns/sample {min(ms[1:]) * 1e6 / (w * h * spp):7.4f}")
        r.close()
PY

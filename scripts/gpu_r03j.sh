#!/bin/bash
TAG=${1:-r03j}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_gpu_$TAG.log
echo "== sweep"; python - <<'PY' 2>&1 | tee $OUT/spp_sweep_$TAG.txt
import numpy as np, raydar_b200 as rb
scene = rb.Scene.load("scenes/benchmark.rscn").override_resolution(1920, 1080); flat = scene.flat()
for spp in (1, 16, 128, 256):
    r = rb.Renderer(rb.RendererConfig(spp, 12)); r.new_frame(flat)
    ms = []
    for rep in range(4):
        r.reset_frame(); r.render_samples(spp); ms.append(r.profiler().device_render_ms)
    print(f"1920x1080 spp {spp:4d} kernel_ms {min(ms[1:]):9.4f}  ns/sample {min(ms[1:]) * 1e6 / (1920 * 1080 * spp):7.4f}")
    r.close()
PY
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/bench_$TAG.json | cut -c1-600

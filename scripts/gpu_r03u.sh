#!/bin/bash
TAG=r03u; OUT=gpurun_out; mkdir -p $OUT
for N in 8 4; do
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_n${N}_$TAG.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=$N', d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['parity_check'], d['multi_render_sample_ms'])"
done
echo "== combine cmp"; timeout 300 python - <<'PY' 2>&1 | tail -3 | tee $OUT/combine_cmp_$TAG.txt
import time, numpy as np, torch, raydar_b200 as rb
scene = rb.Scene.load("scenes/benchmark.rscn").override_resolution(1920, 1080); flat = scene.flat()
img = rb.HostImage(1080, 1920)
for name, comb in (("peer", rb.COMBINE_PEER), ("nccl", rb.COMBINE_NCCL)):
    m = rb.Renderer(rb.RendererConfig(1024, 12), devices=list(range(8))); m.set_combine(comb)
    m.render_frame(flat, out=img.array)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); m.render_frame(flat, out=img.array); ts.append(time.perf_counter() - t0)
    k = m.profiler().device_render_ms
    print(name, "frame ms", 1e3 * float(np.median(ts)), "kernel ms (max over devices)", k, "tail ms", 1e3 * float(np.median(ts)) - k)
    m.close()
PY

#!/bin/bash
# 8-GPU visit with the landed kernel: multi-GPU + CLI tests, strong-scaling bench at N = 8 / 4 / 2 (headline) and N = 8 config 3,
# PEER vs NCCL combine on the multi-GPU handle, launch list of one multi-handle frame
TAG=r03n; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt 2>&1; nproc >> $OUT/gpu_$TAG.txt
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_cli.py -m gpu -q 2>&1 | tail -4 | tee $OUT/pytest_multi_$TAG.log
for N in 8 4 2; do
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_n${N}_$TAG.json | cut -c1-400
done
echo "== bench N=8 config 3"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --config 3 --steps 3 --warmup 2 2>&1 | tail -1 | tee $OUT/bench_n8_config3_$TAG.json | cut -c1-400
echo "== bench N=1 same box"; timeout 600 python bench.py --no-cpu-baseline --no-other-configs --steps 5 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_n1_$TAG.json | cut -c1-300
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
echo "== ncu launch list of one multi-handle frame (8 GPUs, 1024 spp)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_multi8_$TAG.csv python - <<'PY' > $OUT/ncu_multi8_$TAG.log 2>&1
import raydar_b200 as rb
scene = rb.Scene.load("scenes/benchmark.rscn").override_resolution(1920, 1080); flat = scene.flat()
img = rb.HostImage(1080, 1920)
m = rb.Renderer(rb.RendererConfig(1024, 12), devices=list(range(8)))
for _ in range(2): m.render_frame(flat, out=img.array)
m.close()
PY
tail -1 $OUT/ncu_multi8_$TAG.log | cut -c1-200

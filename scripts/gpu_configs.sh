#!/bin/bash
# BASELINE.json configs 3 (4K, one GPU's share), 4 (100k objects, BVH) and 5 (glass/metal, 32 bounces) on one GPU.
# Usage: bash scripts/gpu_configs.sh <tag>
TAG=${1:-c}; OUT=gpurun_out; mkdir -p $OUT
python tests/tools/make_synth_scenes.py /tmp/synth 2>&1 | tail -1
echo "== config3 (4K, 512 spp on one GPU = the per-GPU share of 4096 spp on 8)"
timeout 900 python bench.py --width 3840 --height 2160 --spp 512 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_config3_$TAG.json
echo "== config4 (100k objects, 1080p 256 spp)"
timeout 900 python bench.py --scene /tmp/synth/config4.rscn --spp 256 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_config4_$TAG.json
echo "== config5 (glass/metal lattice, 1080p 1024 spp, 32 bounces)"
timeout 900 python bench.py --scene /tmp/synth/config5.rscn --spp 1024 --bounces 32 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_config5_$TAG.json

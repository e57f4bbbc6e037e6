#!/bin/bash
# driver-style final check on one GPU: reference arm, then this repo's arm (default flags), smoke
TAG=r03w; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1; nproc >> $OUT/gpu_$TAG.txt
echo "== reference arm"; (time timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3) 2>&1 | tail -5 | tee $OUT/bench_ref_$TAG.json | cut -c1-500
echo "== b200 arm"; (time timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3) 2>&1 | tail -5 | tee $OUT/bench_$TAG.json | cut -c1-700
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke_$TAG.log

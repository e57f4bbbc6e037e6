#!/bin/bash
# 1-GPU validation of the round-3 host changes: full GPU suite, smoke, the default bench line (with other_configs), reference arm
TAG=r03b; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1; nproc >> $OUT/gpu_$TAG.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
echo "== bench"; (time timeout 600 python bench.py) 2>&1 | tail -5 | tee $OUT/bench_$TAG.json | cut -c1-1500
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref_$TAG.json | cut -c1-400

#!/bin/bash
# ncu only: launch list + one full capture of the render kernel (64 spp).  Usage: bash scripts/gpu_prof.sh <tag> [accel]
TAG=${1:-p}; ACC=${2:-auto}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --accel $ACC --steps 2 --warmup 1 --spp 64 --no-cpu-baseline > $OUT/ncu_launches_$TAG.log 2>&1
tail -1 $OUT/ncu_launches_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -c 1 -o $OUT/prof_$TAG -f \
    python bench.py --accel $ACC --steps 1 --warmup 0 --spp 64 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log

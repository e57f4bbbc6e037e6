#!/bin/bash
# ncu --set full captures of the render kernel on configs 4 (cooperative hierarchy) and 5 (fused scan, 24-member clusters)
TAG=${1:-r03q}; OUT=gpurun_out; mkdir -p $OUT
for c in 4 5; do
timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_kernel -c 1 -o $OUT/prof_config${c}_$TAG -f \
    python bench.py --config $c --steps 1 --warmup 0 --spp 16 --no-cpu-baseline > $OUT/ncu_config${c}_$TAG.log 2>&1
tail -1 $OUT/ncu_config${c}_$TAG.log | cut -c1-200
done

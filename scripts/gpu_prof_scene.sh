#!/bin/bash
# one full ncu capture of the render kernel on a synthetic scene.  Usage: bash scripts/gpu_prof_scene.sh <tag> <config4|config5> <accel> <spp> [bounces]
TAG=$1; CFG=$2; ACC=${3:-auto}; SPP=${4:-8}; BNC=${5:-12}
OUT=gpurun_out; mkdir -p $OUT
python tests/tools/make_synth_scenes.py /tmp/synth $CFG 2>&1 | tail -1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:render_kernel -c 1 -o $OUT/prof_$TAG -f \
    python bench.py --scene /tmp/synth/$CFG.rscn --accel $ACC --steps 1 --warmup 0 --spp $SPP --bounces $BNC --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log

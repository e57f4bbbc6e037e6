#!/bin/bash
# Sweep an environment variable on a synthetic scene.  Usage: bash scripts/gpu_sweep_scene.sh <tag> <config4|config5> <VAR> "<values>" <accel> <spp> [bounces]
TAG=$1; CFG=$2; VAR=$3; VALS=$4; ACC=${5:-auto}; SPP=${6:-64}; BNC=${7:-12}
OUT=gpurun_out; mkdir -p $OUT
python tests/tools/make_synth_scenes.py /tmp/synth $CFG 2>&1 | tail -1
for v in $VALS; do
  echo "== $VAR=$v"
  env $VAR=$v timeout 600 python bench.py --scene /tmp/synth/$CFG.rscn --accel $ACC --spp $SPP --bounces $BNC --no-cpu-baseline --steps 2 --warmup 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])" | tee -a $OUT/sweep_$TAG.txt
done

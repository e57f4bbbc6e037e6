#!/bin/bash
# Sweep an environment variable over values with short benches.  Usage: bash scripts/gpu_sweep.sh <tag> <VAR> "<values>" [accel]
TAG=$1; VAR=$2; VALS=$3; ACC=${4:-auto}
OUT=gpurun_out; mkdir -p $OUT
for v in $VALS; do
  echo "== $VAR=$v"
  env $VAR=$v timeout 600 python bench.py --accel $ACC --no-cpu-baseline --steps 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])" | tee -a $OUT/sweep_$TAG.txt
done

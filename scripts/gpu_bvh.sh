#!/bin/bash
# Hierarchy visit: bvh2 parity tests, config 4 bench with the cooperative (auto) and the per-lane (bvh) traversal.
TAG=${1:-b}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q -k "bvh" 2>&1 | tail -8 | tee $OUT/pytest_bvh_$TAG.log
python tests/tools/make_synth_scenes.py /tmp/synth config4 2>&1 | tail -1
for A in auto bvh; do
  echo "== config4 $A"
  timeout 900 python bench.py --scene /tmp/synth/config4.rscn --accel $A --spp 64 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_config4_${A}_$TAG.json | cut -c1-330
done

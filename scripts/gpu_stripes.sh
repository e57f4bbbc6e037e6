#!/bin/bash
# GPU visit for the row-stripe sharding: the full parity suite (stripe tests included) and the headline bench.
# Usage: bash scripts/gpu_stripes.sh <tag>
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu_$TAG.log
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -1 | tee $OUT/bench_$TAG.json

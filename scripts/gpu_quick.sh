#!/bin/bash
# Quick GPU visit while iterating on a kernel: parity subset for one accel + short benches.
# Usage: bash scripts/gpu_quick.sh <tag> [accel] [pytest -k expr]
TAG=${1:-q}; ACC=${2:-auto}; KEXPR=${3:-"fused or edge or progressive"}
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | tail -8 | tee $OUT/pytest_quick_$TAG.log
echo "== bench $ACC"; timeout 600 python bench.py --accel $ACC --no-cpu-baseline --steps 3 2>&1 | tail -1 | tee $OUT/bench_quick_$TAG.json
echo "== bench coop"; timeout 600 python bench.py --accel coop --no-cpu-baseline --steps 3 2>&1 | tail -1 | tee $OUT/bench_quick_coop_$TAG.json

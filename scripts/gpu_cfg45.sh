#!/bin/bash
TAG=${1:-cfg}; OUT=gpurun_out; mkdir -p $OUT
echo "== parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config4 or config5 or bvh2 or cluster_sizes" 2>&1 | tail -2
for c in 4 5; do
timeout 600 python bench.py --config $c --no-cpu-baseline --steps 2 --warmup 1 --spp 256 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('config$c', d['value'], d['roofline']['kernel_ms'])" | tee -a $OUT/cfg45_$TAG.txt
done

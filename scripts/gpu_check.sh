#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, bench, ncu launch list and one full capture of the render kernel.
# Usage (from the repo root, under gpurun): bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
nproc >> $OUT/gpu_$TAG.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> $OUT/gpu_$TAG.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu_$TAG.log
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -3 | tee $OUT/bench_$TAG.json
for A in brute bvh cluster; do echo "== bench $A"; timeout 600 python bench.py --accel $A --no-cpu-baseline --steps 3 2>&1 | tail -1 | tee $OUT/bench_${A}_$TAG.json; done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --spp 64 --no-cpu-baseline > $OUT/ncu_launches_$TAG.log 2>&1
tail -2 $OUT/ncu_launches_$TAG.log
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -c 1 -o $OUT/prof_$TAG -f \
    python bench.py --steps 1 --warmup 0 --spp 64 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log
ls -la $OUT

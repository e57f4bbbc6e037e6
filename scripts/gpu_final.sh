#!/bin/bash
# Final GPU visit of a round within a small budget: parity suite, headline bench, smoke, ncu launch list and one full capture.
TAG=${1:-final}; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1; nproc >> $OUT/gpu_$TAG.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/pytest_gpu_$TAG.log
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/bench_$TAG.json
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke_$TAG.log
echo "== ncu full"
timeout 300 ncu --set full --metrics smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum --clock-control none --import-source on -k regex:render_kernel -c 1 -o $OUT/prof_$TAG -f \
    python bench.py --steps 1 --warmup 0 --spp 64 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
tail -1 $OUT/ncu_full_$TAG.log
echo "== ncu launch list"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --spp 64 --no-cpu-baseline > $OUT/ncu_launches_$TAG.log 2>&1
tail -1 $OUT/ncu_launches_$TAG.log | cut -c1-200

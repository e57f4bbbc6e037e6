#!/bin/bash
# Round-3 first GPU visit: refined clustering A/B, the four prepared variants, one full ncu capture with the FP32 op counters.
TAG=r03a; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1; nproc >> $OUT/gpu_$TAG.txt
B="python bench.py --no-cpu-baseline --steps 3"
one() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['roofline']['kernel_ms'])" | tee -a $OUT/variants_$TAG.txt; }
for rep in 1 2; do
  timeout 300 $B 2>&1 | tail -1 | one refine_on
  RDR_CLUSTER_REFINE=0 timeout 300 $B 2>&1 | tail -1 | one refine_off
done
bash scripts/gpu_variants.sh $TAG "ballot rho scanp all3 chunk"
echo "== config5 / config4 with refine"
python tests/tools/make_synth_scenes.py /tmp/synth 2>&1 | tail -1
timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 1 --scene /tmp/synth/config5.rscn --bounces 32 --spp 256 2>&1 | tail -1 | tee $OUT/bench_config5_$TAG.json | cut -c1-200
echo "== ncu full (shipped config) with fp32 op counters"
timeout 300 ncu --set full --metrics smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum,smsp__sass_thread_inst_executed_ops_fadd_fmul_ffma_pred_on.sum \
    --clock-control none --import-source on -k regex:render_kernel -c 1 -o $OUT/prof_$TAG -f \
    python bench.py --steps 1 --warmup 0 --spp 64 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log | cut -c1-300

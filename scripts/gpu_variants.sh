#!/bin/bash
# Kernel experiments: short benches of several builds of the library (raydar_b200/libraydar_cuda_<name>.so, made
# with raydar_b200.build.build_variant) in one GPU visit, each after a parity subset.
# A name may carry a CTA shape of the fused kernel, name@N (RDR_FUSED_CTA: 0 = 3 x 256, 1 = 896, 2 = 1024, 3 = 768 threads).
# Usage: bash scripts/gpu_variants.sh <tag> "<names>" [bench args]
TAG=$1; NAMES=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
for vv in $NAMES; do
  v=${vv%@*}; CTA=3; [ "$vv" != "$v" ] && CTA=${vv#*@}
  export RDR_FUSED_CTA=$CTA
  LIBV=$PWD/raydar_b200/libraydar_cuda_$v.so; [ "$v" = base ] && LIBV=$PWD/raydar_b200/libraydar_cuda.so
  echo "== $vv"
  RAYDAR_CUDA_LIB=$LIBV timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "accumulator_bit_exact and fused or first_hit and fused or edge" 2>&1 | tail -1
  for rep in 1 2; do
  RAYDAR_CUDA_LIB=$LIBV timeout 600 python bench.py --no-cpu-baseline --no-other-configs --steps 3 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$vv', d['value'], d['roofline']['kernel_ms'])" | tee -a $OUT/variants_$TAG.txt
  done
done

#!/bin/bash
# r02f: cold/parked lane state + prefix staging as the default: full parity suite, headline bench, configs 4 and 5 with CTA shapes.
TAG=r02f; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu_$TAG.log
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_$TAG.json
python tests/tools/make_synth_scenes.py /tmp/synth 2>&1 | tail -1
for c in -1 3 4; do
echo "== config4 RDR_BVH2_CTA=$c"
RDR_BVH2_CTA=$c timeout 600 python bench.py --scene /tmp/synth/config4.rscn --spp 128 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('config4 cta=$c', d['value'], d['roofline']['kernel_ms'])" | tee -a $OUT/variants_$TAG.txt
done
for c in -1 4; do
echo "== config5 RDR_FUSED_CTA=$c"
RDR_FUSED_CTA=$c timeout 600 python bench.py --scene /tmp/synth/config5.rscn --spp 256 --bounces 32 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('config5 cta=$c', d['value'], d['roofline']['kernel_ms'])" | tee -a $OUT/variants_$TAG.txt
done

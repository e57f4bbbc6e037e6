#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): multi-handle parity test + bench under torchrun at N ranks.
N=${1:-2}
TAG=${2:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus_${TAG}_n$N.txt 2>&1
echo "== pytest multi"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_multi_${TAG}_n$N.log
for n in 1 $N; do
  echo "== bench n=$n"
  if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_${TAG}_n1of$N.json
  else NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 2>&1 | tail -2 | tee $OUT/bench_${TAG}_n$n.json; fi
done
echo "== CLI multi"; timeout 300 ./raydar_b200/host/raydar-cuda --gpus $N --max-sample-count 64 --resolution 1920x1080 -o $OUT/cli_${TAG}_n$N.png scenes/benchmark.rscn 2>&1 | tail -12 | tee $OUT/cli_${TAG}_n$N.log

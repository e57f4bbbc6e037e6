#!/bin/bash
# Multi-GPU visit for the row-stripe partition (gpurun --gpus N): multi-handle parity tests (sample ranges and stripes),
# bench under torchrun with both partitions, CLI with --partition stripes compared with the 1-GPU image.
N=${1:-2}; TAG=${2:-r02}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest multi"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_multi_${TAG}_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== bench n=$N sample ranges"
NCCL_DEBUG=WARN timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_${TAG}_n$N.json
echo "== bench n=$N row stripes"
NCCL_DEBUG=WARN timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --partition stripes 2>&1 | tail -1 | tee $OUT/bench_stripes_${TAG}_n$N.json
echo "== CLI: 1 GPU vs $N GPUs with stripes (bit-identical PNG expected)"
timeout 300 ./raydar_b200/host/raydar-cuda --max-sample-count 64 --resolution 1920x1080 -o $OUT/cli_${TAG}_n1.png scenes/benchmark.rscn 2>&1 | tail -3
timeout 300 ./raydar_b200/host/raydar-cuda --gpus $N --partition stripes --max-sample-count 64 --resolution 1920x1080 -o $OUT/cli_${TAG}_stripes_n$N.png scenes/benchmark.rscn 2>&1 | tail -3
cmp $OUT/cli_${TAG}_n1.png $OUT/cli_${TAG}_stripes_n$N.png && echo "CLI images identical" | tee $OUT/cli_cmp_${TAG}_n$N.log
rm -f $OUT/cli_${TAG}_n1.png $OUT/cli_${TAG}_stripes_n$N.png

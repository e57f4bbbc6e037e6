#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): multi-handle parity test, bench under torchrun at N ranks (headline config and
# BASELINE config 3: 4K, 4096 spp split over the ranks), C++ CLI on N devices.
N=${1:-2}
TAG=${2:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus_${TAG}_n$N.txt 2>&1
echo "== pytest multi"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_multi_${TAG}_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== bench n=$N (1080p, 1024 spp per rank)"
NCCL_DEBUG=WARN timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_${TAG}_n$N.json
echo "== bench n=$N config 3 (4K, 4096 spp in total)"
NCCL_DEBUG=WARN timeout 900 $TR bench.py --gpus $N --width 3840 --height 2160 --spp $((4096 / N)) --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_config3_${TAG}_n$N.json
echo "== reference arm under torchrun (rank 0 only)"
timeout 900 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref_${TAG}_n$N.json
echo "== CLI multi"; timeout 300 ./raydar_b200/host/raydar-cuda --gpus $N --max-sample-count 64 --resolution 1920x1080 -o $OUT/cli_${TAG}_n$N.png scenes/benchmark.rscn 2>&1 | tail -12 | tee $OUT/cli_${TAG}_n$N.log

#!/bin/bash
# 2-GPU validation: full GPU suite (incl. multi-GPU + CLI 2-GPU tests), torchrun bench at N = 2 (strong scaling, IPC combine, multi-handle e2e)
TAG=r03c; OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu_$TAG.txt 2>&1; nproc >> $OUT/gpu_$TAG.txt
nvidia-smi topo -m >> $OUT/gpu_$TAG.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -40 | tee $OUT/pytest_gpu_$TAG.log
echo "== bench N=2"; (time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3) 2>&1 | tail -8 | tee $OUT/bench_n2_$TAG.json | cut -c1-3000
echo "== bench N=2 stripes"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 --partition stripes 2>&1 | tail -3 | tee $OUT/bench_n2_stripes_$TAG.json | cut -c1-1500

#!/bin/bash
# Short multi-GPU visit (gpurun --gpus N): multi-handle parity tests and the bench with both partitions.
N=${1:-8}; TAG=${2:-r02}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest multi"; timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_multi_${TAG}_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== bench n=$N sample ranges"
NCCL_DEBUG=WARN timeout 300 $TR bench.py --gpus $N --steps 3 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_${TAG}_n$N.json
echo "== bench n=$N row stripes"
NCCL_DEBUG=WARN timeout 300 $TR bench.py --gpus $N --steps 3 --warmup 3 --partition stripes 2>&1 | tail -1 | tee $OUT/bench_stripes_${TAG}_n$N.json

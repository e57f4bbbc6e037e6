#!/bin/bash
TAG=${1:-r03o}; OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_gpu_$TAG.log
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_n2_$TAG.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2', d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['parity_check'])"
echo "== bench N=1"; timeout 600 python bench.py --no-cpu-baseline --no-other-configs 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1', d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['render_sample_ms'])"

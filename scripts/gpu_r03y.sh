#!/bin/bash
OUT=gpurun_out
timeout 300 python bench.py --no-cpu-baseline --no-other-configs --steps 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1', d['value'], d['roofline']['frac'], d['roofline'].get('frac_at_clock_under_load'), d['roofline'].get('executed'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2', d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('frac_at_clock_under_load'), d['parity_check'])"

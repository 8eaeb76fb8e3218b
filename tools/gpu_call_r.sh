#!/bin/bash
# final 2-GPU sanity on the last commit
timeout 300 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=2', round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1))"

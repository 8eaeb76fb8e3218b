#!/bin/bash
# round 2: disable_pairwise_lang_attn (segment-structured mask) in K2/K3/K4/export -- parity, then the whole suite and the headline bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention" 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_model.py -x -q -k "pairwise or pretrain_step" 2>&1 | tail -4
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['ms_per_step'],3), round(d['value'],1))"

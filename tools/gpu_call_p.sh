#!/bin/bash
# round 2: A/B of the pair-kernel heuristics after the uniform-issue fix (env knobs only, no code change)
mkdir -p gpurun_out
run() { env $2 timeout 300 python bench.py --steps 15 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],3), round(d['value'],1), round(d['roofline']['gemm_ms_per_step'],3))"; }
run default_a X=1
run pair192_k0 MERLOT_PAIR192_MINK=0
run pair_everywhere MERLOT_GEMM_PAIR=2
run both "MERLOT_PAIR192_MINK=0 MERLOT_GEMM_PAIR=2"
run no_pair MERLOT_GEMM_PAIR=0
run default_b X=1

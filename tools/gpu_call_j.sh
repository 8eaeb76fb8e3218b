#!/bin/bash
# round 2: fused LayerNorm backward through bulk-copy rings -- parity, timing, then the whole GPU suite and the N=1 bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -k "layernorm" -x -q 2>&1 | tail -5
timeout 200 python tools/perf_rowwise.py 2>&1 | grep -i "ln_" | cut -c1-150 | tee gpurun_out/r02_perf_rowwise_ring.txt
timeout 300 python -m pytest tests/test_gpu_stem.py -k training_step -s -q 2>&1 | grep -a "stem boundary\|passed\|failed" | cut -c1-700
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/r02_bench_ring.json 2> gpurun_out/r02_bench_ring.err; tail -c 1800 gpurun_out/r02_bench_ring.json

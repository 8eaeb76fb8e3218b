#!/bin/bash
# round 2: stem kernels (weight standardisation fwd/bwd one warp per channel, GroupNorm backward block-level dgamma/dbeta) -- parity, then the merlot.yaml-as-shipped step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stem.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --stem hybrid --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_hybrid_stem.json 2> gpurun_out/r02_bench_hybrid_stem.err; tail -1 gpurun_out/r02_bench_hybrid_stem.err | cut -c1-300; head -c 400 gpurun_out/r02_bench_hybrid_stem.json; echo
MERLOT_NO_PDL=1 MERLOT_NO_SIDE_STREAM=1 timeout 600 python tools/timeline_step.py --hybrid-stem 2>&1 | grep -v Warn | cut -c1-150 | tee gpurun_out/r02_timeline_hybrid_stem_serial.txt | head -24

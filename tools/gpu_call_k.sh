#!/bin/bash
# round 2: K1 in-kernel role counters in the final state (cycles the MMA warp spends per MMA in situ, by tile width), and the
# merlot.yaml-as-shipped (hybrid stem) step: bench line + per-kernel timeline
mkdir -p gpurun_out
timeout 300 python tools/gemm_stalls.py 2>&1 | cut -c1-330 | tee gpurun_out/r02_k1_stall_counters.txt
timeout 600 python bench.py --stem hybrid --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_hybrid_stem.json 2> gpurun_out/r02_bench_hybrid_stem.err; tail -2 gpurun_out/r02_bench_hybrid_stem.err | cut -c1-300; head -c 700 gpurun_out/r02_bench_hybrid_stem.json; echo
MERLOT_NO_PDL=1 MERLOT_NO_SIDE_STREAM=1 timeout 600 python tools/timeline_step.py --hybrid-stem 2>&1 | grep -v Warn | cut -c1-150 | tee gpurun_out/r02_timeline_hybrid_stem_serial.txt | head -45

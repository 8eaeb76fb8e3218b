#!/bin/bash
# round 2, call A: full GPU test suite, baseline bench, isolated timings + ncu --set full of every hot kernel
mkdir -p gpurun_out
rm -f gpurun_out/r02_fullsize_parity.json
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_tests_a.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests_a.txt
tail -30 gpurun_out/r02_tests_a.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
tail -3 gpurun_out/r02_bench_a.err; cat gpurun_out/r02_bench_a.json | head -c 1500
timeout 300 python tools/ncu_targets.py --reps 20 > gpurun_out/r02_targets_timing_a.txt 2>&1
cat gpurun_out/r02_targets_timing_a.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm|attn|ln_|adamw" -f -o gpurun_out/r02_targets_a python tools/ncu_targets.py --reps 1 > gpurun_out/r02_ncu_a.log 2>&1
tail -3 gpurun_out/r02_ncu_a.log
ls -la gpurun_out/*.ncu-rep

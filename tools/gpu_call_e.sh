#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -p no:cacheprovider -k "attention" > gpurun_out/r02_tests_attn_e.txt 2>&1
echo "pytest attention rc=$?" >> gpurun_out/r02_tests_attn_e.txt
tail -15 gpurun_out/r02_tests_attn_e.txt
timeout 300 python tools/ncu_targets.py --reps 20 --only attention > gpurun_out/r02_targets_timing_e.txt 2>&1
cat gpurun_out/r02_targets_timing_e.txt
timeout 300 python tools/attn_phases.py > gpurun_out/r02_attn_phases_e.txt 2>&1
cat gpurun_out/r02_attn_phases_e.txt

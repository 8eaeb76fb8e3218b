#!/bin/bash
# round 2, call B: attention v2 kernels -- parity first, then timings, bench, ncu summary (report summarised ON the box)
mkdir -p gpurun_out
rm -f gpurun_out/r02_fullsize_parity.json
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 600 -p no:cacheprovider -k attention > gpurun_out/r02_tests_attn_b.txt 2>&1
echo "pytest attention rc=$?" >> gpurun_out/r02_tests_attn_b.txt
tail -15 gpurun_out/r02_tests_attn_b.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/r02_tests_b.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests_b.txt
tail -25 gpurun_out/r02_tests_b.txt
timeout 300 python tools/ncu_targets.py --reps 20 > gpurun_out/r02_targets_timing_b.txt 2>&1
cat gpurun_out/r02_targets_timing_b.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
tail -3 gpurun_out/r02_bench_b.err; head -c 1200 gpurun_out/r02_bench_b.json
timeout 900 ncu --set full --clock-control none -k regex:"gemm|attn|ln_|adamw" -f -o /tmp/r02_targets_b python tools/ncu_targets.py --reps 1 > gpurun_out/r02_ncu_b.log 2>&1
tail -2 gpurun_out/r02_ncu_b.log
python tools/ncu_summary.py /tmp/r02_targets_b.ncu-rep gpurun_out/r02_ncu_kernels_b.json > gpurun_out/r02_ncu_kernels_b.txt 2>&1
cat gpurun_out/r02_ncu_kernels_b.txt | cut -c1-200
du -sh gpurun_out

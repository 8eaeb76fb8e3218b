#!/bin/bash
# round 2 final: smoke, whole GPU suite, N=1 bench line, refreshed ncu --set full summary of every hot kernel and the launch list
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; head -c 500 gpurun_out/r02_bench_n1_final.json; echo
timeout 600 python tools/ncu_targets.py --reps 20 2>&1 | grep TARGET | tee gpurun_out/r02_targets_timing_final.txt
timeout 900 ncu --set full --clock-control none -k regex:"gemm|attn|ln_|adamw" -f -o /tmp/r02_targets_m python tools/ncu_targets.py --reps 1 > /tmp/ncu_m.log 2>&1
python tools/ncu_summary.py /tmp/r02_targets_m.ncu-rep gpurun_out/r02_ncu_kernels_final.json > gpurun_out/r02_ncu_kernels_final.txt 2>&1
grep -c . gpurun_out/r02_ncu_kernels_final.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1100 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /tmp/launches_bench.log 2>&1
wc -l gpurun_out/r02_launches_final.csv
du -sh gpurun_out

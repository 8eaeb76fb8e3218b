#!/bin/bash
# round 2: evidence run -- ncu --set full summary of every hot kernel (final state), launch list of one step, other BASELINE configs
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:"gemm|attn|ln_|adamw" -f -o /tmp/r02_targets_h python tools/ncu_targets.py --reps 1 > gpurun_out/r02_ncu_h.log 2>&1
tail -2 gpurun_out/r02_ncu_h.log
python tools/ncu_summary.py /tmp/r02_targets_h.ncu-rep gpurun_out/r02_ncu_kernels_final.json > gpurun_out/r02_ncu_kernels_final.txt 2>&1
cut -c1-170 gpurun_out/r02_ncu_kernels_final.txt
for C in 1 4 5; do
timeout 600 python bench.py --config $C --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg$C.json 2> gpurun_out/r02_bench_cfg$C.err
tail -1 gpurun_out/r02_bench_cfg$C.err | cut -c1-300; head -c 900 gpurun_out/r02_bench_cfg$C.json; echo
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1100 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
wc -l gpurun_out/r02_launches_final.csv
du -sh gpurun_out

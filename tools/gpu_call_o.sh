#!/bin/bash
# round 2 final: the non-headline bench lines with the final kernels (merlot.yaml as shipped; BASELINE configs 1, 4, 5)
mkdir -p gpurun_out
timeout 600 python bench.py --stem hybrid --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_hybrid_stem.json 2> gpurun_out/r02_bench_hybrid_stem.err; head -c 330 gpurun_out/r02_bench_hybrid_stem.json; echo
for C in 1 4 5; do
timeout 600 python bench.py --config $C --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg$C.json 2> gpurun_out/r02_bench_cfg$C.err
head -c 330 gpurun_out/r02_bench_cfg$C.json; echo
done
MERLOT_NO_PDL=1 MERLOT_NO_SIDE_STREAM=1 timeout 600 python tools/timeline_step.py --hybrid-stem 2>&1 | grep -v Warn | cut -c1-150 > gpurun_out/r02_timeline_hybrid_stem_serial.txt; grep "ws_\|kernels=" gpurun_out/r02_timeline_hybrid_stem_serial.txt

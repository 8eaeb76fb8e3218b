#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_phases.py > gpurun_out/r02_attn_phases_c.txt 2>&1
cat gpurun_out/r02_attn_phases_c.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_kernel|attn_fwd_kernel" -c 2 -f -o gpurun_out/r02_attn_src_c python tools/ncu_targets.py --reps 1 --only attention > gpurun_out/r02_ncu_c.log 2>&1
tail -2 gpurun_out/r02_ncu_c.log
du -sh gpurun_out

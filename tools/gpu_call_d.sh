#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x > gpurun_out/r02_tests_d.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests_d.txt
tail -6 gpurun_out/r02_tests_d.txt
timeout 300 python tools/ncu_targets.py --reps 20 > gpurun_out/r02_targets_timing_d.txt 2>&1
cat gpurun_out/r02_targets_timing_d.txt
timeout 300 python tools/attn_phases.py > gpurun_out/r02_attn_phases_d.txt 2>&1
cat gpurun_out/r02_attn_phases_d.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err
tail -2 gpurun_out/r02_bench_d.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_d.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['gemm_ms_per_step'], d['attention']['shapes'])
PY

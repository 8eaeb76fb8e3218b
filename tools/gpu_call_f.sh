#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_model.py -m gpu -q --timeout 800 -p no:cacheprovider -k "two_rank or partial_stack" > gpurun_out/r02_tests_dist_f.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_tests_dist_f.txt
tail -25 gpurun_out/r02_tests_dist_f.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2_f.json 2> gpurun_out/r02_bench_n2_f.err
tail -3 gpurun_out/r02_bench_n2_f.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n2_f.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'])
except Exception as e: print('bench parse failed', e)
PY

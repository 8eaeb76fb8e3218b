#!/bin/bash
mkdir -p gpurun_out
run() { N=$1; C=$2; TAG=$3
MERLOT_DP_COMM_CTAS=$C timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n${N}_${TAG}.json 2> gpurun_out/r02_bench_n${N}_${TAG}.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n${N}_${TAG}.json').read().strip().splitlines()[-1])
    dp=d['data_parallel']
    print('N=$N comm_ctas=$C', {k:round(d[k],2) for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), 'no-allreduce ms', round(dp['ms_per_step_without_grad_allreduce'],2), 'exposed', round(dp['grad_allreduce_exposed_ms'],2))
except Exception as e: print('bench parse failed', e); print(open('gpurun_out/r02_bench_n${N}_${TAG}.err').read()[-1500:])
PY
}
run 8 24 c24
run 8 32 c32
run 8 0 c0
run 4 24 c24

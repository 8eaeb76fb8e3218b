#!/bin/bash
# round 2: data-parallel A/B at N GPUs -- K1 ticketed tile schedule under the all-reduce window vs static grids on the reduced SM count
N=${1:-8}
mkdir -p gpurun_out
run() { # tag, env
  env $2 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_dp_n${N}_$1.json 2> gpurun_out/r02_dp_n${N}_$1.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_dp_n${N}_$1.json").read().strip().splitlines()[-1])
    print("$1", d["ms_per_step"], d["value"], d.get("data_parallel"))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/r02_dp_n${N}_$1.err").read()[-1500:])
PY
}
run dyn_a MERLOT_DP_GEMM_DYNAMIC=1
run static_a MERLOT_DP_GEMM_DYNAMIC=0
run dyn_b MERLOT_DP_GEMM_DYNAMIC=1
run static_b MERLOT_DP_GEMM_DYNAMIC=0


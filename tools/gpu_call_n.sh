#!/bin/bash
# round 2 final: 2-GPU sanity -- NCCL parity test of the data-parallel step, then the bench exactly as the driver launches it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err
tail -2 gpurun_out/r02_bench_n2_final.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n2_final.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "e2e", "clocks")}); print(d["data_parallel"])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-300

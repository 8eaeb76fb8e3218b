"""Run warm-up steps, then ONE pretraining step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from merlot_b200.train import model_fn_builder, synthetic_batch  # noqa: E402

cfg = bench.load_config()
if len(sys.argv) > 1 and sys.argv[1] == "nodrop":
    cfg.model["hidden_dropout_prob"] = 0.0
fn = model_fn_builder(cfg)
feats = synthetic_batch(cfg, bench.PER_GPU_BATCH, seed=0)
for _ in range(2):
    fn(feats).train_op()
torch.cuda.synchronize()
torch.cuda.profiler.start()
fn(feats).train_op()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")

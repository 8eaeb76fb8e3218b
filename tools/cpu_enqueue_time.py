"""Host time to ENQUEUE one step (no synchronisation inside the loop) vs the device time of the step: the launch path must stay
well below the device time or the GPU starves."""
import sys, time
import torch
sys.path.insert(0, ".")
import bench
from merlot_b200.train import model_fn_builder, synthetic_batch
cfg = bench.load_config()
fn = model_fn_builder(cfg)
feats = synthetic_batch(cfg, bench.PER_GPU_BATCH, seed=0)
for _ in range(3):
    fn(feats).train_op()
torch.cuda.synchronize()
n = 10
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    fn(feats).train_op()
t1 = time.perf_counter()
e1.record()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / n:.2f} ms/step; device {e0.elapsed_time(e1) / n:.2f} ms/step; wall incl. drain {1e3 * (t2 - t0) / n:.2f} ms/step")
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    fn(feats).train_op()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

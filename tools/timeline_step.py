"""Kernel timeline of one step via torch.profiler (CUPTI): busy time, idle gaps, per-kernel totals (non-ncu, warm caches)."""
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
import bench  # noqa: E402
from merlot_b200.train import model_fn_builder, synthetic_batch  # noqa: E402

if "--hybrid-stem" in sys.argv:  # merlot.yaml as shipped
    bench.STEM = "hybrid"
cfg = bench.load_config()
fn = model_fn_builder(cfg)
feats = synthetic_batch(cfg, bench.PER_GPU_BATCH, seed=0)
for _ in range(3):
    fn(feats).train_op()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        fn(feats).train_op()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
span = evs[-1].time_range.end - evs[0].time_range.start
busy = sum(e.time_range.end - e.time_range.start for e in evs)
# idle = span minus the UNION of the kernel intervals (PDL / side-stream kernels overlap, so a plain sum double-counts)
pos, cur_end = [], evs[0].time_range.end
for e in evs[1:]:
    if e.time_range.start > cur_end:
        pos.append(e.time_range.start - cur_end)
    cur_end = max(cur_end, e.time_range.end)
print(f"kernels={len(evs)} span={span / 1e3:.3f} ms busy={busy / 1e3:.3f} ms idle={sum(pos) / 1e3:.3f} ms "
      f"(mean gap {sum(pos) / max(1, len(pos)):.2f} us, gaps>5us: {sum(1 for g in pos if g > 5)})  [2 steps]")
agg = defaultdict(lambda: [0, 0.0])
for e in evs:
    n = e.name
    if ">(" in n:
        n = n[: n.rindex(">(") + 1]
    else:
        n = n.split("(")[0]
    n = n.replace("void ", "").replace("(int)", "").replace("(bool)", "").replace("mb::", "")
    agg[n][0] += 1
    agg[n][1] += e.time_range.end - e.time_range.start
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{us / 2e3:8.3f} ms/step  n={n // 2:4d}  avg={us / n:7.1f} us  {k[:90]}")

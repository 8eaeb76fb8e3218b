"""How repeatable is the backward pass?  rel-Frobenius distance between the gradient arenas of repeated backward passes of
the SAME forward (only fp32 atomics ordering may differ), with and without the side stream."""
import os
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from merlot_b200.train import model_fn_builder, synthetic_batch  # noqa: E402

cfg = bench.load_config()
cfg.model["hidden_dropout_prob"] = 0.0
fn = model_fn_builder(cfg)
store = fn.store
feats = synthetic_batch(cfg, 2, seed=0)
for mode in ("0", "1", "0", "1"):
    os.environ["MERLOT_NO_SIDE_STREAM"] = mode
    spec = fn(feats)
    gs = []
    for _ in range(4):
        store.g.zero_()
        spec.model.backward()
        torch.cuda.synchronize()
        gs.append(store.g.clone())
    rels = [((g - gs[0]).norm() / gs[0].norm()).item() for g in gs[1:]]
    # per-region: which parameters differ most
    worst = []
    for name, e in store.entries.items():
        a, b = gs[0][e.offset:e.offset + e.numel], gs[1][e.offset:e.offset + e.numel]
        n = a.norm().item()
        if n > 0:
            worst.append((((a - b).norm() / n).item(), name))
    worst.sort(reverse=True)
    print(f"no_side_stream={mode}: rel diffs vs first backward: {['%.2e' % r for r in rels]}; worst params: "
          f"{[(round(w, 5), n[-60:]) for w, n in worst[:4]]}", flush=True)

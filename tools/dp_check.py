"""N-rank NCCL check (torchrun): after data-parallel steps on different per-rank batches every rank must hold bit-identical
parameters, and the 2-rank gradient must equal the mean of the two single-rank gradients (CrossShardOptimizer MEAN)."""
import os
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from merlot_b200.train import DataParallel, model_fn_builder, synthetic_batch  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dp = DataParallel("nccl")
cfg = bench.load_config()
cfg.model.update(num_hidden_layers=2, num_vision_transformer_hidden_layers=2, num_lang_transformer_hidden_layers=2,
                 hidden_dropout_prob=0.0, image_size=[64, 96])
fn = model_fn_builder(cfg, dist=dp, device=torch.device("cuda", local))
store = fn.store
feats = synthetic_batch(cfg, 2, seed=100 + dp.rank, device=store.device)
# gradient of this rank alone
spec = fn(feats)
store.g.zero_()
spec.model.backward()
g_local = store.g.clone()
gl = [torch.empty_like(g_local) for _ in range(dp.world)]
dp.dist.all_gather(gl, g_local)
g_mean = sum(gl) / dp.world
store.g.zero_()
# the real data-parallel path (bucketed async all-reduce inside backward), without the optimizer
pending = []
spec = fn(feats)
spec.model.backward(on_non_vit_grads_ready=lambda: pending.extend(dp.all_reduce_ranges_async(store.g, store.rest_ranges)))
pending.extend(dp.all_reduce_ranges_async(store.g, store.vit_ranges))
dp.wait_all(pending)
torch.cuda.synchronize()
rel = ((store.g / dp.world - g_mean).norm() / g_mean.norm()).item()
store.g.zero_()
for _ in range(3):
    s = fn(feats)
    s.train_op()
torch.cuda.synchronize()
pl = [torch.empty_like(store.p) for _ in range(dp.world)]
dp.dist.all_gather(pl, store.p)
same = all(torch.equal(pl[0], x) for x in pl[1:])
if dp.rank == 0:
    print(f"dp_check world={dp.world}: grad rel err vs mean of per-rank grads = {rel:.2e}; params identical across ranks = {same}; "
          f"loss={s.loss:.4f}")
    assert rel < 2e-3 and same
dp.barrier()

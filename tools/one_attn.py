"""One ViT-shape attention forward + backward a few times (for ncu source-level captures)."""
import sys

import torch

sys.path.insert(0, ".")
from merlot_b200 import ops  # noqa: E402

g = torch.Generator().manual_seed(0)
B, S, heads, H = 32, 266, 12, 768
qkv = (torch.randn(B * S, 3 * H, generator=g) * 0.5).bfloat16().cuda()
dctx = (torch.randn(B * S, H, generator=g) * 0.5).bfloat16().cuda()
for _ in range(3):
    ctx, lse = ops.attention_fwd(qkv, B, S, heads)
    ops.attention_bwd(qkv, ctx, dctx, lse, B, S, heads)
torch.cuda.synchronize()

"""K3 timing experiments: which pipeline (arithmetic warps / tensor pipe) bounds the chunk loop (results are wrong in modes != 0)."""
import sys
import torch
sys.path.insert(0, ".")
from merlot_b200 import ops
from merlot_b200._lib import lib
L = lib()
dev = "cuda"
g = torch.Generator().manual_seed(0)
B, S, heads = 32, 266, 12
H = heads * 64
qkv = (torch.randn(B * S, 3 * H, generator=g) * 0.5).bfloat16().to(dev)
dctx = (torch.randn(B * S, H, generator=g) * 0.5).bfloat16().to(dev)
ctx, lse = ops.attention_fwd(qkv, B, S, heads)
ws = ops.attention_bwd_workspace(B, S, heads, dev)
dqkv = torch.empty(B * S, 3 * H, dtype=torch.bfloat16, device=dev)
dsum = torch.empty(B, heads, S, dtype=torch.float32, device=dev)
for mode, name in ((0, "normal"), (1, "no arithmetic"), (2, "no MMA2"), (4, "no MMA1"), (6, "no MMAs"), (7, "nothing but the skeleton")):
    L.merlot_attention_debug_mode(mode)
    for _ in range(3):
        ops.attention_bwd(qkv, ctx, dctx, lse, B, S, heads, dqkv=dqkv, dq_accum=ws, dsum=dsum)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.attention_bwd(qkv, ctx, dctx, lse, B, S, heads, dqkv=dqkv, dq_accum=ws, dsum=dsum)
    e1.record()
    torch.cuda.synchronize()
    print(f"mode {mode} ({name:26s}): {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us per attention_bwd (dsum + K3 + finish)")
L.merlot_attention_debug_mode(0)

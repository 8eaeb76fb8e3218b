"""Per-phase cycle breakdown of K2 / K3 (one softmax thread per CTA, merlot_attention_debug_counters) at the ViT shape and the
joint shape.  Prints average cycles per key tile (K2) / query chunk (K3) and per CTA."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from merlot_b200 import ops  # noqa: E402
from merlot_b200._lib import lib  # noqa: E402

dev = "cuda"
L = lib()
g = torch.Generator().manual_seed(0)
for B, S, heads, masked in ((32, 266, 12, False), (8, 396, 12, True), (2, 3608, 12, False)):
    H = heads * 64
    qkv = (torch.randn(B * S, 3 * H, generator=g) * 0.5).bfloat16().to(dev)
    dctx = (torch.randn(B * S, H, generator=g) * 0.5).bfloat16().to(dev)
    valid = None
    if masked:
        v = torch.ones(B, S, dtype=torch.uint8)
        v[:, S - 40:] = 0
        valid = v.reshape(-1).to(dev)
    ctx, lse = ops.attention_fwd(qkv, B, S, heads, valid)
    ws = ops.attention_bwd_workspace(B, S, heads, dev)
    dqkv = torch.empty(B * S, 3 * H, dtype=torch.bfloat16, device=dev)
    dsum = torch.empty(B, heads, S, dtype=torch.float32, device=dev)
    ops.attention_bwd(qkv, ctx, dctx, lse, B, S, heads, valid, dqkv=dqkv, dq_accum=ws, dsum=dsum)
    torch.cuda.synchronize()
    cnt = torch.zeros(24, dtype=torch.int64, device=dev)
    L.merlot_attention_debug_counters(C.c_void_p(cnt.data_ptr()))
    ops.attention_fwd(qkv, B, S, heads, valid, ctx=ctx, lse=lse)
    ops.attention_bwd(qkv, ctx, dctx, lse, B, S, heads, valid, dqkv=dqkv, dq_accum=ws, dsum=dsum)
    torch.cuda.synchronize()
    L.merlot_attention_debug_counters(None)
    c = cnt.cpu().tolist()
    nf = (S + 127) // 128 * heads * B
    nb = (S + 127) // 128 * heads * B
    tf, tb = max(c[6], 1), max(c[14], 1)
    print(f"B={B} S={S} masked={masked}")
    print("  K2 per key tile : " + "  ".join(f"{n} {c[i] / tf:7.0f}" for i, n in enumerate(("waitS", "ld+max", "rescale", "exp+stP", "fence+sync"))) +
          f"   | per CTA: final wait {c[5] / nf:7.0f}  epilogue {c[7] / nf:7.0f}  tiles/CTA {tf / nf:.1f}")
    print("  K3 per q chunk  : " + "  ".join(f"{n} {c[8 + i] / tb:7.0f}" for i, n in enumerate(("wait1", "math", "fence+sync", "wait2", "dQ out", "sync2"))) +
          f"   | per CTA: epilogue {c[15] / nb:7.0f}  chunks/CTA {tb / nb:.1f}")
    ti = max(c[22], 1)
    print("  K3 issuer/chunk : " + "  ".join(f"{n} {c[16 + i] / ti:7.0f}" for i, n in enumerate(("wait P", "MMA2 issue", "stats", "MMA1 (+wait Q)", "refill"))))

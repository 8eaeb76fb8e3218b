"""Bitwise repeatability of individual ops on identical inputs (back-to-back launches, PDL active)."""
import sys

import torch

sys.path.insert(0, ".")
from merlot_b200 import ops  # noqa: E402

dev = "cuda"
g = torch.Generator().manual_seed(0)


def rep(name, fn, n=6):
    outs = []
    for _ in range(n):
        o = fn()
        torch.cuda.synchronize()
        outs.append([t.clone() for t in (o if isinstance(o, (list, tuple)) else [o])])
    bad = []
    for k in range(len(outs[0])):
        d = max(((outs[i][k].float() - outs[0][k].float()).norm() / (outs[0][k].float().norm() + 1e-30)).item() for i in range(1, n))
        bad.append(d)
    print(f"{name}: max rel diff across {n} runs per output: {['%.2e' % b for b in bad]}", flush=True)


M, H, I = 8512, 768, 3072
x = (torch.randn(M, H, generator=g) * 0.5).bfloat16().to(dev)
w = (torch.randn(H, I, generator=g) * 0.05).bfloat16().to(dev)
bias = torch.randn(I, generator=g).to(dev)
pre = torch.empty(M, I, dtype=torch.bfloat16, device=dev)
rep("gemm fwd gelu dual", lambda: [ops.gemm(x, w, b_mn_major=True, bias=bias, gelu=True, out_pre=pre), pre])
dy = (torch.randn(M, I, generator=g) * 0.1).bfloat16().to(dev)
rep("gemm dgrad", lambda: ops.gemm(dy, w, out_dtype=torch.bfloat16, M=M, N=H, K=I))
rep("gemm dgrad dgelu", lambda: ops.gemm(x, w.t().contiguous(), out_dtype=torch.bfloat16, dgelu_aux=dy, M=M, N=I, K=H))
rep("gemm wgrad splitK (fp32 red)", lambda: ops.gemm(x, dy, a_mn_major=True, b_mn_major=True, out=torch.zeros(H, I, device=dev), atomic=True, M=H, N=I, K=M))
B, S, heads = 32, 266, 12
qkv = torch.randn(B * S, 3 * H, generator=g).bfloat16().to(dev)
rep("attn fwd", lambda: list(ops.attention_fwd(qkv, B, S, heads)))
ctx, lse = ops.attention_fwd(qkv, B, S, heads)
dctx = (torch.randn(B * S, H, generator=g) * 0.1).bfloat16().to(dev)


def bwd():
    d = ops.attention_bwd(qkv, ctx, dctx, lse, B, S, heads)
    return [d[:, :H], d[:, H:2 * H], d[:, 2 * H:]]


rep("attn bwd (dq, dk, dv)", bwd)
gam, bet = torch.randn(H, generator=g).to(dev), torch.randn(H, generator=g).to(dev)
y = torch.empty(M, H, dtype=torch.bfloat16, device=dev)
mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
rep("ln fwd", lambda: [ops.layernorm_fwd(x, y, gam, bet, mean, rstd), mean, rstd])

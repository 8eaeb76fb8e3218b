"""One ViT-size GEMM a few times (for ncu source-level captures): python tools/one_gemm.py {qkv|ffn1|ffn1gelu|dgelu}"""
import sys

import torch

sys.path.insert(0, ".")
from merlot_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "qkv"
dev = "cuda"
g = torch.Generator().manual_seed(0)
M, H, I = 8512, 768, 3072
x = (torch.randn(M, H, generator=g) * 0.5).bfloat16().to(dev)
wqkv = (torch.randn(H, 3 * H, generator=g) * 0.05).bfloat16().to(dev)
w1 = (torch.randn(H, I, generator=g) * 0.05).bfloat16().to(dev)
w2 = (torch.randn(I, H, generator=g) * 0.05).bfloat16().to(dev)
b1 = torch.randn(I).to(dev)
oqkv = torch.empty(M, 3 * H, dtype=torch.bfloat16, device=dev)
oi = torch.empty(M, I, dtype=torch.bfloat16, device=dev)
oi2 = torch.randn(M, I, generator=g).bfloat16().to(dev)
for _ in range(4):
    if which == "qkv":
        ops.gemm(x, wqkv, b_mn_major=True, out=oqkv)
    elif which == "ffn1":
        ops.gemm(x, w1, b_mn_major=True, out=oi)
    elif which == "ffn1gelu":
        ops.gemm(x, w1, b_mn_major=True, bias=b1, gelu=True, out_pre=oi2, out=oi)
    else:
        ops.gemm(x, w2, out=oi, dgelu_aux=oi2, M=M, N=I, K=H)
torch.cuda.synchronize()

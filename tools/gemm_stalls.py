"""Per-CTA stall breakdown of the 1-CTA K1 kernel (merlot_gemm_debug_counters) on the ViT-size GEMMs."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from merlot_b200 import ops  # noqa: E402
from merlot_b200._lib import lib  # noqa: E402

dev = "cuda"
g = torch.Generator().manual_seed(0)
M, H, I = 8512, 768, 3072
L = lib()
L.merlot_gemm_debug_counters.argtypes = [ctypes.c_void_p]
L.merlot_gemm_debug_counters.restype = None
buf = torch.zeros(148 * 8, dtype=torch.int64, device=dev)

x = (torch.randn(M, H, generator=g) * 0.5).bfloat16().to(dev)
xi = (torch.randn(M, I, generator=g) * 0.5).bfloat16().to(dev)
wqkv = (torch.randn(H, 3 * H, generator=g) * 0.05).bfloat16().to(dev)
w1 = (torch.randn(H, I, generator=g) * 0.05).bfloat16().to(dev)
w2 = (torch.randn(I, H, generator=g) * 0.05).bfloat16().to(dev)
b3, b1, bh = torch.randn(3 * H).to(dev), torch.randn(I).to(dev), torch.randn(H).to(dev)
oqkv = torch.empty(M, 3 * H, dtype=torch.bfloat16, device=dev)
oh = torch.empty(M, H, dtype=torch.bfloat16, device=dev)
oi = torch.empty(M, I, dtype=torch.bfloat16, device=dev)
oi2 = torch.empty(M, I, dtype=torch.bfloat16, device=dev)
big = (torch.randn(8192, 8192, generator=g) * 0.1).bfloat16().to(dev)
obig = torch.empty(8192, 8192, dtype=torch.bfloat16, device=dev)


def probe(name, fn, k=None):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    buf.zero_()
    L.merlot_gemm_debug_counters(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    L.merlot_gemm_debug_counters(None)
    c = buf.view(148, 8).double().cpu()
    busy = c[:, 7] > 0
    c = c[busy]
    tiles = c[:, 7]
    mx = c[c[:, 7] == tiles.max()]  # CTAs on the critical path (most tiles)
    f = lambda t: f"{t.mean().item():9.0f}"
    print(f"{name:34s} tiles/CTA {tiles.min().item():.0f}-{tiles.max().item():.0f} | total{f(mx[:,0])} cyc/tile{f(mx[:,0]/mx[:,7])}"
          f" | mma: wait_full{f(mx[:,1])} wait_tmem_empty{f(mx[:,2])} | tma: wait_empty{f(mx[:,3])}"
          f" | epi: wait_tmem_full{f(mx[:,4])} wait_staging{f(mx[:,5])} work{f(mx[:,6])}"
          + (f" | mma-warp busy per MMA {((mx[:,0] - mx[:,1] - mx[:,2]) / (mx[:,7] * (k / 16))).mean().item():6.1f} cyc" if k else ""), flush=True)


for bn in (128, 192, 256):  # in-situ cost of one 128 x BN x 16 MMA (isolated floors: 64 / 96 / 128 cycles, tools/micro/mma_bench*.cu)
    probe(f"QKV plain BN={bn}", lambda: ops.gemm(x, wqkv, b_mn_major=True, out=oqkv, block_n=bn), k=H)
    probe(f"FFN2 plain (K=3072) BN={bn}", lambda: ops.gemm(xi, w2, b_mn_major=True, out=oh, block_n=bn), k=I)
probe("QKV plain", lambda: ops.gemm(x, wqkv, b_mn_major=True, out=oqkv))
probe("QKV +bias", lambda: ops.gemm(x, wqkv, b_mn_major=True, bias=b3, out=oqkv))
probe("FFN1 plain", lambda: ops.gemm(x, w1, b_mn_major=True, out=oi))
probe("FFN1 +bias+gelu dual", lambda: ops.gemm(x, w1, b_mn_major=True, bias=b1, gelu=True, out_pre=oi2, out=oi))
probe("FFN2 plain (K=3072)", lambda: ops.gemm(xi, w2, b_mn_major=True, out=oh))
probe("FFN2-dgrad plain", lambda: ops.gemm(x, w2, out=oi, M=M, N=I, K=H))
probe("FFN2-dgrad dgelu", lambda: ops.gemm(x, w2, out=oi, dgelu_aux=oi2, M=M, N=I, K=H))
probe("8192^3 1-CTA BN=256", lambda: ops.gemm(big, big, out=obig, block_n=256), k=8192)

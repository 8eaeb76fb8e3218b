"""Every hot kernel of the step launched a few times at the configs[1] ViT shapes (M = 8512, H = 768, I = 3072; attention
32 x 266 x 12 heads), in a fixed, printed order -- the command `ncu --set full` is wrapped around (tools/ncu_summary.py turns
the report into profiles/rNN_ncu_kernels.json).  Also prints CUDA-event timings of the same launches when run without ncu.

  ncu --set full --clock-control none --import-source on -o gpurun_out/r02_targets python tools/ncu_targets.py --reps 1
  python tools/ncu_targets.py --reps 20            # plain timings
"""
import argparse
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from merlot_b200 import ops  # noqa: E402
from merlot_b200._lib import check, lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--only", default="")
args = ap.parse_args()
dev = "cuda"
g = torch.Generator().manual_seed(0)
M, H, I = 8512, 768, 3072
L = lib()


def bf(*shape, s=0.5):
    return (torch.randn(*shape, generator=g) * s).bfloat16().to(dev)


x, xi = bf(M, H), bf(M, I)
wqkv, wo, w1, w2 = bf(H, 3 * H, s=0.05), bf(H, H, s=0.05), bf(H, I, s=0.05), bf(I, H, s=0.05)
b3, b1, bh = torch.randn(3 * H).to(dev), torch.randn(I).to(dev), torch.randn(H).to(dev)
oqkv = torch.empty(M, 3 * H, dtype=torch.bfloat16, device=dev)
oh = torch.empty(M, H, dtype=torch.bfloat16, device=dev)
oi, oi2 = torch.empty(M, I, dtype=torch.bfloat16, device=dev), torch.empty(M, I, dtype=torch.bfloat16, device=dev)
gw = torch.zeros(I, H, dtype=torch.float32, device=dev)
gqkv = torch.zeros(H, 3 * H, dtype=torch.float32, device=dev)
dy, dres = bf(M, H), bf(M, H)
dx, dmask, y = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
mean, rstd = torch.randn(M).to(dev), (torch.rand(M) + 0.5).to(dev)
gamma, beta = torch.randn(H).to(dev), torch.randn(H).to(dev)
dgam, dbet, dbias = torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
L.merlot_layernorm_bwd_fused.argtypes = [C.c_void_p] * 11 + [C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_uint64, C.c_uint32, C.c_void_p]
lnws = torch.empty(L.merlot_layernorm_bwd_workspace_bytes(H), dtype=torch.uint8, device=dev)


def lnb():
    check(L.merlot_layernorm_bwd_fused(dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), dres.data_ptr(),
                                       dx.data_ptr(), dmask.data_ptr(), dgam.data_ptr(), dbet.data_ptr(), dbias.data_ptr(),
                                       lnws.data_ptr(), M, H, 0.1, 1, 3, st))


B, S, heads = 32, 266, 12
qkv = bf(B * S, 3 * H)
ctx, lse = ops.attention_fwd(qkv, B, S, heads)
dctx = bf(B * S, H)
dqkv = torch.empty(B * S, 3 * H, dtype=torch.bfloat16, device=dev)
dq_acc = ops.attention_bwd_workspace(B, S, heads, dev)
dsum = torch.empty(B, heads, S, dtype=torch.float32, device=dev)
colsum = torch.zeros(B * S, dtype=torch.float32, device=dev)
NP = 85_000_000  # one 12-layer stack's worth of parameters
p32, g32 = torch.randn(NP, device=dev), torch.randn(NP, device=dev) * 1e-3
m16, v16, pb16 = torch.zeros(NP, dtype=torch.bfloat16, device=dev), torch.zeros(NP, dtype=torch.bfloat16, device=dev), torch.zeros(NP, dtype=torch.bfloat16, device=dev)

F = 2.0 * M
targets = [
    # name, fn, flops, algorithmic bytes
    ("K1 QKV fwd +bias                      8512x2304x768", lambda: ops.gemm(x, wqkv, b_mn_major=True, bias=b3, out=oqkv), F * 3 * H * H, 0),
    ("K1 out-proj +bias+resid+dropout       8512x768x768", lambda: ops.gemm(x, wo, b_mn_major=True, bias=bh, resid=x, out=oh, dropout_p=0.1, dropout_seed=1), F * H * H, 0),
    ("K1 FFN1 +bias+gelu (pre+act)          8512x3072x768", lambda: ops.gemm(x, w1, b_mn_major=True, bias=b1, gelu=True, out_pre=oi2, out=oi), F * I * H, 0),
    ("K1 FFN2 +bias+resid+dropout           8512x768x3072", lambda: ops.gemm(xi, w2, b_mn_major=True, bias=bh, resid=x, out=oh, dropout_p=0.1, dropout_seed=1), F * I * H, 0),
    ("K1 FFN2-dgrad x gelu'(pre)            8512x3072x768", lambda: ops.gemm(x, w2, out=oi, dgelu_aux=oi2, M=M, N=I, K=H), F * I * H, 0),
    ("K1 FFN1-dgrad plain                   8512x768x3072", lambda: ops.gemm(xi, w1, out=oh, M=M, N=H, K=I), F * I * H, 0),
    ("K1 QKV-dgrad plain                    8512x768x2304", lambda: ops.gemm(oqkv, wqkv, out=oh, M=M, N=H, K=3 * H), F * 3 * H * H, 0),
    ("K1 out-proj dgrad plain               8512x768x768", lambda: ops.gemm(x, wo, out=oh, M=M, N=H, K=H), F * H * H, 0),
    ("K1 FFN2 wgrad (split-K, red.add f32)  3072x768x8512", lambda: ops.gemm(xi, dy, a_mn_major=True, b_mn_major=True, out=gw, atomic=True, M=I, N=H, K=M), F * I * H, 0),
    ("K1 QKV wgrad (split-K, red.add f32)   768x2304x8512", lambda: ops.gemm(x, oqkv, a_mn_major=True, b_mn_major=True, out=gqkv, atomic=True, M=H, N=3 * H, K=M), F * 3 * H * H, 0),
    ("K2 attention fwd                      32x266 12 heads", lambda: ops.attention_fwd(qkv, B, S, heads, ctx=ctx, lse=lse), 4.0 * B * heads * S * S * 64, 0),
    ("K3 attention bwd (all its kernels)    32x266 12 heads", lambda: ops.attention_bwd(qkv, ctx, dctx, lse, B, S, heads, dqkv=dqkv, dq_accum=dq_acc, dsum=dsum), 10.0 * B * heads * S * S * 64, 0),
    ("K4 attention colsum                   32x266 12 heads", lambda: ops.attention_colsum(qkv, lse, colsum, B, S, heads), 2.0 * B * heads * S * S * 64, 0),
    ("K5 ln_fwd bf16->bf16                  8512x768", lambda: ops.layernorm_fwd(x, y, gamma, beta, mean, rstd), 0, M * H * 2 * 2),
    ("K5 ln_bwd_fused dres+dropout+bias     8512x768", lnb, 0, M * H * 2 * 5),
    ("K10 adamw (85 M params, bf16 copy, zero grad)", lambda: ops.adamw_step(p32, g32, m16, v16, pb16, NP, 0.9, 0.1, 0.98, 0.02, 1e-6, 1e-4, 0.1, 1.0, True), 0, NP * 26),
]
if args.only:
    targets = [t for t in targets if args.only in t[0]]
for name, fn, fl, by in targets:
    fn()  # warm-up (also under ncu: the SECOND launch of every kernel is the warm one)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    extra = f"{fl / ms / 1e9:8.1f} TFLOP/s" if fl else f"{by / ms / 1e6:8.0f} GB/s"
    print(f"TARGET {name:58s} {ms * 1e3:8.1f} us {extra}", flush=True)

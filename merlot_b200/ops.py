"""Torch-tensor front end of the C-ABI operators (device memory + stream plumbing only; all math is in the .so)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.MerlotError(L.MERLOT_EINVAL, "merlot_b200 operators take CUDA tensors only (no CPU fallback)")


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn_major: bool = False, b_mn_major: bool = False,
         out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16,
         bias: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
         gelu: bool = False, out_pre: Optional[torch.Tensor] = None, dgelu_aux: Optional[torch.Tensor] = None,
         alpha: float = 1.0, atomic: bool = False, dropout_p: float = 0.0, dropout_seed: int = 0,
         dropout_site: int = 0, splits: int = 0, block_n: int = 0,
         M: Optional[int] = None, N: Optional[int] = None, K: Optional[int] = None) -> torch.Tensor:
    """C[M,N] = epilogue(alpha * A @ B^T) on tcgen05 tensor cores (see include/merlot_b200.h, K1).

    a: [M,K] (or [K,M] when a_mn_major); b: [N,K] (or [K,N] when b_mn_major); both bf16, row-major, 2-D.
    With gelu=True and out_pre given: out_pre <- pre-activation, return value <- gelu(pre).
    """
    _require_cuda(a, b, out, bias, resid, out_pre, dgelu_aux)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1
    if M is None:
        M = a.shape[1] if a_mn_major else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn_major else a.shape[1]
    if N is None:
        N = b.shape[1] if b_mn_major else b.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.stride(1) == 1
    g = L.GemmDesc()
    g.M, g.N, g.K = M, N, K
    g.a, g.lda, g.a_mn_major = a.data_ptr(), a.stride(0), int(a_mn_major)
    g.b, g.ldb, g.b_mn_major = b.data_ptr(), b.stride(0), int(b_mn_major)
    flags = 0
    if out.dtype == torch.float32:
        flags |= L.GEMM_OUT_F32
    else:
        assert out.dtype == torch.bfloat16
    if atomic:
        flags |= L.GEMM_ATOMIC
    if gelu:
        flags |= L.GEMM_GELU
    if gelu and out_pre is not None:
        # C-ABI convention: out <- pre, out2 <- gelu(pre)
        g.out, g.ld_out = out_pre.data_ptr(), out_pre.stride(0)
        g.out2, g.ld_out2 = out.data_ptr(), out.stride(0)
    else:
        g.out, g.ld_out = out.data_ptr(), out.stride(0)
        g.out2, g.ld_out2 = None, 0
    if dgelu_aux is not None:
        flags |= L.GEMM_MUL_DGELU
        g.aux, g.ld_aux = dgelu_aux.data_ptr(), dgelu_aux.stride(0)
    if bias is not None:
        assert bias.dtype == torch.float32
        g.bias = bias.data_ptr()
    if resid is not None:
        assert resid.dtype == torch.bfloat16
        g.resid, g.ld_resid = resid.data_ptr(), resid.stride(0)
    if dropout_p > 0.0:
        flags |= L.GEMM_DROPOUT
    g.alpha = alpha
    g.flags = flags
    g.dropout_p, g.dropout_seed, g.dropout_site = dropout_p, dropout_seed, dropout_site
    g.splits, g.block_n = splits, block_n
    L.check(L.lib().merlot_gemm_bf16(C.byref(g), _stream()))
    return out


def _attn_desc(qkv: torch.Tensor, B: int, S: int, heads: int, valid: Optional[torch.Tensor]) -> L.AttnDesc:
    _require_cuda(qkv, valid)
    assert qkv.dtype == torch.bfloat16 and qkv.dim() == 2 and qkv.stride(1) == 1 and qkv.shape[0] == B * S
    a = L.AttnDesc()
    a.B, a.S, a.heads, a.head_dim = B, S, heads, qkv.shape[1] // (3 * heads)
    a.qkv, a.ld_qkv = qkv.data_ptr(), qkv.stride(0)
    if valid is not None:
        assert valid.dtype == torch.uint8 and valid.numel() == B * S and valid.is_contiguous()
        a.valid = valid.data_ptr()
    a.scale = 1.0 / (a.head_dim ** 0.5)
    return a


def attention_fwd(qkv: torch.Tensor, B: int, S: int, heads: int, valid: Optional[torch.Tensor] = None,
                  ctx: Optional[torch.Tensor] = None, lse: Optional[torch.Tensor] = None):
    """K2: ctx[B*S,H] = softmax(mask(q k^T / sqrt(d))) v, reading q/k/v in place from the fused qkv buffer."""
    a = _attn_desc(qkv, B, S, heads, valid)
    H = heads * a.head_dim
    if ctx is None:
        ctx = torch.empty((B * S, H), dtype=torch.bfloat16, device=qkv.device)
    if lse is None:
        lse = torch.empty((B, heads, S), dtype=torch.float32, device=qkv.device)
    a.ctx, a.ld_ctx, a.lse = ctx.data_ptr(), ctx.stride(0), lse.data_ptr()
    L.check(L.lib().merlot_attention_fwd(C.byref(a), _stream()))
    return ctx, lse


def attention_bwd(qkv, ctx, d_ctx, lse, B, S, heads, valid=None, dqkv=None, dq_accum=None, dsum=None):
    """K3: dqkv[B*S,3H] from d_ctx.  dq_accum must be zero on entry (it is returned zeroed)."""
    a = _attn_desc(qkv, B, S, heads, valid)
    H = heads * a.head_dim
    dev = qkv.device
    if dqkv is None:
        dqkv = torch.empty((B * S, 3 * H), dtype=torch.bfloat16, device=dev)
    if dq_accum is None:
        dq_accum = torch.zeros((B * S, H), dtype=torch.float32, device=dev)
    if dsum is None:
        dsum = torch.empty((B, heads, S), dtype=torch.float32, device=dev)
    a.ctx, a.ld_ctx, a.lse = ctx.data_ptr(), ctx.stride(0), lse.data_ptr()
    assert d_ctx.stride(0) == ctx.stride(0)
    a.d_ctx, a.dsum = d_ctx.data_ptr(), dsum.data_ptr()
    a.dq_accum, a.ld_dq = dq_accum.data_ptr(), dq_accum.stride(0)
    a.dqkv, a.ld_dqkv = dqkv.data_ptr(), dqkv.stride(0)
    L.check(L.lib().merlot_attention_bwd(C.byref(a), _stream()))
    return dqkv


def attention_colsum(qkv, lse, colsum, B, S, heads, valid=None):
    """K4: colsum[B,S] += mean_h sum_q P[b,h,q,k] (recomputed from q,k,lse)."""
    a = _attn_desc(qkv, B, S, heads, valid)
    assert colsum.dtype == torch.float32 and colsum.numel() == B * S
    a.lse, a.colsum = lse.data_ptr(), colsum.data_ptr()
    L.check(L.lib().merlot_attention_colsum(C.byref(a), _stream()))
    return colsum

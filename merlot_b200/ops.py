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

"""VCR downstream pieces: the classifier head of downstream/vcr/modeling.py over MerlotModel(num_texts=4).

`MerlotModel(config with num_texts: 4, image=[b, h, w, 3], input_ids=[b*4, L])` tiles every image's tokens to its four
candidate texts (model/modeling.py:111-119; merlot_b200/modeling.py does that with four LayerNorm row remaps).  This module adds
`cls_head_val` (downstream/vcr/modeling.py:57-77): first language token -> dense(H/2)+gelu -> dense(1) -> [img_batch, 4] logits,
and `cls_loss` (softmax cross entropy over the four candidates).  All arithmetic runs in libmerlot_b200.so (K1 GEMM + fused
bias, the erf-GeLU kernel, the CE kernel); the head's variables live in a small dict keyed by the reference's names
(`<mode>_cls/classifier_mlp{0,1}/{kernel,bias}`), e.g. filled from a checkpoint by ParamStore-independent loading.
The TRAINING head (`cls_head`, :79-127, answer + rationale towers with dropout) is the same two layers per tower; its backward
enters the model through MerlotModel.backward(d_hidden_state=...).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import ops


def init_head(hidden_size: int, mode: str = "answer", initializer_range: float = 0.02, bias_pi: float = 0.25, seed: int = 0,
              device="cuda") -> Dict[str, torch.Tensor]:
    """Reference initialisers (downstream/vcr/modeling.py:63-74): truncated normal kernels, bias1 = -log((1-pi)/pi)."""
    g = torch.Generator().manual_seed(seed)
    k0 = torch.empty(hidden_size, hidden_size // 2)
    k1 = torch.empty(hidden_size // 2, 1)
    for t in (k0, k1):
        torch.nn.init.trunc_normal_(t, 0.0, initializer_range, -2 * initializer_range, 2 * initializer_range, generator=g)
    p = {f"{mode}_cls/classifier_mlp0/kernel": k0, f"{mode}_cls/classifier_mlp0/bias": torch.zeros(hidden_size // 2),
         f"{mode}_cls/classifier_mlp1/kernel": k1, f"{mode}_cls/classifier_mlp1/bias": torch.full((1,), -math.log((1 - bias_pi) / bias_pi))}
    return {k: v.to(device) for k, v in p.items()}


def cls_head_val(model, head: Dict[str, torch.Tensor], mode: str = "answer") -> torch.Tensor:
    """downstream/vcr/modeling.py:57-77 on model.encoder_hidden_states['lang'] -> fp32 logits [img_batch_size, 4]."""
    y = model.encoder_info["hidden_state"]  # bf16 [B, P+L, H]; encoder_hidden_states['lang'][:, 0] is its row P (cast to fp32 there)
    B, Sj, H = y.shape
    if B % 4 != 0:
        raise ValueError(f"VCR heads score 4 candidates per image: batch {B} is not a multiple of 4")
    dev = y.device
    first = torch.empty((B, H), dtype=torch.bfloat16, device=dev)
    idx = (torch.arange(B, device=dev, dtype=torch.int32) * Sj + model.P).contiguous()
    ops.gather_rows(y.reshape(B * Sj, H), idx, first)  # hidden_state[:, 0, :] of the language piece
    k0 = head[f"{mode}_cls/classifier_mlp0/kernel"].to(torch.bfloat16).contiguous()
    pre = ops.gemm(first, k0, b_mn_major=True, bias=head[f"{mode}_cls/classifier_mlp0/bias"].float().contiguous(), out_dtype=torch.float32)
    act = torch.empty_like(pre)
    ops.gelu_f32(pre, act)
    actb = torch.empty(pre.shape, dtype=torch.bfloat16, device=dev)
    ops.cast_f32_to_bf16(act, actb)
    k1 = torch.zeros((H // 2, 8), dtype=torch.bfloat16, device=dev)  # N = 1 padded to a TMA-legal row of 8
    k1[:, :1] = head[f"{mode}_cls/classifier_mlp1/kernel"].to(torch.bfloat16)
    b1 = torch.zeros(8, dtype=torch.float32, device=dev)
    b1[:1] = head[f"{mode}_cls/classifier_mlp1/bias"].float()
    logits = ops.gemm(actb, k1, b_mn_major=True, bias=b1, out_dtype=torch.float32)  # [B, 8], column 0 is the logit
    return logits[:, 0].reshape(B // 4, 4)


def cls_loss(logits_flat: torch.Tensor, target: torch.Tensor):
    """downstream/vcr/modeling.py:133-150: mean softmax cross entropy over the 4 candidates + accuracy (0-d CUDA tensors)."""
    n = logits_flat.shape[0]
    dev = logits_flat.device
    padded = torch.zeros((n, 8), dtype=torch.float32, device=dev)
    padded[:, :4] = logits_flat
    per, lse, corr = (torch.empty(n, dtype=torch.float32, device=dev) for _ in range(3))
    ops.softmax_ce_fwd(padded, target.to(torch.int32).contiguous(), 4, per, lse, corr)
    out2 = torch.empty(2, dtype=torch.float32, device=dev)
    coeff = torch.empty(n, dtype=torch.float32, device=dev)
    ops.weighted_loss(per, corr, None, None, 0, 1.0, out2, coeff)
    return out2[0], out2[1]

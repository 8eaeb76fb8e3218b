"""Batch-level input step of the reference dataloader (`_process_example`, model/dataloader.py:210-272) -- SURVEY.md 8(f)
next-row 4.  Runs on whatever device the feature tensors live on (plain index arithmetic: argsort / gather / where), so
the step can stay on the GPU next to the model instead of in a tf.data map on the host.

  * shuffle_chunks (:211-224)   -- permute whole source videos inside a row: [A A A A B B] -> [B B A A A A]
  * frame shuffle (:226-257)    -- per group of `num_chunks_in_group` segments draw how many frames are out of place
                                   (categorical over [1-p, 1e-6, p/(n-1), ...]), pick that many positions, give them the ids
                                   16 + random order; everything else keeps arange(n).  The offset 16 (< 64) is why every
                                   temporal pair weighs 0.01 in pretraining.
  * flatten + transpose (:259-264) -- images [b, n, h, w, 3] -> [b*n, h, w, 3] (-> [h, w, 3, b*n] when training with
                                   `transpose_input`, undone by model_fn, modeling.py:683-685)

The reference draws its random numbers with tf.random.*; `draws` injects them so the integer results can be compared bit for
bit with the NumPy restatement in oracle/oracle_np.py (tests/test_host_logic.py).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

SHUFFLE_OFFSET = 16  # model/dataloader.py:227
_PER_CHUNK_KEYS = ("youtube_id", "chunk_num", "mean_time", "images", "input_ids", "is_eoc", "video_src_ids")  # :222


def num_shuffle_probs(num_chunks_in_group: int, shuffle_prob: float):
    """:243-244 -- P(k frames out of place) for k = 0..n: [1-p, 1e-6, p/(n-1) x (n-1)] (one frame out of place is no shuffle)."""
    n = num_chunks_in_group
    return [1.0 - shuffle_prob, 1e-6] + [shuffle_prob / (n - 1) for _ in range(n - 1)]


def make_draws(batch_size: int, num_chunks: int, num_chunks_in_group: int, shuffle_prob: float, seed: int, device="cpu"):
    """The four random tensors `_process_example` draws, from a seed (tf.random_uniform x3, tf.random.categorical x1)."""
    g = torch.Generator().manual_seed(seed)
    B = batch_size * num_chunks // num_chunks_in_group
    probs = torch.tensor(num_shuffle_probs(num_chunks_in_group, max(shuffle_prob, 1e-6)), dtype=torch.float64)
    d = {
        "chunk_u": torch.rand(batch_size, num_chunks, generator=g),
        "num_shuffle": torch.multinomial(probs / probs.sum(), B, True, generator=g).to(torch.int32),
        "pick_u": torch.rand(B, num_chunks_in_group, generator=g),
        "order_u": torch.rand(B, num_chunks_in_group, generator=g),
    }
    return {k: v.to(device) for k, v in d.items()}


def process_example(features: Dict[str, torch.Tensor], data_cfg: dict, model_cfg: dict, is_training: bool = True,
                    draws: Optional[Dict[str, torch.Tensor]] = None, seed: int = 0) -> Dict[str, torch.Tensor]:
    """model/dataloader.py:210-272.  `features`: images [b, n, h, w, 3], input_ids [b, n, L], video_src_ids [b, n] (+ any of the
    other per-chunk keys).  Returns a new dict with `shuffled_idx_img` [B*g] int32 added and images flattened."""
    out = dict(features)
    ids = out["input_ids"]
    if ids.dim() != 3:
        raise ValueError(f"input_ids must be [batch, num_chunks, L], got {tuple(ids.shape)}")
    b, n, _ = ids.shape
    g = model_cfg.get("num_chunks_in_group", n)
    if (b * n) % g != 0:
        raise ValueError(f"batch*num_chunks = {b * n} is not divisible by num_chunks_in_group = {g}")
    B = b * n // g
    p = model_cfg.get("image_shuffle_prob", 0.5)  # :226 (merged data+model config; the key lives in `model:`)
    dev = ids.device
    if draws is None:
        draws = make_draws(b, n, g, p, seed, dev)

    if data_cfg.get("shuffle_chunks", False):  # :211-224
        vid = out["video_src_ids"].long()
        mapping = torch.argsort(draws["chunk_u"].to(dev), dim=-1, stable=True)
        new_chunkid = torch.gather(mapping, 1, vid)
        trg = new_chunkid * n + torch.arange(n, device=dev)[None]
        idx = torch.argsort(trg, dim=1, stable=True)
        for k in _PER_CHUNK_KEYS:
            if k in out:
                v = out[k]
                ix = idx.reshape(b, n, *([1] * (v.dim() - 2))).expand(b, n, *v.shape[2:])
                out[k] = torch.gather(v, 1, ix)

    base = torch.arange(g, dtype=torch.int32, device=dev)[None].expand(B, g)
    if p < 1e-6:  # :234-237
        shuffled = base
    else:  # :238-257
        num_shuffle = draws["num_shuffle"].to(dev).to(torch.int32)
        do_shuffle = torch.argsort(draws["pick_u"].to(dev), dim=1, stable=True).to(torch.int32) < num_shuffle[:, None]
        order = torch.argsort(draws["order_u"].to(dev), dim=1, stable=True).to(torch.int32)
        shuffled = torch.where(do_shuffle, SHUFFLE_OFFSET + order, base)
    out["shuffled_idx_img"] = shuffled.reshape(-1).contiguous()

    img = out["images"]
    if img.dim() != 5 or img.shape[-1] != 3:
        raise ValueError(f"images must be [batch, num_chunks, h, w, 3], got {tuple(img.shape)}")
    img = img.reshape(b * img.shape[1], *img.shape[2:])  # :260-261
    if is_training and model_cfg.get("transpose_input", False):  # :262-264
        img = img.permute(1, 2, 3, 0)
    out["images"] = img
    return out


def expected_out_of_place(num_chunks_in_group: int, shuffle_prob: float) -> float:
    """The figure the reference logs (:245-247)."""
    return math.fsum(i * q for i, q in enumerate(num_shuffle_probs(num_chunks_in_group, shuffle_prob)))

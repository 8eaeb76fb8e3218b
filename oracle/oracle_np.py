"""ORACLE (second, independent restatement) -- TEST INFRASTRUCTURE ONLY.  NumPy fp64 versions of the hot-path primitives,
written separately from oracle/merlot_oracle.py so the two can check each other (SURVEY.md 8(c) "oracle cross-check").
PARITY UNPINNED against the reference itself (no reference tests/fixtures exist; TF 1.15 cannot run here).
Citations relative to /root/reference."""
import math

import numpy as np


def erf(x):
    """High-precision erf via math.erf (vectorised)."""
    return np.vectorize(math.erf)(x)


def gelu(x):  # utils/model_utils.py:96-110
    return x * 0.5 * (1.0 + erf(x / math.sqrt(2.0)))


def layer_norm(x, gamma, beta, eps=1e-5):  # utils/model_utils.py:113-130
    mean = x.mean(-1, keepdims=True)
    var = x.var(-1, keepdims=True)  # biased
    s = gamma / np.sqrt(var + eps)
    return x * s - mean * s + beta


def softmax(x):
    e = np.exp(x - x.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


def attention(x, wq, bq, wk, bk, wv, bv, wo, bo, mask, heads):  # utils/transformer.py:33-138
    B, S, H = x.shape
    d = H // heads
    xf = x.reshape(B * S, H)

    def proj(w, b):
        return (xf @ w + b).reshape(B, S, heads, d).transpose(0, 2, 1, 3)

    q, k, v = proj(wq, bq), proj(wk, bk), proj(wv, bv)
    s = q @ k.transpose(0, 1, 3, 2) / math.sqrt(d)
    m = mask[:, None]
    s = s * m - 1e10 * (1 - m)
    p = softmax(s)
    ctx = (p @ v).transpose(0, 2, 1, 3).reshape(B * S, H)
    return (ctx @ wo + bo).reshape(B, S, H), p


def cross_entropy(logits, labels):  # utils/model_utils.py:313-332
    z = logits - logits.max(-1, keepdims=True)
    logp = z - np.log(np.exp(z).sum(-1, keepdims=True))
    return -np.take_along_axis(logp, labels[..., None], -1)[..., 0]


def bf16_round(x):
    """float32 -> nearest-even bfloat16 -> float32, bit-level."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32)


def decode_v(v_bf16_as_f32):  # utils/optimization.py:268-281
    a = np.abs(v_bf16_as_f32).astype(np.float32)
    return np.where(np.sign(v_bf16_as_f32) > 0, a, a * np.float32(1.00390625)).astype(np.float32)


def encode_v(v):  # utils/optimization.py:283-288
    e = bf16_round(v)
    err0 = np.abs(e - v)
    err1 = np.abs(e * np.float32(1.00390625) - v)
    return np.where(err0 <= err1, e, -e).astype(np.float32)


# ---- hybrid ResNet-lite stem pieces, written independently of merlot_oracle.py --------------------
def group_norm(x, gamma, beta, num_groups=32, eps=1e-4):  # utils/model_utils.py:133-222, mean_close_to_zero=True
    n, h, w, c = x.shape
    g = c // num_groups
    out = np.empty_like(x)
    for b in range(n):
        for k in range(num_groups):
            blk = x[b, :, :, k * g:(k + 1) * g]
            m = blk.sum() / blk.size
            v = (blk * blk).sum() / blk.size - m * m  # one-pass variance (sufficient_statistics / normalize_moments)
            out[b, :, :, k * g:(k + 1) * g] = (blk - m) / np.sqrt(v + eps)
    return out * gamma + beta


def conv2d_ws(x, kernel, strides=1, ws=True):  # utils/vision_transformer.py:30-66 (NHWC, HWIO), direct loops
    kh, kw, cin, cout = kernel.shape
    if ws:
        flat = kernel.reshape(-1, cout)
        kernel = ((flat - flat.mean(0)) / np.sqrt(flat.var(0) + 1e-5)).reshape(kernel.shape)
    n, h, w, _ = x.shape
    if strides > 1:
        beg = (kh - 1) // 2
        end = kh - 1 - beg
    else:
        beg = end = kh // 2
    xp = np.pad(x, ((0, 0), (beg, end), (beg, end), (0, 0)))
    ho = (h + beg + end - kh) // strides + 1
    wo = (w + beg + end - kw) // strides + 1
    out = np.zeros((n, ho, wo, cout), dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            patch = xp[:, i:i + (ho - 1) * strides + 1:strides, j:j + (wo - 1) * strides + 1:strides, :]
            out += patch @ kernel[i, j]
    return out


def avg_pool_same(x, s):  # tf.nn.avg_pool2d SAME, ksize = strides = s: bottom/right padding, padded cells not counted
    n, h, w, c = x.shape
    ho, wo = -(-h // s), -(-w // s)
    out = np.empty((n, ho, wo, c), dtype=x.dtype)
    for i in range(ho):
        for j in range(wo):
            out[:, i, j] = x[:, i * s:min(h, (i + 1) * s), j * s:min(w, (j + 1) * s)].mean((1, 2))
    return out


# ---- batch-level input step (model/dataloader.py:210-272), written independently of merlot_b200/dataloader.py ---------------
def process_example_np(input_ids, video_src_ids, chunk_u, num_shuffle, pick_u, order_u, num_chunks_in_group, shuffle_prob,
                       shuffle_chunks):
    """Returns (permutation idx [b, n] applied to every per-chunk feature, shuffled_idx_img [B*g]).  tf.argsort of distinct
    values == np.argsort(kind='stable')."""
    b, n = video_src_ids.shape
    idx = np.tile(np.arange(n), (b, 1))
    if shuffle_chunks:
        idx = np.empty((b, n), dtype=np.int64)
        for r in range(b):
            mapping = np.argsort(chunk_u[r], kind="stable")          # chunkid_to_new_id_mapping (:216)
            new_chunkid = mapping[video_src_ids[r]]                   # (:217)
            trg = new_chunkid * n + np.arange(n)                      # (:218)
            idx[r] = np.argsort(trg, kind="stable")                   # (:219)
    g = num_chunks_in_group
    B = b * n // g
    out = np.tile(np.arange(g, dtype=np.int32), (B, 1))
    if shuffle_prob >= 1e-6:
        for r in range(B):
            rank = np.argsort(pick_u[r], kind="stable")               # (:251): argsort VALUES compared with the count
            order = np.argsort(order_u[r], kind="stable")
            for j in range(g):
                if rank[j] < num_shuffle[r]:
                    out[r, j] = 16 + order[j]
    return idx, out.reshape(-1)

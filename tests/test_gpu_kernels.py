"""GPU parity tests (run on the B200 box: pytest -m gpu): every CUDA kernel family against the oracle on the same
seeded inputs, through the C-ABI.  Tolerances: bit-exact for integer/index work and the packed bf16 Adam moments;
bf16 tensor-core outputs compared in relative Frobenius norm against the fp32 oracle evaluated on the SAME bf16-rounded
inputs (GEMM fp32-out 1e-4; bf16-out / attention 1e-2 -- one bf16 rounding is 2^-9 = 2e-3 per element)."""
import numpy as np
import pytest
import torch

from oracle import merlot_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda"


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from merlot_b200 import ops as o
    return o


# ---------------------------------------------------------------------------------------------------------------
# K1 GEMM: all operand-major modes, ragged shapes (every Appendix-B family incl. N = 50370 and M = 200)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K,bn", [(128, 256, 64, 256), (200, 264, 136, 0), (1024, 2304, 768, 0), (3168, 768, 3072, 128),
                                      (8, 8, 8, 0), (130, 50376, 768, 0)])
def test_gemm_modes(ops, a_mn, b_mn, M, N, K, bn):
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("MN-major operands need 16-byte aligned rows")
    g = torch.Generator().manual_seed(M + N + K)
    a = (torch.randn((K, M) if a_mn else (M, K), generator=g) * 0.5).bfloat16()
    b = (torch.randn((K, N) if b_mn else (N, K), generator=g) * 0.5).bfloat16()
    out = ops.gemm(a.to(DEV), b.to(DEV), a_mn_major=a_mn, b_mn_major=b_mn, out_dtype=torch.float32, block_n=bn)
    A = a.float().t() if a_mn else a.float()
    Bm = b.float() if b_mn else b.float().t()
    assert rel(out, A @ Bm) < 1e-4


def test_gemm_epilogues(ops):
    g = torch.Generator().manual_seed(1)
    M, N, K = 520, 768, 768
    a = (torch.randn(M, K, generator=g) * 0.3).bfloat16()
    w = (torch.randn(K, N, generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, generator=g)
    resid = torch.randn(M, N, generator=g).bfloat16()
    p = {"d/kernel": w.float(), "d/bias": bias}
    base = O.dense(a.float(), p, "d")  # tf.layers.dense restatement
    ad, wd, bd, rd = a.to(DEV), w.to(DEV), bias.to(DEV), resid.to(DEV)
    assert rel(ops.gemm(ad, wd, b_mn_major=True, bias=bd), base) < 6e-3
    assert rel(ops.gemm(ad, wd, b_mn_major=True, bias=bd, resid=rd), base + resid.float()) < 6e-3
    pre = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    act = ops.gemm(ad, wd, b_mn_major=True, bias=bd, gelu=True, out_pre=pre)
    assert rel(pre, base) < 6e-3 and rel(act, O.gelu(base)) < 6e-3
    aux = torch.randn(M, N, generator=g).bfloat16()
    x = aux.float().requires_grad_(True)
    O.gelu(x).sum().backward()
    assert rel(ops.gemm(ad, wd, b_mn_major=True, dgelu_aux=aux.to(DEV)), (a.float() @ w.float()) * x.grad) < 6e-3
    # the pair the stacks use: the forward saves gelu'(pre) (GELU_GRAD_OUT), the FFN2 dgrad multiplies by it (MUL_AUX)
    for kw in (dict(), dict(block_n=128), dict(block_n=256)):
        gsave = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        act2 = ops.gemm(ad, wd, b_mn_major=True, bias=bd, gelu=True, out_pre=gsave, gelu_grad_out=True, **kw)
        xb = base.clone().requires_grad_(True)
        O.gelu(xb).sum().backward()
        assert rel(act2, O.gelu(base)) < 6e-3 and rel(gsave, xb.grad) < 6e-3
        assert rel(ops.gemm(ad, wd, b_mn_major=True, mul_aux=gsave, **kw), (a.float() @ w.float()) * gsave.float().cpu()) < 6e-3
    dy = (torch.randn(M, N, generator=g) * 0.1).bfloat16()
    dw = torch.zeros(K, N, dtype=torch.float32, device=DEV)
    for _ in range(2):  # accumulation semantics of the flat gradient arena (shared `encoder` weights get two passes)
        ops.gemm(ad, dy.to(DEV), a_mn_major=True, b_mn_major=True, out=dw, atomic=True, M=K, N=N, K=M)
    assert rel(dw, 2 * (a.float().t() @ dy.float())) < 1e-4
    o1 = ops.gemm(ad, wd, b_mn_major=True, bias=bd, dropout_p=0.1, dropout_seed=7, dropout_site=3)
    o2 = ops.gemm(ad, wd, b_mn_major=True, bias=bd, dropout_p=0.1, dropout_seed=7, dropout_site=3)
    o3 = ops.gemm(ad, wd, b_mn_major=True, bias=bd, dropout_p=0.1, dropout_seed=8, dropout_site=3)
    assert torch.equal(o1, o2) and not torch.equal(o1, o3)
    assert abs((o1 == 0).float().mean().item() - 0.1) < 0.01
    kept = (o1 != 0).cpu()
    assert rel(o1.cpu()[kept], (base / 0.9)[kept]) < 6e-3  # inverted dropout scaling (tf.nn.dropout)


@pytest.mark.parametrize("bn", [128, 192, 256, -256])  # -256 = CTA-pair kernel
@pytest.mark.parametrize("M,N", [(300, 264), (130, 1000), (515, 72)])
def test_gemm_epilogue_instances(ops, bn, M, N):
    """Every feature-specialised epilogue instance (plain, bias, bias+resid(+generic), bias+gelu dual, gelu', resid) on
    every tile width, with ragged M and N edges (N % 64 != 0, N % 32 != 0) -- the warp-private staged epilogue clips per
    16-byte chunk and per row."""
    if bn == -256 and M <= 256:
        pytest.skip("the pair kernel needs more than one 256-row tile to be selected")
    g = torch.Generator().manual_seed(M * 7 + N)
    K = 200
    a = (torch.randn(M, K, generator=g) * 0.3).bfloat16()
    w = (torch.randn(K, N, generator=g) * 0.1).bfloat16()
    bias = torch.randn(N, generator=g)
    resid = torch.randn(M, N, generator=g).bfloat16()
    aux = torch.randn(M, N, generator=g).bfloat16()
    base = a.float() @ w.float()
    ad, wd, bd, rd, xd = a.to(DEV), w.to(DEV), bias.to(DEV), resid.to(DEV), aux.to(DEV)
    kw = dict(b_mn_major=True, block_n=bn)
    assert rel(ops.gemm(ad, wd, **kw), base) < 6e-3
    assert rel(ops.gemm(ad, wd, bias=bd, **kw), base + bias) < 6e-3
    assert rel(ops.gemm(ad, wd, resid=rd, **kw), base + resid.float()) < 6e-3
    assert rel(ops.gemm(ad, wd, bias=bd, resid=rd, **kw), base + bias + resid.float()) < 6e-3
    assert rel(ops.gemm(ad, wd, bias=bd, resid=rd, alpha=0.5, **kw), 0.5 * base + bias + resid.float()) < 6e-3  # generic instance
    pre = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    act = ops.gemm(ad, wd, bias=bd, gelu=True, out_pre=pre, **kw)
    assert rel(pre, base + bias) < 6e-3 and rel(act, O.gelu(base + bias)) < 6e-3
    x = aux.float().requires_grad_(True)
    O.gelu(x).sum().backward()
    assert rel(ops.gemm(ad, wd, dgelu_aux=xd, **kw), base * x.grad) < 6e-3
    assert rel(ops.gemm(ad, wd, dgelu_aux=xd, resid=rd, **kw), base * x.grad + resid.float()) < 6e-3  # gelu' + unprefetched residual
    dw = torch.zeros(K, N, dtype=torch.float32, device=DEV)  # split-K fp32 accumulation (EPI 2) with ragged N
    dy = (torch.randn(M, N, generator=g) * 0.1).bfloat16()
    ops.gemm(ad, dy.to(DEV), a_mn_major=True, b_mn_major=True, out=dw, atomic=True, M=K, N=N, K=M, block_n=bn if bn > 0 else 0)
    assert rel(dw, a.float().t() @ dy.float()) < 1e-4


@pytest.mark.parametrize("M,N,K", [(300, 768, 200), (515, 264, 136), (1000, 72, 2304)])
def test_gemm_pair192(ops, M, N, K):
    """256 x 192 CTA-pair tile (block_n = -192): K-major B only (the dgrad orientation), every epilogue family, ragged edges."""
    from merlot_b200._lib import MerlotError
    g = torch.Generator().manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=g) * 0.3).bfloat16()
    b = (torch.randn(N, K, generator=g) * 0.1).bfloat16()  # [N, K]: K-major
    aux = torch.randn(M, N, generator=g).bfloat16()
    resid = torch.randn(M, N, generator=g).bfloat16()
    bias = torch.randn(N, generator=g)
    base = a.float() @ b.float().t()
    ad, bd = a.to(DEV), b.to(DEV)
    assert rel(ops.gemm(ad, bd, block_n=-192), base) < 6e-3
    assert rel(ops.gemm(ad, bd, block_n=-192, out_dtype=torch.float32), base) < 1e-4
    assert rel(ops.gemm(ad, bd, block_n=-192, bias=bias.to(DEV), resid=resid.to(DEV)), base + bias + resid.float()) < 6e-3
    x = aux.float().requires_grad_(True)
    O.gelu(x).sum().backward()
    assert rel(ops.gemm(ad, bd, block_n=-192, dgelu_aux=aux.to(DEV)), base * x.grad) < 6e-3
    assert rel(ops.gemm(ad, bd, block_n=192), base) < 6e-3  # same tile width, 1-CTA kernel
    with pytest.raises((MerlotError, ValueError)):
        ops.gemm(ad, b.t().contiguous().to(DEV), b_mn_major=True, block_n=-192)  # MN-major B is not tileable in 96-column halves


def test_gemm_shape_errors(ops):
    from merlot_b200._lib import MerlotShapeError
    a = torch.zeros(16, 12, dtype=torch.bfloat16, device=DEV)  # lda = 12 not a multiple of 8
    b = torch.zeros(16, 12, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(MerlotShapeError):
        ops.gemm(a, b)
    assert issubclass(MerlotShapeError, ValueError)  # the reference raises ValueError on shape mismatches


# ---------------------------------------------------------------------------------------------------------------
# K2/K3/K4 attention
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,S,heads,masked", [(1, 128, 1, False), (2, 64, 2, True), (2, 266, 12, False), (2, 396, 12, True),
                                              (3, 93, 4, True), (1, 885, 2, True), (1, 1, 1, False)])
def test_attention_fwd_bwd_colsum(ops, B, S, heads, masked):
    g = torch.Generator().manual_seed(S)
    H = heads * 64
    qkv = torch.randn(B * S, 3 * H, generator=g).bfloat16()
    valid = None
    mask = None
    if masked:
        lens = torch.randint(max(1, S // 3), S + 1, (B,), generator=g)
        v2 = (torch.arange(S)[None] < lens[:, None])
        if S > 10:
            v2[-1, 5:9] = False
        valid = v2.to(torch.uint8).reshape(-1).contiguous()
        vf = v2.float()
        mask = vf[:, None, :] * vf[:, :, None]
    x = qkv.float().reshape(B, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    q, k, v = (x[i].clone().requires_grad_(True) for i in range(3))
    probs, ctx4 = O.attention_core(q, k, v, mask)
    ctx_ref = ctx4.permute(0, 2, 1, 3).reshape(B * S, H)
    vd = valid.to(DEV) if valid is not None else None
    ctx, lse = ops.attention_fwd(qkv.to(DEV), B, S, heads, vd)
    assert rel(ctx, ctx_ref) < 1e-2
    d_ctx = (torch.randn(B * S, H, generator=g) * 0.1).bfloat16()
    # padding QUERY rows get a gradient too (they never do in the model): the reference's scores*m - 1e10*(1-m) keeps their
    # uniform probabilities in dV and sends nothing into q / k (utils/transformer.py:109-112)
    dqkv = ops.attention_bwd(qkv.to(DEV), ctx, d_ctx.to(DEV), lse, B, S, heads, vd)
    ctx_ref.backward(d_ctx.float())
    ref = torch.stack([q.grad, k.grad, v.grad], 0).permute(1, 3, 0, 2, 4).reshape(B * S, 3 * H)
    for i in range(3):
        assert rel(dqkv[:, i * H:(i + 1) * H], ref[:, i * H:(i + 1) * H]) < 1.5e-2
    colsum = torch.zeros(B, S, device=DEV)
    ops.attention_colsum(qkv.to(DEV), lse, colsum, B, S, heads, vd)
    assert rel(colsum, probs.detach().mean(1).sum(1)) < 2e-3  # head-mean, summed over queries (transformer.py:208-209)


@pytest.mark.parametrize("B,P,chunk,nch,heads", [(2, 13, 8, 4, 2), (2, 100, 32, 5, 4), (1, 266, 32, 4, 12), (2, 0, 16, 6, 1), (1, 70, 33, 3, 2)])
def test_attention_disable_pairwise_lang_attn(ops, B, P, chunk, nch, heads):
    """model/modeling.py:160-168: segment 0 = P vision tokens, segment 1 + c = language chunk c; a pair attends iff it shares a
    segment or either side is a vision token.  K2 / K3 / K4 and the export kernel take (P, chunk) and derive the partner set of
    every row arithmetically; the oracle gets the explicit [B, S, S] mask the reference builds.  Chunk boundaries fall inside
    32-position words, inside and across the 64 / 128-wide tiles, and some tokens are padding."""
    S = P + chunk * nch
    g = torch.Generator().manual_seed(S + chunk)
    H = heads * 64
    qkv = torch.randn(B * S, 3 * H, generator=g).bfloat16()
    v2 = torch.ones(B, S, dtype=torch.bool)
    for b in range(B):  # ragged captions: the tail of some chunks is padding (ids == 0)
        for c in range(nch):
            n_pad = int(torch.randint(0, chunk // 2 + 1, (1,), generator=g))
            if n_pad:
                v2[b, P + (c + 1) * chunk - n_pad:P + (c + 1) * chunk] = False
    seg = torch.cat([torch.zeros(P, dtype=torch.int64), 1 + torch.arange(chunk * nch) // chunk])
    can = (seg[:, None] == seg[None]) | (seg == 0)[None] | (seg == 0)[:, None]
    mask = (v2[:, None, :] & v2[:, :, None] & can[None]).float()
    valid = v2.to(torch.uint8).reshape(-1).contiguous().to(DEV)
    x = qkv.float().reshape(B, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    q, k, v = (x[i].clone().requires_grad_(True) for i in range(3))
    probs, ctx4 = O.attention_core(q, k, v, mask)
    ctx_ref = ctx4.permute(0, 2, 1, 3).reshape(B * S, H)
    pair = (P, chunk)
    ctx, lse = ops.attention_fwd(qkv.to(DEV), B, S, heads, valid, pair=pair)
    assert rel(ctx, ctx_ref) < 1e-2
    # the mask did something: without it the same inputs give a different context
    ctx_nopair, _ = ops.attention_fwd(qkv.to(DEV), B, S, heads, valid)
    assert nch < 2 or rel(ctx_nopair, ctx_ref) > 1e-2
    d_ctx = (torch.randn(B * S, H, generator=g) * 0.1).bfloat16()
    dqkv = ops.attention_bwd(qkv.to(DEV), ctx, d_ctx.to(DEV), lse, B, S, heads, valid, pair=pair)
    ctx_ref.backward(d_ctx.float())
    ref = torch.stack([q.grad, k.grad, v.grad], 0).permute(1, 3, 0, 2, 4).reshape(B * S, 3 * H)
    for i in range(3):
        assert rel(dqkv[:, i * H:(i + 1) * H], ref[:, i * H:(i + 1) * H]) < 1.5e-2
    colsum = torch.zeros(B, S, device=DEV)
    ops.attention_colsum(qkv.to(DEV), lse, colsum, B, S, heads, valid, pair=pair)
    assert rel(colsum, probs.detach().mean(1).sum(1)) < 2e-3
    pm = ops.attention_probs(qkv.to(DEV), lse, B, S, heads, valid, pair=pair)
    assert rel(pm, probs.detach().mean(1)) < 2e-3
    # a language token of chunk 0 puts (numerically) nothing on a language token of chunk 1
    if nch >= 2:
        assert float(pm[0, P, P + chunk].abs()) < 1e-6


# ---------------------------------------------------------------------------------------------------------------
# K5 LayerNorm, CE, l2norm
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,H", [(1, 768), (1000, 768), (37, 128), (8512, 768)])
def test_layernorm_fwd_bwd(ops, rows, H):
    g = torch.Generator().manual_seed(rows)
    x = (torch.randn(rows, H, generator=g) * 2 + 0.5).bfloat16()
    gam, bet = torch.randn(H, generator=g), torch.randn(H, generator=g)
    xr = x.float().requires_grad_(True)
    gr, br = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
    y_ref = O.layer_norm(xr, {"l/gamma": gr, "l/beta": br}, "l")
    y = torch.empty(rows, H, dtype=torch.bfloat16, device=DEV)
    mean, rstd = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
    ops.layernorm_fwd(x.to(DEV), y, gam.to(DEV), bet.to(DEV), mean, rstd)
    assert rel(y, y_ref) < 4e-3
    dy = torch.randn(rows, H, generator=g).bfloat16()
    dres = torch.randn(rows, H, generator=g).bfloat16()
    y_ref.backward(dy.float())
    dx = torch.empty(rows, H, dtype=torch.bfloat16, device=DEV)
    dg, db = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    ops.layernorm_bwd(dy.to(DEV), x.to(DEV), mean, rstd, gam.to(DEV), dx, dg, db, dres=dres.to(DEV))
    assert rel(dx, xr.grad + dres.float()) < 6e-3
    assert rel(dg, gr.grad) < 2e-3 and rel(db, br.grad) < 2e-3


@pytest.mark.parametrize("rows,H,with_dres,p", [(1, 768, True, 0.1), (13, 64, False, 0.0), (1777, 768, True, 0.1), (8512, 768, True, 0.1),
                                                  (8512, 768, False, 0.0), (5000, 1024, True, 0.1), (20000, 256, True, 0.0), (3, 512, True, 0.1)])
def test_layernorm_bwd_fused(ops, rows, H, with_dres, p):
    """The stacks' fused LayerNorm backward (rows arrive through per-warp bulk-copy rings): every output against the fp32 graph
    of utils/model_utils.py:113-130, the dropout mask against merlot_dropout_apply, the bias gradient against the exact column
    sums of what the kernel wrote; row counts that leave warps without rows, with one row, and with many ring refills."""
    g = torch.Generator().manual_seed(rows + H)
    x = (torch.randn(rows, H, generator=g) * 2 + 0.5).bfloat16()
    gam = torch.randn(H, generator=g)
    xr, gr, br = x.float().requires_grad_(True), gam.clone().requires_grad_(True), torch.zeros(H, requires_grad=True)
    y_ref = O.layer_norm(xr, {"l/gamma": gr, "l/beta": br}, "l")
    dy = torch.randn(rows, H, generator=g).bfloat16()
    dres = torch.randn(rows, H, generator=g).bfloat16() if with_dres else None
    y_ref.backward(dy.float())
    mu = x.float().mean(-1)
    rs = torch.rsqrt(x.float().var(-1, unbiased=False) + 1e-5)
    dx = torch.full((rows, H), float("nan"), dtype=torch.bfloat16, device=DEV)
    dmask = torch.full((rows, H), float("nan"), dtype=torch.bfloat16, device=DEV)
    dg, db, dbias = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    for rep in range(2):  # twice: the accumulators add up, the ring barriers start fresh every launch
        ops.layernorm_bwd_fused(dy.to(DEV), x.to(DEV), mu.to(DEV), rs.to(DEV), gam.to(DEV), dx, dg, db,
                                dres=None if dres is None else dres.to(DEV), dmask=dmask, dbias=dbias, dropout=(p, 7, 3))
    want = xr.grad + (dres.float() if with_dres else 0.0)
    assert torch.isfinite(dx.float()).all()
    assert rel(dx, want) < 6e-3
    assert rel(dg, 2 * gr.grad) < 2e-3 and rel(db, 2 * br.grad) < 2e-3
    if p > 0:
        ref_mask = torch.empty_like(dx)
        ops.dropout_apply(dx, ref_mask, p, 7, 3)
        assert torch.equal(dmask, ref_mask)
        assert rel(dbias, 2 * dmask.float().sum(0)) < 1e-5
        kept = float((dmask != 0).float().mean())
        assert abs(kept - (1 - p)) < (0.2 if rows * H < 4096 else 0.02)
    else:
        assert rel(dbias, 2 * dx.float().sum(0)) < 1e-5


def test_softmax_ce_and_l2norm(ops):
    g = torch.Generator().manual_seed(0)
    R, Cn, ld = 37, 50370, 50432
    logits = torch.zeros(R, ld)
    logits[:, :Cn] = torch.randn(R, Cn, generator=g) * 3
    labels = torch.randint(0, Cn, (R,), generator=g, dtype=torch.int32)
    lr = logits[:, :Cn].clone().requires_grad_(True)
    per_ref = O.raw_cross_entropy_with_logits(lr, labels)
    per, lse, corr = (torch.empty(R, device=DEV) for _ in range(3))
    ops.softmax_ce_fwd(logits.to(DEV), labels.to(DEV), Cn, per, lse, corr)
    assert torch.allclose(per.cpu(), per_ref.detach(), rtol=1e-5, atol=1e-5)
    assert torch.equal(corr.cpu(), (lr.argmax(-1) == labels).float())
    coeff = torch.rand(R, generator=g)
    (per_ref * coeff).sum().backward()
    dlog = torch.empty(R, ld, dtype=torch.float32, device=DEV)
    ops.softmax_ce_bwd(logits.to(DEV), labels.to(DEV), Cn, lse, coeff.to(DEV), dlog)
    assert rel(dlog[:, :Cn], lr.grad) < 1e-5 and float(dlog[:, Cn:].abs().max()) == 0.0
    x = torch.randn(32, 768, generator=g)
    xr = x.clone().requires_grad_(True)
    y_ref = O.l2_normalize(xr)
    y, inv = torch.empty(32, 768, device=DEV), torch.empty(32, device=DEV)
    ops.l2norm_fwd(x.to(DEV), y, inv)
    assert rel(y, y_ref) < 1e-6
    dy = torch.randn(32, 768, generator=g)
    y_ref.backward(dy)
    dx = torch.empty(32, 768, device=DEV)
    ops.l2norm_bwd(dy.to(DEV), y, inv, dx)
    assert rel(dx, xr.grad) < 1e-5


# ---------------------------------------------------------------------------------------------------------------
# K12 masking: bit-exact given injected draws (integer path)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,L,spanbert,use_attn", [(8, 128, True, True), (3, 32, True, True), (4, 128, False, True),
                                                   (2, 64, True, False), (2, 1024, True, True)])
def test_mask_inputs_bit_exact(ops, tiny_cfg, B, L, spanbert, use_attn):
    cfg = dict(tiny_cfg, masking_do_spanbert=spanbert, masking_use_attn=use_attn)
    g = torch.Generator().manual_seed(L + B)
    ids = torch.randint(100, 50370, (B, L), generator=g, dtype=torch.int32)
    ids[:, ::32] = O.START
    ids[:, -L // 8:] = 0
    summ = torch.rand(B, L, generator=g) * 12
    summ[:, 5] = summ[:, 9]  # force an exact tie: tf.math.top_k keeps the lower index
    k = int(L * 0.2)
    draws = O.make_mask_draws(B, L, k, 50370, seed=B)
    ref = O.mask_inputs(ids, summ if use_attn else None, cfg, draws)
    if use_attn:
        w = torch.tensor([1.0, 0.0]) * np.float32(ref["topk_val"] - 0.01) + np.float32(0.01)
        consts = (float(np.float32(ref["topk_val"] - 0.01)), float(np.float32(0.01)), float(torch.log(w)[0]), float(torch.log(w)[1]),
                  float(w.max()))
    else:
        consts = (0.0, 1.0, 0.0, 0.0, 1.0)
    m_ids = torch.empty(B, L, dtype=torch.int32, device=DEV)
    m_idx = torch.empty(B, k, dtype=torch.int32, device=DEV)
    ops.mask_inputs(ids.to(DEV), summ.to(DEV) if use_attn else None, {k_: v.to(DEV) for k_, v in draws.items()}, m_ids, m_idx, None,
                    int(L * 0.2), k, spanbert, O.MASK, consts)
    assert torch.equal(m_idx.cpu(), ref["masked_idx"])
    assert torch.equal(m_ids.cpu(), ref["masked_ids"])


# ---------------------------------------------------------------------------------------------------------------
# K10 AdamW: packed bf16 moments bit-exact over many steps, parameters to 1e-6
# ---------------------------------------------------------------------------------------------------------------
def test_adamw_vs_oracle(ops):
    g = torch.Generator().manual_seed(0)
    n = 100003
    cfg = dict(learning_rate=3e-4, num_train_steps=1000, num_warmup_steps=10, weight_decay_rate=0.1, beta_2=0.98, epsilon=1e-6,
               use_bfloat16_adam=True, param_overrides=[[["bias"], {"weight_decay_rate": 0}]])
    p_ref = {"w/kernel": torch.randn(n, generator=g) * 0.02, "w/bias": torch.randn(n, generator=g) * 0.02}
    opt = O.AdamOracle(p_ref, cfg)
    dev = {k: dict(p=v.clone().to(DEV), m=torch.zeros(n, dtype=torch.bfloat16, device=DEV),
                   v=torch.zeros(n, dtype=torch.bfloat16, device=DEV), pb=torch.zeros(n, dtype=torch.bfloat16, device=DEV))
           for k, v in p_ref.items()}
    for step in range(25):
        grads = {k: torch.randn(n, generator=g) * (10.0 ** float(torch.randint(-6, 1, (1,), generator=g))) for k in p_ref}
        s = opt.step_scalars()
        opt.apply_gradients(p_ref, grads)
        for k, d in dev.items():
            gd = grads[k].to(DEV)
            wd = 0.1 if "kernel" in k else 0.0
            ops.adamw_step(d["p"], gd, d["m"], d["v"], d["pb"], n, float(s["beta1"]), float(np.float32(1) - s["beta1"]),
                           float(s["beta2"]), float(np.float32(1) - s["beta2"]), float(s["eps"]), float(s["lr_t"]), wd, 1.0, True)
            assert float(gd.abs().max()) == 0.0  # zero_grad
    for k, d in dev.items():
        assert torch.equal(d["m"].cpu(), opt.m[k]), "first moment must be bit-exact"
        assert torch.equal(d["v"].cpu(), opt.v[k]), "packed second moment must be bit-exact"
        assert (d["p"].cpu() - p_ref[k]).abs().max().item() < 1e-6
        assert torch.equal(d["pb"].cpu(), d["p"].cpu().bfloat16())


def test_clip_by_global_norm(ops):
    """tf.clip_by_global_norm (utils/optimization.py:233-237): g * clip / max(norm, clip)."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1000003, generator=g) * 0.01
    x = torch.cat([x, torch.zeros(1)])  # 16-byte friendly length not required
    for clip in (0.5, 1e6):
        gd = x.clone().to(DEV)
        scratch = torch.zeros(1, dtype=torch.float64, device=DEV)
        norm = torch.zeros(1, device=DEV)
        ops.clip_by_global_norm(gd, clip, scratch, norm)
        ref_norm = x.double().norm().item()
        assert abs(float(norm) - ref_norm) < 1e-5 * ref_norm
        ref = x * (clip / max(ref_norm, clip))
        assert rel(gd, ref) < 1e-6


def test_device_mask_draws_distributions_and_bit_exact_masking(ops):
    """merlot_mask_draws: the five tf.random tensors of mask_inputs drawn on device -- right distributions, reproducible from
    the seed -- and K12 fed with them is bit-exact against the oracle fed the very same tensors."""
    from merlot_b200 import ops as o
    B, L, k, V = 64, 128, 25, 50370
    d = o.mask_draws(B, L, k, V, [0.625, 0.25, 0.125], seed=11, device=DEV)
    d2 = o.mask_draws(B, L, k, V, [0.625, 0.25, 0.125], seed=11, device=DEV)
    d3 = o.mask_draws(B, L, k, V, [0.625, 0.25, 0.125], seed=12, device=DEV)
    assert all(torch.equal(d[x], d2[x]) for x in d) and not torch.equal(d["gumbel"], d3["gumbel"])
    g = d["gumbel"].double().cpu()
    assert abs(float(g.mean()) - 0.5772) < 0.03 and abs(float(g.var()) - 1.6449) < 0.1          # Gumbel(0,1): mean gamma, var pi^2/6
    opt = torch.bincount(d["option"].cpu(), minlength=3).double() / (B * L)
    assert (opt - torch.tensor([0.1, 0.8, 0.1])).abs().max() < 0.02
    for name in ("span_lower", "span_upper"):
        f = torch.bincount(d[name].reshape(-1).cpu(), minlength=3).double() / (B * k)
        assert (f - torch.tensor([0.625, 0.25, 0.125])).abs().max() < 0.04
    r = d["rand_ids"].cpu()
    assert int(r.min()) >= 100 and int(r.max()) < V and abs(float(r.double().mean()) - (100 + V) / 2) < 300
    gen = torch.Generator().manual_seed(0)
    ids = torch.randint(100, V, (B, L), generator=gen, dtype=torch.int32)
    ids[:, 0] = 2
    ids[:, 100:] = 0
    summ = torch.rand(B, L, generator=gen)
    cfg = dict(masking_use_topk_from_attn_perc=0.2, masking_choose_topk_prob=0.5, masking_rate=0.2, masking_do_spanbert=True, masking_use_attn=True)
    ref = O.mask_inputs(ids, summ, cfg, {x: v.cpu() for x, v in d.items()})
    import numpy as np
    nontop, top = 0.01, 0.01 * 0.5 * 0.8 / (0.2 * 0.5)
    w = torch.tensor([1.0, 0.0]) * np.float32(top - nontop) + np.float32(nontop)
    consts = (float(np.float32(top - nontop)), float(np.float32(nontop)), float(torch.log(w)[0]), float(torch.log(w)[1]), float(w.max()))
    mi = torch.empty(B, L, dtype=torch.int32, device=DEV)
    mx = torch.empty(B, k, dtype=torch.int32, device=DEV)
    o.mask_inputs(ids.to(DEV), summ.to(DEV), d, mi, mx, None, 25, k, True, 1, consts)
    assert torch.equal(mi.cpu(), ref["masked_ids"]) and torch.equal(mx.cpu(), ref["masked_idx"])

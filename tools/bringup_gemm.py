"""GPU bring-up for K1 (run under gpurun): each operand-major combination in its own process so a trap in one
does not poison the CUDA context of the others.  Prints max/rel errors against a torch fp32 matmul of the same bf16
inputs; exits non-zero on mismatch.  Usage: python tools/bringup_gemm.py [kk|kmn|mnmn|mnk|epi|perf]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from merlot_b200 import ops  # noqa: E402


def ref_mm(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    B = b.float() if b_mn else b.float().t()
    return A @ B


def check(name, got, want, tol=2e-2):
    got = got.float()
    err = (got - want).abs().max().item()
    rel = ((got - want).norm() / (want.norm() + 1e-30)).item()
    ok = rel < tol and torch.isfinite(got).all().item()
    print(f"{'OK  ' if ok else 'FAIL'} {name}: max_abs={err:.4e} rel_fro={rel:.4e}", flush=True)
    if not ok:
        bad = ((got - want).abs() > 0.05 * want.abs().max()).nonzero()
        print("   first bad idx:", bad[:8].tolist(), " n_bad:", bad.shape[0], flush=True)
        rows = torch.unique(bad[:, 0])[:16].tolist()
        cols = torch.unique(bad[:, 1])[:16].tolist()
        print("   bad rows (first 16):", rows, " bad cols (first 16):", cols, flush=True)
    return ok


def run_major(a_mn, b_mn):
    dev = "cuda"
    ok = True
    g = torch.Generator(device="cpu").manual_seed(0)
    for (M, N, K, bn) in [(128, 256, 64, 256), (128, 128, 64, 128), (128, 256, 256, 256), (256, 512, 768, 256),
                          (200, 264, 136, 0), (8512, 768, 768, 0), (1024, 2304, 768, 0), (300, 50376, 768, 0),
                          (3168, 3072, 768, 128), (1000, 768, 3072, 0), (520, 768, 768, 192), (8512, 2304, 768, 192)]:
        a = (torch.randn((K, M) if a_mn else (M, K), generator=g) * 0.5).bfloat16().to(dev)
        b = (torch.randn((K, N) if b_mn else (N, K), generator=g) * 0.5).bfloat16().to(dev)
        if (M % 8 and a_mn) or (N % 8 and b_mn):
            continue
        out = ops.gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn, out_dtype=torch.float32, block_n=bn)
        torch.cuda.synchronize()
        ok &= check(f"a_mn={a_mn} b_mn={b_mn} M={M} N={N} K={K} bn={bn}", out, ref_mm(a, b, a_mn, b_mn), 1e-3)
    return ok


def run_epi():
    dev = "cuda"
    ok = True
    g = torch.Generator(device="cpu").manual_seed(1)
    M, N, K = 520, 768, 768
    a = (torch.randn(M, K, generator=g) * 0.3).bfloat16().to(dev)
    w = (torch.randn(K, N, generator=g) * 0.05).bfloat16().to(dev)  # TF layout [in,out] -> b_mn_major
    bias = torch.randn(N, generator=g).to(dev)
    resid = torch.randn(M, N, generator=g).bfloat16().to(dev)
    base = a.float() @ w.float() + bias
    out = ops.gemm(a, w, b_mn_major=True, bias=bias)
    ok &= check("bias bf16", out, base, 6e-3)
    out = ops.gemm(a, w, b_mn_major=True, bias=bias, resid=resid)
    ok &= check("bias+resid", out, base + resid.float(), 6e-3)
    pre = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    act = ops.gemm(a, w, b_mn_major=True, bias=bias, gelu=True, out_pre=pre)
    ok &= check("gelu pre", pre, base, 6e-3)
    ok &= check("gelu act", act, torch.nn.functional.gelu(base), 6e-3)
    aux = torch.randn(M, N, generator=g).bfloat16().to(dev)
    x = aux.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    out = ops.gemm(a, w, b_mn_major=True, dgelu_aux=aux)
    ok &= check("mul dgelu", out, (a.float() @ w.float()) * x.grad, 6e-3)
    # split-K atomic wgrad: dW[K,N] = a^T @ dy
    dy = (torch.randn(M, N, generator=g) * 0.1).bfloat16().to(dev)
    dw = torch.zeros(K, N, dtype=torch.float32, device=dev)
    ops.gemm(a, dy, a_mn_major=True, b_mn_major=True, out=dw, atomic=True, M=K, N=N, K=M)
    ops.gemm(a, dy, a_mn_major=True, b_mn_major=True, out=dw, atomic=True, M=K, N=N, K=M)
    ok &= check("wgrad atomic x2", dw, 2 * (a.float().t() @ dy.float()), 1e-3)
    # dropout: rate + determinism + scaling
    o1 = ops.gemm(a, w, b_mn_major=True, bias=bias, dropout_p=0.1, dropout_seed=7, dropout_site=3)
    o2 = ops.gemm(a, w, b_mn_major=True, bias=bias, dropout_p=0.1, dropout_seed=7, dropout_site=3)
    o3 = ops.gemm(a, w, b_mn_major=True, bias=bias, dropout_p=0.1, dropout_seed=8, dropout_site=3)
    rate = (o1 == 0).float().mean().item()
    same = torch.equal(o1, o2)
    diff = (o1 != o3).float().mean().item()
    kept = o1 != 0
    ok_scale = check("dropout kept scale", o1[kept], (base / 0.9)[kept], 6e-3)
    print(f"dropout zero-rate={rate:.4f} deterministic={same} differs_with_seed={diff:.3f}")
    ok &= ok_scale and same and abs(rate - 0.1) < 0.01 and diff > 0.1
    return ok


def run_perf():
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(2)
    for (M, N, K, a_mn, b_mn, f32) in [(8512, 2304, 768, 0, 1, 0), (8512, 768, 768, 0, 1, 0), (8512, 3072, 768, 0, 1, 0),
                                       (8512, 768, 3072, 0, 1, 0), (768, 3072, 8512, 1, 1, 1), (3072, 768, 8512, 1, 1, 1),
                                       (3168, 768, 768, 0, 1, 0), (3168, 2304, 768, 0, 1, 0), (768, 768, 8512, 1, 1, 1), (8192, 8192, 8192, 0, 0, 0)]:
        a = (torch.randn((K, M) if a_mn else (M, K), generator=g) * 0.1).bfloat16().to(dev)
        b = (torch.randn((K, N) if b_mn else (N, K), generator=g) * 0.1).bfloat16().to(dev)
        for bn in (128, 192, 256):
            out = torch.zeros(M, N, dtype=torch.float32 if f32 else torch.bfloat16, device=dev)
            kw = dict(a_mn_major=bool(a_mn), b_mn_major=bool(b_mn), out=out, atomic=bool(f32), block_n=bn)
            for _ in range(3):
                ops.gemm(a, b, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ops.gemm(a, b, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"perf M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} bn={bn}: {ms * 1e3:.1f} us  "
                  f"{2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    return True


def run_pair():
    """CTA-pair (cta_group::2) kernel: block_n = -256."""
    dev = "cuda"
    ok = True
    g = torch.Generator(device="cpu").manual_seed(3)
    for (a_mn, b_mn) in [(False, False), (False, True), (True, True), (True, False)]:
        for (M, N, K) in [(256, 256, 64), (256, 256, 256), (512, 512, 768), (520, 768, 136), (8512, 2304, 768), (1000, 264, 3072)]:
            if (M % 8 and a_mn) or (N % 8 and b_mn):
                continue
            a = (torch.randn((K, M) if a_mn else (M, K), generator=g) * 0.5).bfloat16().to(dev)
            b = (torch.randn((K, N) if b_mn else (N, K), generator=g) * 0.5).bfloat16().to(dev)
            for f32 in (True, False):
                out = ops.gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn, out_dtype=torch.float32 if f32 else torch.bfloat16, block_n=-256)
                torch.cuda.synchronize()
                ok &= check(f"pair a_mn={a_mn} b_mn={b_mn} M={M} N={N} K={K} f32={f32}", out, ref_mm(a, b, a_mn, b_mn), 1e-3 if f32 else 6e-3)
    # epilogues through the pair kernel
    M, N, K = 520, 768, 768
    a = (torch.randn(M, K, generator=g) * 0.3).bfloat16().to(dev)
    w = (torch.randn(K, N, generator=g) * 0.05).bfloat16().to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    resid = torch.randn(M, N, generator=g).bfloat16().to(dev)
    base = a.float() @ w.float() + bias
    ok &= check("pair bias+resid", ops.gemm(a, w, b_mn_major=True, bias=bias, resid=resid, block_n=-256), base + resid.float(), 6e-3)
    pre = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    act = ops.gemm(a, w, b_mn_major=True, bias=bias, gelu=True, out_pre=pre, block_n=-256)
    ok &= check("pair gelu pre", pre, base, 6e-3)
    ok &= check("pair gelu act", act, torch.nn.functional.gelu(base), 6e-3)
    dy = (torch.randn(M, N, generator=g) * 0.1).bfloat16().to(dev)
    dw = torch.zeros(K, N, dtype=torch.float32, device=dev)
    for _ in range(2):
        ops.gemm(a, dy, a_mn_major=True, b_mn_major=True, out=dw, atomic=True, M=K, N=N, K=M, block_n=-256)
    ok &= check("pair wgrad reduce-add x2", dw, 2 * (a.float().t() @ dy.float()), 1e-3)
    return ok


def run_pair_perf():
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(2)
    for (M, N, K, a_mn, b_mn, f32) in [(8512, 2304, 768, 0, 1, 0), (8512, 768, 768, 0, 1, 0), (8512, 3072, 768, 0, 1, 0),
                                       (8512, 768, 3072, 0, 0, 0), (768, 3072, 8512, 1, 1, 1), (3168, 2304, 768, 0, 1, 0),
                                       (8192, 8192, 8192, 0, 0, 0)]:
        a = (torch.randn((K, M) if a_mn else (M, K), generator=g) * 0.1).bfloat16().to(dev)
        b = (torch.randn((K, N) if b_mn else (N, K), generator=g) * 0.1).bfloat16().to(dev)
        for bn in (0, -256):
            out = torch.zeros(M, N, dtype=torch.float32 if f32 else torch.bfloat16, device=dev)
            kw = dict(a_mn_major=bool(a_mn), b_mn_major=bool(b_mn), out=out, atomic=bool(f32), block_n=bn)
            for _ in range(3):
                ops.gemm(a, b, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ops.gemm(a, b, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"perf M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} {'PAIR' if bn else 'auto'}: {ms * 1e3:.1f} us  "
                  f"{2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    return True


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "kk"
    t0 = time.time()
    ok = {"kk": lambda: run_major(False, False), "kmn": lambda: run_major(False, True),
          "mnmn": lambda: run_major(True, True), "mnk": lambda: run_major(True, False),
          "epi": run_epi, "perf": run_perf, "pair": run_pair, "pairperf": run_pair_perf}[mode]()
    print(f"[{mode}] {'PASS' if ok else 'FAIL'} in {time.time() - t0:.1f}s", flush=True)
    sys.exit(0 if ok else 1)

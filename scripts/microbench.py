"""Per-kernel micro-benchmarks at the BASELINE config-2 shapes (CUDA events, L2-flushing between
iterations is unnecessary: every operand set is rotated through > 126 MB of distinct buffers)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ergm_b200 import ops, _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--dropout", type=float, default=0.1)
args = ap.parse_args()
dev = "cuda"
B, T, H, nh, I, V = 32, 256, 768, 12, 3072, 50260
M = B * T


def timeit(name, fn, nbuf, flops=None, bytes_=None, iters=args.iters):
    """GPU time per launch: the launches are captured into a CUDA graph so that host-side work
    (tensor-map encoding, ctypes) cannot leave the GPU idle between them."""
    for i in range(3):
        fn(i % nbuf)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i % nbuf)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    msg = "%-44s %9.1f us" % (name, us)
    if flops:
        msg += "  %8.1f TFLOP/s" % (flops / us / 1e6)
    if bytes_:
        msg += "  %8.1f GB/s" % (bytes_ / us / 1e3)
    print(msg, flush=True)


def want(k):
    return not args.only or any(s in k for s in args.only.split(","))


def gemm_case(name, Mm, N, K, a_mn, b_mn, out_dtype=torch.bfloat16, nbuf=4, **kw):
    if not want("gemm"):
        return
    ldk = (K + 7) // 8 * 8
    ldm = (Mm + 7) // 8 * 8
    As = [(torch.randn(K, ldm, device=dev).bfloat16()[:, :Mm] if a_mn else torch.randn(Mm, ldk, device=dev).bfloat16()[:, :K])
          for _ in range(nbuf)]
    ldb = (N + 7) // 8 * 8
    Bs = [(torch.randn(K, ldb, device=dev).bfloat16()[:, :N] if b_mn else torch.randn(N, ldk, device=dev).bfloat16()[:, :K])
          for _ in range(nbuf)]
    ldd = (N + 63) // 64 * 64
    Ds = [torch.zeros(Mm, ldd, device=dev, dtype=out_dtype) for _ in range(nbuf)]
    extra = {}
    if kw.pop("bias", False):
        extra["bias"] = torch.randn(N, device=dev)
    if kw.pop("residual", False):
        extra["residual"] = "self"
    def fn(i):
        e = dict(extra)
        if e.get("residual") == "self":
            e["residual"] = Ds[i]
        ops.gemm(As[i], Bs[i], Ds[i], M=Mm, N=N, K=K, a_major=int(a_mn), b_major=int(b_mn), **e, **kw)
    timeit("gemm %s [%d,%d,%d]" % (name, Mm, N, K), fn, nbuf, flops=2.0 * Mm * N * K)


gemm_case("qkv fwd", M, 3 * H, H, 0, 1, bias=True)
gemm_case("qkv fwd bn128", M, 3 * H, H, 0, 1, bias=True, block_n=128)
gemm_case("proj fwd+res", M, H, H, 0, 1, out_dtype=torch.float32, bias=True, residual=True)
gemm_case("proj fwd+res bn128", M, H, H, 0, 1, out_dtype=torch.float32, bias=True, residual=True, block_n=128)
gemm_case("proj fwd+res bn64", M, H, H, 0, 1, out_dtype=torch.float32, bias=True, residual=True, block_n=64)
gemm_case("fc fwd gelu", M, I, H, 0, 1, bias=True, epilogue=L.EPI_GELU)
gemm_case("mlp proj fwd+res", M, H, I, 0, 1, out_dtype=torch.float32, bias=True, residual=True)
gemm_case("lm head fwd", M, V, H, 0, 0, nbuf=2)
gemm_case("lm head dgrad", M, H, V, 0, 1, out_dtype=torch.float32, nbuf=2)
gemm_case("lm head wgrad", V, H, M, 1, 1, out_dtype=torch.float32, nbuf=2, epilogue=L.EPI_ATOMIC, block_n=128)
gemm_case("fc dgrad", M, H, I, 0, 0)
gemm_case("proj dgrad (4H out)", M, I, H, 0, 0)
gemm_case("qkv wgrad", H, 3 * H, M, 1, 1, out_dtype=torch.float32, epilogue=L.EPI_ATOMIC, block_n=128)
gemm_case("proj wgrad sk4", H, H, M, 1, 1, out_dtype=torch.float32, epilogue=L.EPI_ATOMIC, block_n=128, split_k=4)
gemm_case("fc wgrad", H, I, M, 1, 1, out_dtype=torch.float32, epilogue=L.EPI_ATOMIC, block_n=128)
gemm_case("decode qkv M=64", 64, 3 * H, H, 0, 1, bias=True)
gemm_case("decode lm head M=64", 64, V, H, 0, 0, out_dtype=torch.float32, nbuf=2)

if want("attn"):
    nb = 6
    qkvs = [torch.randn(M, 3 * H, device=dev).bfloat16() for _ in range(nb)]
    outs = [torch.zeros(M, H, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
    o32 = [torch.zeros(M, H, device=dev) for _ in range(nb)]
    lse = torch.zeros(B, nh, T, device=dev)
    fl_c = 4.0 * B * nh * T * T * 64 / 2
    for pd in (0.0, args.dropout):
        timeit("attn_fwd causal p=%.1f" % pd, lambda i: ops.attn_fwd(qkvs[i], qkvs[i], qkvs[i], outs[i], lse, B=B, nh=nh, Tq=T, Tk=T,
               k_col0=H, v_col0=2 * H, causal=True, dropout_p=pd, seed=1, offset=2, out_f32=o32[i]), nb, flops=fl_c)
        timeit("attn_fwd cross  p=%.1f" % pd, lambda i: ops.attn_fwd(qkvs[i], qkvs[i], qkvs[i], outs[i], lse, B=B, nh=nh, Tq=T, Tk=T,
               k_col0=H, v_col0=2 * H, causal=False, dropout_p=pd, seed=1, offset=2, out_f32=o32[i]), nb, flops=2 * fl_c)
    douts = [torch.randn(M, H, device=dev).bfloat16() for _ in range(nb)]
    dqkv = [torch.zeros(M, 3 * H, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
    dq = torch.zeros(M, H, device=dev, dtype=torch.bfloat16)
    delta = torch.zeros(B, nh, T, device=dev)
    for pd in (0.0, args.dropout):
        for causal in (True, False):
            timeit("attn_bwd %s p=%.1f" % ("causal" if causal else "cross ", pd),
                   lambda i: ops.attn_bwd(qkvs[i], qkvs[i], qkvs[i], outs[i], douts[i], lse, delta, dq, dqkv[i], dqkv[i], B=B,
                                          nh=nh, Tq=T, Tk=T, k_col0=H, v_col0=2 * H, dk_col0=H, dv_col0=2 * H, causal=causal,
                                          dropout_p=pd, seed=1, offset=2, out_f32=o32[i]), nb,
                   flops=(2.5 * fl_c if causal else 5 * fl_c))

if want("ln"):
    nb = 6
    xs = [torch.randn(M, H, device=dev) for _ in range(nb)]
    ys = [torch.zeros(M, H, device=dev, dtype=torch.bfloat16) for _ in range(nb)]
    g, bt = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    mean, rstd = torch.zeros(M, device=dev), torch.ones(M, device=dev)
    timeit("ln_fwd", lambda i: ops.ln_fwd(xs[i], g, bt, ys[i], None, mean, rstd, 1e-5), nb, bytes_=M * H * 6.0)
    dxs = [torch.zeros(M, H, device=dev) for _ in range(nb)]
    dg, db, dn = torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev)
    for pd in (0.0, args.dropout):
        timeit("ln_bwd p=%.1f" % pd, lambda i: ops.ln_bwd(ys[i], xs[i], mean, rstd, g, dxs[i], dxs[i], ys[(i + 1) % nb], dg, db, dn,
               dropout_p=pd, seed=1, offset=3), nb, bytes_=M * H * (2 + 4 + 4 + 4 + 2.0))
    big = torch.randn(M, I, device=dev).bfloat16()
    out = torch.zeros(I, device=dev)
    timeit("colsum [M,4H]", lambda i: ops.colsum_bf16(big, out), 1, bytes_=M * I * 2.0)

"""Where do the K=768, N=768 GEMMs lose their time?  Tile shapes x epilogue variants (NOST = no epilogue
global traffic at all, NOGL = transposition only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = sys.argv[:1]
import torch
from ergm_b200 import ops, _lib as L
dev = "cuda"
def run(M, N, K, a_mn, b_mn, bn, out_dtype=torch.bfloat16, iters=40, nbuf=6, tag="", **kw):
    As = [torch.randn(K, M, device=dev).bfloat16() if a_mn else torch.randn(M, K, device=dev).bfloat16() for _ in range(nbuf)]
    Bs = [torch.randn(K, N, device=dev).bfloat16() if b_mn else torch.randn(N, K, device=dev).bfloat16() for _ in range(nbuf)]
    Ds = [torch.zeros(M, (N + 63) // 64 * 64, device=dev, dtype=out_dtype) for _ in range(nbuf)]
    if kw.pop("res", False):
        R = [torch.randn(M, N, device=dev) for _ in range(nbuf)]
        f = lambda i: ops.gemm(As[i], Bs[i], Ds[i], M=M, N=N, K=K, a_major=a_mn, b_major=b_mn, block_n=bn, residual=R[i], **kw)
    else:
        f = lambda i: ops.gemm(As[i], Bs[i], Ds[i], M=M, N=N, K=K, a_major=a_mn, b_major=b_mn, block_n=bn, **kw)
    for i in range(3): f(i % nbuf)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters): f(i % nbuf)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters / 2
    print("M=%5d N=%5d K=%5d a%d b%d bn=%4d %-8s %7.1f us %7.1f TF  %s" % (M, N, K, a_mn, b_mn, bn, str(out_dtype)[6:], us, 2.0 * M * N * K / us / 1e6, tag), flush=True)
NOST, NOGL = 1 << 30, 1 << 29
bias = torch.randn(3072, device=dev)
for bn in (64, 128, 256, 2128, 2256):
    run(8192, 768, 768, 0, 0, bn, tag="dgrad bf16 out")
    run(8192, 768, 768, 0, 0, bn, tag="dgrad NOST", epilogue=NOST)
    run(8192, 768, 768, 0, 1, bn, out_dtype=torch.float32, res=True, bias=bias, tag="fwd proj f32+res+bias")
    run(8192, 768, 768, 0, 1, bn, out_dtype=torch.float32, bias=bias, tag="fwd proj f32 no res")
for bn in (128, 256, 2128, 2256):
    run(8192, 2304, 768, 0, 1, bn, bias=bias, tag="qkv")
    run(8192, 768, 3072, 0, 1, bn, out_dtype=torch.float32, res=True, bias=bias, tag="proj2 f32+res")
    run(8192, 768, 3072, 0, 0, bn, tag="dgrad fc K=3072")
for sk in (1, 2, 4, 8):
    run(768, 768, 8192, 1, 1, 128, out_dtype=torch.float32, epilogue=L.EPI_ATOMIC, split_k=sk, tag="wgrad sk%d" % sk)
    run(768, 768, 8192, 1, 1, 2128, out_dtype=torch.float32, epilogue=L.EPI_ATOMIC, split_k=sk, tag="wgrad pair sk%d" % sk)

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ergm_b200 import ops, _lib as L
dev = "cuda"
def run(M, N, K, a_mn, b_mn, bn, out_dtype=torch.bfloat16, iters=20, nbuf=4, **kw):
    ldk = (K + 7) // 8 * 8
    As = [torch.randn(K, M, device=dev).bfloat16() if a_mn else torch.randn(M, ldk, device=dev).bfloat16()[:, :K] for _ in range(nbuf)]
    Bs = [torch.randn(K, N, device=dev).bfloat16() if b_mn else torch.randn(N, ldk, device=dev).bfloat16()[:, :K] for _ in range(nbuf)]
    Ds = [torch.zeros(M, (N + 63) // 64 * 64, device=dev, dtype=out_dtype) for _ in range(nbuf)]
    f = lambda i: ops.gemm(As[i], Bs[i], Ds[i], M=M, N=N, K=K, a_major=a_mn, b_major=b_mn, block_n=bn, **kw)
    for i in range(3): f(i % nbuf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): f(i % nbuf)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    print("M=%5d N=%5d K=%5d a_mn=%d b_mn=%d bn=%4d %-8s %8.1f us %7.1f TF %s" % (M, N, K, a_mn, b_mn, bn, str(out_dtype)[6:], us, 2.0 * M * N * K / us / 1e6, sorted(kw)), flush=True)
NOST = 1 << 30
NOGL = 1 << 29
u = torch.randn(8192, 3072, device=dev).bfloat16()
for bn in (256, 2256, 128):
    run(8192, 3072, 768, 0, 0, bn)
    run(8192, 3072, 768, 0, 0, bn, gelu_grad_of=u)
    run(8192, 3072, 768, 0, 0, bn, epilogue=L.EPI_GELU)
res = torch.randn(8192, 768, device=dev)
bias = torch.randn(768, device=dev)
for bn in (128, 2128):
    run(8192, 768, 768, 0, 1, bn, out_dtype=torch.float32)
    run(8192, 768, 768, 0, 1, bn, out_dtype=torch.float32, residual=res)
    run(8192, 768, 768, 0, 1, bn, out_dtype=torch.float32, residual=res, bias=bias, dropout_p=0.1, seed=1, offset=2)

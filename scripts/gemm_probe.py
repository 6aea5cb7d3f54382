import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ergm_b200 import ops, _lib as L
dev = "cuda"
def run(M, N, K, a_mn, b_mn, bn, out_dtype=torch.bfloat16, iters=10, **kw):
    ldk = (K + 7) // 8 * 8
    a = torch.randn(K, M, device=dev).bfloat16() if a_mn else torch.randn(M, ldk, device=dev).bfloat16()[:, :K]
    b = torch.randn(K, N, device=dev).bfloat16() if b_mn else torch.randn(N, ldk, device=dev).bfloat16()[:, :K]
    d = torch.zeros(M, (N + 63) // 64 * 64, device=dev, dtype=out_dtype)
    f = lambda: ops.gemm(a, b, d, M=M, N=N, K=K, a_major=a_mn, b_major=b_mn, block_n=bn, **kw)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    print("M=%5d N=%5d K=%5d a_mn=%d b_mn=%d bn=%4d %-8s %8.1f us %7.1f TF %s" % (M, N, K, a_mn, b_mn, bn, str(out_dtype)[6:], us, 2.0 * M * N * K / us / 1e6, kw if kw else ""), flush=True)

for bn in (128, 256, 2128, 2256):
    run(8192, 8192, 8192, 0, 0, bn)
for bn in (256, 2256):
    run(8192, 8192, 8192, 0, 1, bn)
    run(8192, 8192, 8192, 1, 1, bn, out_dtype=torch.float32)
for bn in (128, 256, 2128, 2256):
    run(8192, 2304, 768, 0, 1, bn)
for bn in (128, 256, 2128, 2256):
    run(8192, 768, 768, 0, 1, bn, out_dtype=torch.float32)
for bn in (128, 256, 2128, 2256):
    run(8192, 3072, 768, 0, 1, bn)
    run(8192, 768, 3072, 0, 1, bn, out_dtype=torch.float32)
for bn in (256, 2256):
    run(8192, 50260, 768, 0, 0, bn, iters=4)
for bn, sk in ((128, 1), (2128, 1), (2128, 2), (2256, 2), (2256, 4)):
    run(768, 2304, 8192, 1, 1, bn, out_dtype=torch.float32, epilogue=L.EPI_ATOMIC, split_k=sk)
for bn, sk in ((128, 4), (2128, 4), (2128, 8), (2256, 8)):
    run(768, 768, 8192, 1, 1, bn, out_dtype=torch.float32, epilogue=L.EPI_ATOMIC, split_k=sk)

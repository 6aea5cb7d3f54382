"""ncu driver: the out-projection GEMM shape 8192x768x768 (fp32 out + bias + residual), chosen tile via argv."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ergm_b200 import ops
dev = "cuda"
bn = int(sys.argv[1]) if len(sys.argv) > 1 else 256
res_on = "--nores" not in sys.argv
M, N, K = 8192, 768, 768
a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(K, N, device=dev).bfloat16()
d = torch.zeros(M, N, device=dev); res = torch.randn(M, N, device=dev); bias = torch.randn(N, device=dev)
for _ in range(3):
    ops.gemm(a, w, d, M=M, N=N, K=K, a_major=0, b_major=1, block_n=bn, bias=bias, residual=res if res_on else None)
torch.cuda.synchronize()
print("done")

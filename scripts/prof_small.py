"""Tiny driver for ncu: a few launches of each hot kernel at config-2 shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ergm_b200 import ops, _lib as L
dev = "cuda"
B, T, H, nh, I = 32, 256, 768, 12, 3072
M = B * T
which = sys.argv[1] if len(sys.argv) > 1 else "all"
qkv = torch.randn(M, 3 * H, device=dev).bfloat16()
out = torch.zeros(M, H, device=dev, dtype=torch.bfloat16)
o32 = torch.zeros(M, H, device=dev)
lse = torch.zeros(B, nh, T, device=dev)
if which in ("all", "attn"):
    for causal in (True, False):
        for _ in range(2):
            ops.attn_fwd(qkv, qkv, qkv, out, lse, B=B, nh=nh, Tq=T, Tk=T, k_col0=H, v_col0=2 * H, causal=causal, out_f32=o32, dropout_p=0.1, seed=1, offset=2)
    dout = torch.randn(M, H, device=dev).bfloat16()
    dqkv = torch.zeros(M, 3 * H, device=dev, dtype=torch.bfloat16)
    dq = torch.zeros(M, H, device=dev, dtype=torch.bfloat16)
    delta = torch.zeros(B, nh, T, device=dev)
    for causal in (True, False):
        for _ in range(2):
            ops.attn_bwd(qkv, qkv, qkv, out, dout, lse, delta, dq, dqkv, dqkv, B=B, nh=nh, Tq=T, Tk=T, k_col0=H, v_col0=2 * H,
                         dk_col0=H, dv_col0=2 * H, causal=causal, out_f32=o32, dropout_p=0.1, seed=1, offset=2)
if which in ("all", "gemm"):
    a = torch.randn(M, H, device=dev).bfloat16()
    w = torch.randn(H, 3 * H, device=dev).bfloat16()
    bias = torch.randn(3 * H, device=dev)
    d = torch.zeros(M, 3 * H, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(a, w, d, M=M, N=3 * H, K=H, bias=bias)
    w2 = torch.randn(H, H, device=dev).bfloat16()
    d2 = torch.zeros(M, H, device=dev)
    for _ in range(3):
        ops.gemm(a, w2, d2, M=M, N=H, K=H, residual=d2)
if which in ("all", "ln"):
    x = torch.randn(M, H, device=dev)
    g, bt = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    mean, rstd = torch.zeros(M, device=dev), torch.ones(M, device=dev)
    dx = torch.zeros(M, H, device=dev)
    dxb = torch.zeros(M, H, device=dev, dtype=torch.bfloat16)
    dg, db, dn = torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev)
    for _ in range(3):
        ops.ln_bwd(out, x, mean, rstd, g, dx, dx, dxb, dg, db, dn)
torch.cuda.synchronize()
print("done")

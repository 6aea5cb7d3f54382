import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gemm_probe3.py")).read().split("NOST, NOGL = 1 << 30, 1 << 29")[0]
exec(src)
bias = torch.randn(3072, device=dev)
u = torch.empty(8192, 3072, device=dev, dtype=torch.bfloat16)
for bn in (192, 256, 2256):
    run(8192, 768, 3072, 0, 1, bn, out_dtype=torch.float32, res=True, bias=bias, dropout_p=0.1, seed=1, offset=2, tag="proj2 fwd f32+res+bias+drop")
    run(8192, 768, 3072, 0, 0, bn, tag="dgrad fc (K=3072) bf16")
    run(8192, 768, 2304, 0, 0, bn, tag="dgrad qkv (K=2304) bf16")
    run(8192, 768, 1536, 0, 0, bn, tag="dgrad kv2 (K=1536) bf16")
    run(8192, 3072, 768, 0, 1, bn, bias=bias, preact=u, epilogue=L.EPI_GELU, tag="fc fwd gelu+preact")
    run(8192, 3072, 768, 0, 0, bn, tag="dgrad proj2 (N=3072) bf16")
    run(8192, 2304, 768, 0, 1, bn, bias=bias, tag="qkv fwd")
    run(8192, 1536, 768, 0, 1, bn, bias=bias, tag="kv2 fwd")

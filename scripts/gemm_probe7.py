import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gemm_probe3.py")).read().split("NOST, NOGL = 1 << 30, 1 << 29")[0]
exec(src)
bias = torch.randn(3072, device=dev)
for bn in (128, 192, 256):
    run(8192, 768, 768, 0, 0, bn, tag="dgrad bf16")
    run(8192, 768, 768, 0, 1, bn, out_dtype=torch.float32, res=True, bias=bias, tag="fwd f32+res+bias")
    run(8192, 768, 768, 0, 1, bn, out_dtype=torch.float32, res=True, bias=bias, dropout_p=0.1, seed=1, offset=2, tag="fwd f32+res+bias+drop")
    run(8192, 768, 768, 0, 1, bn, bias=bias, tag="fwd bf16+bias (q2)")
    run(64, 50260, 768, 0, 0, bn, out_dtype=torch.float32, iters=10, nbuf=2, tag="decode LM head")
    run(8192, 1536, 768, 0, 1, bn, bias=bias, tag="kv2 fwd")

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib.util
spec = importlib.util.spec_from_file_location("p3", os.path.join(os.path.dirname(os.path.abspath(__file__)), "gemm_probe3.py"))
src = open(spec.origin).read().split("NOST, NOGL = 1 << 30, 1 << 29")[0]
exec(src)
NOST, NOGL = 1 << 30, 1 << 29
bias = torch.randn(3072, device=dev)
for bn in (128, 256):
    for (M, N, K) in ((8192, 768, 768), (8192, 2304, 768)):
        run(M, N, K, 0, 1, bn, tag="NOST", epilogue=NOST)
        run(M, N, K, 0, 1, bn, tag="NOGL (tmem ld + transpose only)", epilogue=NOGL)
        run(M, N, K, 0, 1, bn, tag="bf16 out")
        run(M, N, K, 0, 1, bn, out_dtype=torch.float32, tag="f32 out")
        run(M, N, K, 0, 1, bn, out_dtype=torch.float32, bias=bias, tag="f32 out + bias")
        run(M, N, K, 0, 1, bn, out_dtype=torch.float32, res=True, tag="f32 out + res")

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gemm_probe3.py")).read().split("NOST, NOGL = 1 << 30, 1 << 29")[0]
exec(src)
for (M, N) in ((768, 768), (768, 3072), (3072, 768), (768, 2304), (768, 1536), (50260, 768)):
    for bn in (128, 256, 2256):
        for sk in ((1,) if M > 4096 else (1, 2, 4, 8)):
            tiles = ((M + (255 if bn > 2000 else 127)) // (256 if bn > 2000 else 128)) * ((N + (bn % 1000) - 1) // (bn % 1000)) * sk
            if tiles > 600 and sk > 1: continue
            run(M, N, 8192, 1, 1, bn, out_dtype=torch.float32, epilogue=L.EPI_ATOMIC, split_k=sk, iters=20, nbuf=3, tag="wgrad sk%d tiles%d" % (sk, tiles))

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gemm_probe3.py")).read().split("NOST, NOGL = 1 << 30, 1 << 29")[0]
exec(src)
for bn in (0, 256, 2256, 2128, 128):
    run(8192, 3072, 768, 0, 0, bn, tag="dgrad proj2 (bf16 out)")
    run(8192, 3072, 768, 0, 1, bn, tag="b1 variant")
for bn in (0, 256, 2256):
    run(8192, 768, 3072, 0, 0, bn, tag="dgrad fc")
    run(8192, 2304, 768, 0, 0, bn, tag="N2304 b0")

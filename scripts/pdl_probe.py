"""Does programmatic dependent launch pay on the training chain?  48 dependent ln_fwd launches in a graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ergm_b200 import ops
dev = "cuda"
M, H = 8192, 768
xs = [torch.randn(M, H, device=dev) for _ in range(4)]
g, b = torch.ones(H, device=dev), torch.zeros(H, device=dev)
y = torch.empty(M, H, device=dev, dtype=torch.bfloat16); y32 = torch.empty(M, H, device=dev)
mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
def chain():
    for i in range(48):
        # dependent chain: fp32 output of one call is the input of the next
        ops.ln_fwd(xs[i % 2] if i == 0 else xs[(i + 1) % 2 + 2], g, b, y, xs[i % 2 + 2], mean, rstd, 1e-5)
chain(); torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    chain()
gr.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): gr.replay()
e1.record(); torch.cuda.synchronize()
print("ERGM_PDL=%s: %.2f us per ln_fwd launch" % (os.environ.get("ERGM_PDL", "1"), e0.elapsed_time(e1) / 480 * 1e3))

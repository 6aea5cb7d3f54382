"""Per-kernel in-graph cost of the decode step: each op repeated R times inside one CUDA graph
(dependent chain on one stream, like the real step), time per launch = replay time / R."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ergm_b200 import ops
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B, H, I, V, nh, L = 64, 768, 3072, 50260, 12, 12
R = 48
bf = torch.bfloat16
x = torch.randn(B, H, device=dev)
gam, bet = torch.ones(H, device=dev), torch.zeros(H, device=dev)
def W(K, N, nk=False):
    w = 0.02 * torch.randn((N, K) if nk else (K, N), device=dev)
    return ops.dec_pack_weight(w, K, N, w_is_nk=nk)[0]
# distinct weights per repetition so that every launch streams from HBM like the real step (12 layers)
wq = [W(H, 3 * H) for _ in range(12)]; wo = [W(H, H) for _ in range(12)]
wf = [W(H, I) for _ in range(12)]; wp = [W(I, H) for _ in range(12)]
wl = W(H, V, nk=True)
hn = torch.randn(B, H, device=dev).to(bf); wte_b = (0.02 * torch.randn(V, H, device=dev)).to(bf)
bq, bo, bfc = torch.zeros(3 * H, device=dev), torch.zeros(H, device=dev), torch.zeros(I, device=dev)
qkv = torch.zeros(B, 3 * H, device=dev, dtype=bf); ctx = torch.zeros(B, H, device=dev, dtype=bf)
gbuf = torch.zeros(B, I, device=dev, dtype=bf); logits = torch.zeros(B, (V + 63) // 64 * 64, device=dev)
max_ctx = 192; pages = max_ctx // 16
pool = [torch.randn(B * pages, 2, nh, 16, 64, device=dev).to(bf) for _ in range(12)]
bt = torch.arange(B * pages, dtype=torch.int32, device=dev).view(B, pages).contiguous()
seq = torch.full((B,), 127, dtype=torch.int32, device=dev)
step = torch.zeros(1, dtype=torch.int32, device=dev); out_ids = torch.zeros(B, 4096, dtype=torch.int64, device=dev)
nxt = torch.zeros(B, 1, dtype=torch.int64, device=dev)
cases = {
    "ln_qkv   [64x768]x[768x2304]": lambda i: ops.dec_gemm(qkv, wq[i % 12], M=B, K=H, N=3 * H, x=x, bias=bq),
    "attn paged ctx=128": lambda i: ops.attn_decode_paged(qkv, pool[i % 12], bt, seq, ctx, B=B, nh=nh, H=H),
    "proj +=  [64x768]x[768x768]": lambda i: ops.dec_gemm(x, wo[i % 12], M=B, K=H, N=H, a=ctx, bias=bo, out_mode=2),
    "ln_fc gelu [64x768]x[768x3072]": lambda i: ops.dec_gemm(gbuf, wf[i % 12], M=B, K=H, N=I, x=x, bias=bfc, gelu=True),
    "proj2 += [64x3072]x[3072x768]": lambda i: ops.dec_gemm(x, wp[i % 12], M=B, K=I, N=H, a=gbuf, bias=bo, out_mode=2),
    "lm head  [64x768]x[768x50260]": lambda i: ops.dec_gemm(logits, wl, M=B, K=H, N=V, x=x, out_mode=1),
    "lm head tcgen05 gemm bn256": lambda i: ops.gemm(hn, wte_b, logits, M=B, N=V, K=H, a_major=0, b_major=0, block_n=256),
    "lm head tcgen05 gemm bn128": lambda i: ops.gemm(hn, wte_b, logits, M=B, N=V, K=H, a_major=0, b_major=0, block_n=128),
    "lm head tcgen05 gemm bn64": lambda i: ops.gemm(hn, wte_b, logits, M=B, N=V, K=H, a_major=0, b_major=0, block_n=64),
    "lm head ln_fwd rows": lambda i: ops.ln_fwd(x, gam, bet, hn, None, None, None, 1e-5),
    "sample argmax": lambda i: ops.sample(logits, V=V, step=step, out_ids=out_ids, next_ids=nxt, advance_step=True),
    "layer (5 kernels)": None,
}
def layer(i):
    for k in list(cases)[:5]:
        cases[k](i)
cases["layer (5 kernels)"] = layer
for name, fn in cases.items():
    x.normal_()
    reps = 8 if name.startswith("lm head") else R
    if len(sys.argv) > 1 and sys.argv[1] not in name: continue
    fn(0); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    print("%-34s %7.2f us / launch" % (name, e0.elapsed_time(e1) / 5 / reps * 1e3))
    seq.fill_(127); step.zero_()

"""Persistent cluster decode kernel (ergm_decode_stack) vs the per-kernel chain: token agreement + step time.
usage: python scripts/decode_stack_check.py [tiny|small] """
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transformers import GPT2Config
from ergm_b200.model import GPT2LMHeadModel
from ergm_b200 import generation, synthetic

which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
torch.manual_seed(0)
if which == "tiny":
    cfg = GPT2Config(vocab_size=1024, n_positions=256, n_embd=128, n_layer=2, n_head=2, initializer_range=0.2)
    B, T, new, vocab, fd = 5, 24, 12, 1024, 128
else:
    cfg = GPT2Config(vocab_size=50260, initializer_range=0.05)
    B, T, new, vocab, fd = 64, 128, 64, 50260, 768
m = GPT2LMHeadModel(cfg).to("cuda").eval()
b = synthetic.make_batch(B, T, seed=99, ragged=False, vocab=vocab, feat_dim=fd)
g = torch.Generator().manual_seed(7)
lens = torch.randint(T // 2, T + 1, (B,), generator=g).cuda()
ids, tt = b["input_ids"].cuda(), b["token_type_ids"].cuda()
outs = {}
for mode in ("0", "1"):
    os.environ["ERGM_DEC_STACK"] = mode
    for graph in (False, True):
        out, st = generation.generate(m, ids, tt, max_new_tokens=new, sp2_id=vocab - 1, prompt_lens=lens, return_state=True,
                                      use_cuda_graph=graph)
        torch.cuda.synchronize()
        outs[(mode, graph)] = out.cpu()
        print("mode", mode, "graph", graph, "stack used:", getattr(st, "stack", None) is not None, flush=True)
    if st.graph is not None:
        st.step.zero_(); st.seq_lens.copy_(lens.int())
        lat = []
        for _ in range(new - 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); st.graph.replay(); e1.record(); torch.cuda.synchronize()
            lat.append(e0.elapsed_time(e1))
        print("mode", mode, "decode step p50 %.1f us  min %.1f us" % (1e3 * statistics.median(lat), 1e3 * min(lat)), flush=True)
        tr = getattr(st, "trace", None)
        if tr is not None:
            t = tr.cpu().view(-1, 16).double()
            names = ["gridsync1", "ln1", "qkv_mma", "qkv_push+csync", "attn_loop", "attn_tail+push+csync", "oproj", "gridsync2",
                     "ln2", "fc_mma", "g_push+2csync", "proj"]
            d = (t[:, 1:13] - t[:, 0:12]) / 1e3
            print("per-phase us of CTA 0, median over blocks 1..L-1:")
            for i, n in enumerate(names):
                print("  %-22s %6.2f" % (n, d[1:, i].median().item()))
            print("  block total            %6.2f" % ((t[1:, 12] - t[1:, 0]) / 1e3).median().item())
ref = outs[("0", True)]
for k, v in outs.items():
    agree = (v == ref).float().mean().item()
    first = [(int((v[i] != ref[i]).nonzero()[0]) if (v[i] != ref[i]).any() else -1) for i in range(B)]
    print(k, "token agreement with the chain: %.3f" % agree, "first diffs", [f for f in first if f >= 0][:10])

"""Decode-step driver for ncu / CUDA-event timing: GPT-2 small, B=64, ragged prompts 64..128.
   python scripts/prof_decode.py [n_new] [--caption] [--events]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ergm_b200 import generation, ops
from oracle import synthetic
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
new = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4
model = bench.build(dev, dropout=0.0).eval()
g = torch.Generator().manual_seed(7)
b = synthetic.make_batch(64, 128, seed=99, ragged=False)
lens = torch.randint(64, 129, (64,), generator=g)
ids, tt, cap = b["input_ids"].to(dev), b["token_type_ids"].to(dev), b["caption_ids"].to(dev)
capt = cap if "--caption" in sys.argv else None
if "--bench" in sys.argv:
    import json
    print(json.dumps(bench.bench_generation(model, dev), indent=1))
elif "--events" in sys.argv:
    out, st = generation.generate(model, ids, tt, max_new_tokens=8, sp2_id=50259, caption_ids=capt, prompt_lens=lens,
                                  return_state=True, use_cuda_graph=False)
    torch.cuda.synchronize()
    ops.PROFILE = []
    for _ in range(3):
        generation.decode_step(model.engine, st, dict(top_k=0, temperature=1.0, seed=0, eos_id=-1))
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    by = {}
    for n, i, a, bb in prof:
        t, c = by.get(n, (0.0, 0))
        by[n] = (t + a.elapsed_time(bb) / 3, c + 1 / 3)
    for n, (t, c) in sorted(by.items(), key=lambda kv: -kv[1][0]):
        print("%-28s %8.1f us/step  %5.1f launches  %6.2f us each" % (n, t * 1e3, c, t * 1e3 / c))
else:
    generation.generate(model, ids, tt, max_new_tokens=new, sp2_id=50259, caption_ids=capt, prompt_lens=lens,
                        use_cuda_graph=False)
torch.cuda.synchronize()
print("done")

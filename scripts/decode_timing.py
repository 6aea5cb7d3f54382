"""Where does a decode step's time go?  CPU cost of a graph replay, GPU time of back-to-back replays,
eager step time, SM clocks while decoding."""
import os, sys, time, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ergm_b200 import generation, ops
from oracle import synthetic
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
model = bench.build(dev, dropout=0.0).eval()
g = torch.Generator().manual_seed(7)
b = synthetic.make_batch(64, 128, seed=99, ragged=False)
lens = torch.randint(64, 129, (64,), generator=g)
ids, tt, cap = b["input_ids"].to(dev), b["token_type_ids"].to(dev), b["caption_ids"].to(dev)
capt = cap if "--caption" in sys.argv else None
out, st = generation.generate(model, ids, tt, max_new_tokens=64, sp2_id=50259, caption_ids=capt, prompt_lens=lens,
                              return_state=True)
torch.cuda.synchronize()
clk = bench.ClockSampler(0); clk.start()
def reset():
    st.step.zero_(); st.seq_lens.copy_(lens.to(dev).int())
# warm the clocks
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
for _ in range(50): a @ a
torch.cuda.synchronize()
N = 60
for rep in range(3):
    reset(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(N): st.graph.replay()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("graph: cpu submit %.1f us/replay, gpu %.1f us/step, wall %.1f us/step" % ((t1 - t0) / N * 1e6, e0.elapsed_time(e1) / N * 1e3, (t2 - t0) / N * 1e6))
kw = dict(top_k=0, temperature=1.0, seed=0, eos_id=-1)
for rep in range(2):
    reset(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(N): generation.decode_step(model.engine, st, kw)
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print("eager: cpu submit %.1f us/step, gpu %.1f us/step" % ((t1 - t0) / N * 1e6, e0.elapsed_time(e1) / N * 1e3))
print(clk.stop())

import os, sys, faulthandler
faulthandler.dump_traceback_later(40, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import ergm_oracle as O, synthetic
from test_generation_gpu import build_model, tiny_cfg
from ergm_b200 import generation, ops
import ergm_b200.ops as OPS

orig = OPS._call
def traced(name, *a):
    orig(name, *a)
    torch.cuda.synchronize()
    print("ok", name, flush=True)
OPS._call = traced
og = OPS.gemm
def tg(*a, **k):
    og(*a, **k); torch.cuda.synchronize(); print("ok gemm", k.get("M"), k.get("N"), k.get("K"), flush=True)
OPS.gemm = tg
generation.ops.gemm = tg

cfg = tiny_cfg()
sd = O.init_state_dict(cfg, seed=5, perturb=True)
m = build_model(cfg, sd)
b = synthetic.make_batch(4, 24, seed=21, vocab=cfg.vocab_size, feat_dim=cfg.n_embd, ragged=False)
ids = generation.generate(m, b["input_ids"].cuda(), b["token_type_ids"].cuda(), max_new_tokens=4, sp2_id=cfg.vocab_size - 1, use_cuda_graph=False)
print(ids.cpu())

"""2-rank DP debug driver with stage prints (run under torchrun)."""
import os, sys, faulthandler
faulthandler.dump_traceback_later(70, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from transformers import GPT2Config
from ergm_b200.model import GPT2LMHeadModel
from ergm_b200.optim import FusedAdamW
from ergm_b200.parallel import DataParallel
from ergm_b200.trainer import GraphedTrainStep
from oracle import ergm_oracle as O, synthetic
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
def say(*a):
    print("[r%d]" % rank, *a, flush=True)
cfg = O.OracleConfig(vocab_size=1024, n_positions=256, n_embd=128, n_layer=3, n_head=2)
sd = O.init_state_dict(cfg, seed=3, perturb=True)
def build():
    hf = GPT2Config(vocab_size=1024, n_positions=256, n_embd=128, n_layer=3, n_head=2, attn_pdrop=0.0, resid_pdrop=0.0, embd_pdrop=0.0)
    m = GPT2LMHeadModel(hf); m.load_state_dict(sd); return m.cuda().train()
b = synthetic.make_batch(4 * world, 64, seed=31, vocab=1024, feat_dim=128)
keys = ("input_ids", "token_type_ids", "labels", "emotion_labels", "caption_ids", "imgs", "auds")
full = {k: b[k].cuda() for k in keys}
mine = {k: v[rank * 4:(rank + 1) * 4].contiguous() for k, v in full.items()}
ref = build(); out = ref(**full); out.loss.backward(); torch.cuda.synchronize(); say("ref loss", out.loss.item())
m = build(); dp = DataParallel(m, bucket_mb=0.25); torch.cuda.synchronize(); say("dp built", len(dp.buckets))
o = m(**mine); torch.cuda.synchronize(); say("dp fwd", o.loss.item())
o.loss.backward(); torch.cuda.synchronize(); say("dp bwd done")
worst = max(((p.grad - q.grad).norm() / (q.grad.norm() + 1e-20)).item() for p, q in zip(m.parameters(), ref.parameters()))
say("worst grad rel", worst)
if "--graph" in sys.argv:
    m2 = build(); dp2 = DataParallel(m2, bucket_mb=0.25)
    step = GraphedTrainStep(m2, FusedAdamW(m2, lr=1e-3), dp=dp2)
    pinned = {k: v.cpu().pin_memory() for k, v in mine.items()}
    for i in range(3):
        l = step(pinned); say("graph step", i, l)
    step.close(); say("closed")
dist.destroy_process_group()
say("done")

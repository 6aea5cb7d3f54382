"""Multi-GPU data-parallel parity (needs >= 2 GPUs; skipped otherwise): an N-rank step on a split
batch must equal the single-GPU step on the concatenated batch (loss, gradients, post-step
weights), SURVEY.md §8e.  Spawns torchrun-style workers itself."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from transformers import GPT2Config
from ergm_b200.model import GPT2LMHeadModel
from ergm_b200.optim import FusedAdamW
from ergm_b200.parallel import DataParallel
from ergm_b200.trainer import GraphedTrainStep
from oracle import ergm_oracle as O, synthetic
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
cfg = O.OracleConfig(vocab_size=1024, n_positions=256, n_embd=128, n_layer=3, n_head=2)
sd = O.init_state_dict(cfg, seed=3, perturb=True)
def build():
    hf = GPT2Config(vocab_size=1024, n_positions=256, n_embd=128, n_layer=3, n_head=2, attn_pdrop=0.0, resid_pdrop=0.0, embd_pdrop=0.0)
    m = GPT2LMHeadModel(hf); m.load_state_dict(sd); return m.cuda().train()
B = 4 * world
b = synthetic.make_batch(B, 64, seed=31, vocab=1024, feat_dim=128)
keys = ("input_ids", "token_type_ids", "labels", "emotion_labels", "caption_ids", "imgs", "auds")
full = {k: b[k].cuda() for k in keys}
mine = {k: v[rank * 4:(rank + 1) * 4].contiguous() for k, v in full.items()}
# reference: single-GPU step on the concatenated batch (every rank computes it redundantly)
ref = build()
out = ref(**full); out.loss.backward()
ref_loss = out.loss.item()
ref_grads = {n: p.grad.clone() for n, p in ref.named_parameters()}
FusedAdamW(ref, lr=1e-3).step()
# data-parallel step through the public model API
m = build()
dp = DataParallel(m, bucket_mb=0.25)
o = m(**mine); o.loss.backward()
ok = abs(o.loss.item() - ref_loss) < 1e-5
worst = 0.0
for n, p in m.named_parameters():
    r = ((p.grad - ref_grads[n]).norm() / (ref_grads[n].norm() + 1e-20)).item()
    worst = max(worst, r)
FusedAdamW(m, lr=1e-3).step()
wdiff = max((p.detach() - q.detach()).abs().max().item() for p, q in zip(m.parameters(), ref.parameters()))
# graph-captured DP train step (the bench path) runs and agrees with itself across ranks
m2 = build(); dp2 = DataParallel(m2, bucket_mb=0.25)
step = GraphedTrainStep(m2, FusedAdamW(m2, lr=1e-3), dp=dp2)
pinned = {k: v.cpu().pin_memory() for k, v in mine.items()}
losses = [step(pinned) for _ in range(3)]
t = torch.tensor(losses, device="cuda"); t2 = t.clone(); dist.broadcast(t2, 0)
same = bool(torch.equal(t, t2))
w0 = m2.transformer.h[1].mlp.c_fc.weight.detach().clone(); w1 = w0.clone(); dist.broadcast(w1, 0)
print("RANK%%d loss_ok=%%s worst_grad_rel=%%.2e wdiff=%%.2e graph_losses=%%s same=%%s wsync=%%s" %% (rank, ok, worst, wdiff, ["%%.4f" %% x for x in losses], same, bool(torch.equal(w0, w1))), flush=True)
assert ok and worst < 2e-3 and wdiff < 1e-5 and same and torch.equal(w0, w1) and losses[2] < losses[0]
step.close()
dist.barrier()
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2])
def test_dp_step_equals_single_gpu_step(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    script = tmp_path / "dp_worker.py"
    script.write_text(WORKER % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0

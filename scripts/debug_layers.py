"""Layer-by-layer comparison of the engine's saved activations with the oracle (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ergm_oracle as O, synthetic
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_model_gpu import build_model, tiny_cfg, cuda_batch

cfg = tiny_cfg()
sd = O.init_state_dict(cfg, seed=3, perturb=True)
m = build_model(cfg, sd).train()
b = synthetic.make_batch(3, 48, seed=11, vocab=cfg.vocab_size, feat_dim=cfg.n_embd, tc=40)
kw = cuda_batch(b, caption=True)
eng = m.engine
out = eng.forward(kw["input_ids"], kw["token_type_ids"], kw["labels"], kw["emotion_labels"], kw["imgs"], kw["auds"],
                  kw["caption_ids"], None, training=True, save=True, dropout=0.0)
torch.cuda.synchronize()
sv = eng.saved
B, T, H = 3, 48, cfg.n_embd

def stat(name, t, ref=None):
    t = t.float()
    msg = "%-10s nan=%d absmax=%.4g" % (name, int(torch.isnan(t).sum()), float(t.abs().max()))
    if ref is not None:
        ref = ref.to(t.device).float().reshape(t.shape)
        msg += "  relerr=%.3g" % float((t - ref).norm() / (ref.norm() + 1e-30))
    print(msg)

# oracle intermediates
wte, wpe = sd["transformer.wte.weight"], sd["transformer.wpe.weight"]
import torch.nn.functional as F
e = F.embedding(b["input_ids"], wte).clone()
for i in range(B):
    e[i, 0] += b["imgs"][i][0]
    e[i, 1] += b["auds"][i]
h = e + F.embedding(torch.arange(T)[None], wpe) + F.embedding(b["token_type_ids"], wte)
enc = F.embedding(b["caption_ids"], wte)
for l in range(cfg.n_layer):
    r = sv["layers"][l]
    p = "transformer.h.%d." % l
    stat("x_%d" % l, r["x"], h)
    a1 = O.layer_norm(sd, p + "ln_1.", cfg, h)
    stat("a1", r["a1"], a1)
    qkv = O.conv1d(a1, sd[p + "attn.c_attn.weight"], sd[p + "attn.c_attn.bias"])
    stat("qkv", r["qkv"], qkv)
    ao, _ = O.self_attention(sd, p + "attn.", cfg, a1)
    stat("x1", r["x1"], ao + h)
    stat("ctx", r["ctx"])
    h1 = ao + h
    a2 = O.layer_norm(sd, p + "ln_cross_attn.", cfg, h1)
    stat("a2", r["a2"], a2)
    stat("q2", r["q2"], O.conv1d(a2, sd[p + "crossattention.q_attn.weight"], sd[p + "crossattention.q_attn.bias"]))
    stat("kv2", r["kv2"], O.conv1d(enc, sd[p + "crossattention.c_attn.weight"], sd[p + "crossattention.c_attn.bias"]))
    stat("ctx2", r["ctx2"])
    co = O.cross_attention(sd, p + "crossattention.", cfg, a2, enc)
    h2 = h1 + co
    stat("x2", r["x2"], h2)
    a3 = O.layer_norm(sd, p + "ln_2.", cfg, h2)
    stat("a3", r["a3"], a3)
    u = O.conv1d(a3, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])
    stat("u", r["u"], u)
    stat("g", r["g"], O.gelu_new(u))
    h = h2 + O.mlp(sd, p + "mlp.", a3)
stat("xf", sv["xf"], h)
stat("hn", sv["hn"], O.layer_norm(sd, "transformer.ln_f.", cfg, h))
stat("logits", sv["logits"][:, :cfg.vocab_size], F.linear(O.layer_norm(sd, "transformer.ln_f.", cfg, h), wte))
print("sums", out["loss_sums"].tolist())
print("losses", eng.finalize_loss(out).tolist())
ref = O.forward(sd, cfg, b["input_ids"], b["token_type_ids"], b["labels"], b["emotion_labels"], b["imgs"], b["auds"], b["caption_ids"])
print("ref loss", float(ref["loss"]), float(ref["lm_loss"]), float(ref["emotion_loss"]))

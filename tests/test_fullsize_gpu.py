"""Parity at the sizes bench.py times (BASELINE.json configs 2, 4, 5), against the oracle restatement (fp32, CPU).

The CUDA path is driven through the public model class (C ABI underneath); the oracle runs on the GPU box's
host cores and finishes in seconds at these sizes.  Tolerances are north_star's: LM loss 1e-3, logits 1e-2
norm-relative, gradients 3e-2 norm-relative per tensor (bf16-operand / fp32-accumulate mode); greedy ids
bit-exact in fp32 mode, and in bf16 mode every first divergence must be an fp32 near-tie of the oracle.
"""
import os

import pytest
import torch

from oracle import ergm_oracle as O
from ergm_b200 import synthetic

pytestmark = pytest.mark.gpu
LOGITS_REL_TOL = 1e-2
LOSS_TOL = 1e-3
GRAD_TOL = 3e-2


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def build_model(cfg, sd, dropout=0.0):
    from transformers import GPT2Config
    from ergm_b200.model import GPT2LMHeadModel
    hf = GPT2Config(vocab_size=cfg.vocab_size, n_positions=cfg.n_positions, n_embd=cfg.n_embd, n_layer=cfg.n_layer,
                    n_head=cfg.n_head, attn_pdrop=dropout, resid_pdrop=dropout, embd_pdrop=dropout,
                    initializer_range=cfg.initializer_range)
    m = GPT2LMHeadModel(hf)
    m.load_state_dict(sd, strict=True)
    return m.to("cuda")


def _oracle_sd(sd):
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "lm_head.weight"}
    sdo["lm_head.weight"] = sdo["transformer.wte.weight"]
    return sdo


def test_config2_exact_training_step_vs_oracle(cuda_device):
    """BASELINE config 2, the step bench.py times: GPT-2 small (12 layers, V = 50260), B = 32, T = Tc = 256,
    caption mode + img/aud fusion, the synthetic MELD-shaped batch of bench.py (seed 1234), p = 0.
    Forward, both losses, logits at every position, and a dozen gradients spread over the depth."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cfg = O.OracleConfig()
    sd = O.init_state_dict(cfg, seed=0, perturb=True)
    m = build_model(cfg, sd).train()
    b = synthetic.make_batch(32, 256, seed=1234)
    kw = dict(input_ids=b["input_ids"].cuda(), token_type_ids=b["token_type_ids"].cuda(), labels=b["labels"].cuda(),
              emotion_labels=b["emotion_labels"].cuda(), caption_ids=b["caption_ids"].cuda(), imgs=b["imgs"].cuda(),
              auds=b["auds"].cuda())
    out = m(**kw)
    out.loss.backward()
    sdo = _oracle_sd(sd)
    o = O.forward(sdo, cfg, b["input_ids"], b["token_type_ids"], b["labels"], b["emotion_labels"], b["imgs"], b["auds"],
                  b["caption_ids"])
    o["loss"].backward()
    d_lm = abs(out.lm_loss.item() - o["lm_loss"].item())
    d_sum = abs(out.loss.item() - o["loss"].item())
    r_log = rel(out.logits, o["logits"])
    print("config 2: lm loss %.6f (oracle %.6f, |d| %.2e)  summed loss |d| %.2e  logits rel %.2e  emotion logits rel %.2e"
          % (out.lm_loss.item(), o["lm_loss"].item(), d_lm, d_sum, r_log, rel(out.emotion_logits, o["emotion_logits"])))
    assert d_lm < LOSS_TOL
    assert d_sum < LOSS_TOL          # 32 samples: the emotion CE mean is no longer dominated by single-sample bf16 noise
    assert r_log < LOGITS_REL_TOL
    names = ["transformer.wte.weight", "transformer.wpe.weight", "transformer.ln_f.weight", "emotion_head.weight",
             "transformer.h.0.attn.c_attn.weight", "transformer.h.0.ln_1.bias", "transformer.h.3.crossattention.q_attn.weight",
             "transformer.h.5.mlp.c_fc.weight", "transformer.h.5.mlp.c_fc.bias", "transformer.h.7.crossattention.c_attn.weight",
             "transformer.h.9.attn.c_proj.weight", "transformer.h.11.mlp.c_proj.weight", "transformer.h.11.ln_cross_attn.weight",
             "transformer.h.11.attn.c_attn.bias"]
    params = dict(m.named_parameters())
    worst = ("", 0.0)
    for n in names:
        r = rel(params[n].grad, sdo[n].grad)
        worst = max(worst, (n, r), key=lambda t: t[1])
        assert r < GRAD_TOL, (n, r)
    print("config 2: worst gradient rel err %.2e (%s)" % (worst[1], worst[0]))


def _oracle_greedy_ragged(sd, cfg, ids, tt, lens, new, sp2, caption_ids=None):
    """Batched KV-cached greedy decode of right-padded ragged prompts with the oracle: the reference's own
    attention_mask / position_ids surface (model.py:469-482) makes a padded batch equal to per-sequence runs."""
    B, T = ids.shape
    ar = torch.arange(T)[None]
    mask = (ar < lens[:, None]).long()
    with torch.no_grad():
        r = O.forward(sd, cfg, ids, tt, caption_ids=caption_ids, attention_mask=mask)
        past = r["past_key_values"]
        logits = r["logits"][torch.arange(B), lens - 1]
        out = []
        for step in range(new):
            nxt = logits.argmax(-1)
            out.append(nxt)
            if step + 1 == new:
                break
            mask = torch.cat([mask, torch.ones(B, 1, dtype=torch.long)], 1)
            pos = (lens + step)[:, None]
            r = O.forward(sd, cfg, nxt[:, None], torch.full((B, 1), sp2), caption_ids=caption_ids, past_key_values=past,
                          attention_mask=mask, position_ids=pos)
            past = r["past_key_values"]
            logits = r["logits"][:, -1]
    return torch.stack(out, 1)


NEAR_TIE = 0.04  # units of the row's logit std (bf16-mode logits are within 1e-2 relative of the oracle's)


def _check_divergences(got, want, sd, cfg, ids, tt, lens, sp2, caption_ids=None):
    """Agreeing-prefix fraction; every FIRST divergence must be a near-tie in the oracle's own fp32 logits."""
    got, want = got.cpu(), want.cpu()
    B, N = want.shape
    same, flips = 0, 0
    for i in range(B):
        neq = (got[i] != want[i]).nonzero()
        first = int(neq[0]) if len(neq) else N
        same += first
        if first < N:
            flips += 1
            n = int(lens[i])
            s_ids = torch.cat([ids[i, :n], want[i, :first]])[None]
            s_tt = torch.cat([tt[i, :n], torch.full((first,), sp2)])[None]
            with torch.no_grad():
                lg = O.forward(sd, cfg, s_ids, s_tt, caption_ids=caption_ids[i:i + 1] if caption_ids is not None else None)["logits"][0, -1]
            margin = (lg[want[i, first]] - lg[got[i, first]]).item()
            assert 0 <= margin <= NEAR_TIE * lg.std().item(), \
                "sequence %d diverges at token %d on a clear arg-max (margin %.4f, std %.4f)" % (i, first, margin, lg.std().item())
    return same / (B * N), flips


@pytest.mark.parametrize("caption", [False, True])
def test_config4_generation_vs_oracle(cuda_device, caption):
    """BASELINE config 4 as bench.py runs it: GPT-2 small, 64 requests, ragged prompts 64..128, 64 new tokens,
    greedy, paged KV + CUDA-graph decode, with and without captions.  A larger initializer_range keeps random-init
    greedy decoding from collapsing onto one token."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cfg = O.OracleConfig()
    cfg.initializer_range = 0.05
    sd = O.init_state_dict(cfg, seed=5, perturb=True)
    m = build_model(cfg, sd).eval()
    B, prompt, new = 64, 128, 64
    g = torch.Generator().manual_seed(7)
    b = synthetic.make_batch(B, prompt, seed=99, ragged=False)
    lens = torch.randint(prompt // 2, prompt + 1, (B,), generator=g)
    cap = b["caption_ids"][:, :64].contiguous() if caption else None
    ids = m.generate(b["input_ids"].cuda(), b["token_type_ids"].cuda(), max_new_tokens=new, sp2_id=50259,
                     caption_ids=cap.cuda() if caption else None, prompt_lens=lens.cuda())
    want = _oracle_greedy_ragged(sd, cfg, b["input_ids"], b["token_type_ids"], lens, new, 50259, caption_ids=cap)
    frac, flips = _check_divergences(ids, want, sd, cfg, b["input_ids"], b["token_type_ids"], lens, 50259, caption_ids=cap)
    print("config 4 (%s): agreeing-prefix fraction %.3f, %d of %d sequences flip on an oracle near-tie, distinct tokens %d"
          % ("caption" if caption else "no caption", frac, flips, B, want.unique().numel()))
    assert want.unique().numel() > 8, "degenerate decode: the test would prove nothing"
    assert frac >= 0.5, frac


def test_config4_fp32_mode_greedy_bit_exact(cuda_device):
    """north_star: greedy-decoded token ids bit-exact in fp32 mode.  Config-4 model size, 16 uniform prompts of 96
    tokens, 24 new tokens (fp32 mode recomputes the full sequence per token like main.py:253-282)."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cfg = O.OracleConfig()
    cfg.initializer_range = 0.05
    sd = O.init_state_dict(cfg, seed=5, perturb=True)
    m = build_model(cfg, sd).eval()
    m.ergm_precision = "fp32"
    B, prompt, new = 16, 96, 24
    b = synthetic.make_batch(B, prompt, seed=98, ragged=False)
    ids = m.generate(b["input_ids"].cuda(), b["token_type_ids"].cuda(), max_new_tokens=new, sp2_id=50259)
    with torch.no_grad():
        want = O.greedy_generate_cached(sd, cfg, b["input_ids"], b["token_type_ids"], new, sp2_id=50259, eos_id=-1)
    eq = (ids.cpu() == want)
    if not bool(eq.all()):
        # a genuine fp32-level tie is the only admissible difference: verify it is one (margin < 1e-4)
        lens = torch.full((B,), prompt)
        old = globals()["NEAR_TIE"]
        try:
            globals()["NEAR_TIE"] = 2e-4
            _check_divergences(ids, want, sd, cfg, b["input_ids"], b["token_type_ids"], lens, 50259)
        finally:
            globals()["NEAR_TIE"] = old
    print("config 4 fp32 mode: %d / %d tokens identical" % (int(eq.sum()), eq.numel()))
    assert bool(eq.all()), "fp32-mode greedy ids differ from the oracle"


def test_config5_medium_full_depth_forward_vs_oracle(cuda_device):
    """BASELINE config 5 backbone at full depth: GPT-2 medium (24 layers, H = 1024, 16 heads, V = 50260),
    B = 2 x T = 512, caption mode, forward + losses + logits against the oracle."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cfg = O.OracleConfig(n_embd=1024, n_layer=24, n_head=16)
    sd = O.init_state_dict(cfg, seed=2, perturb=True)
    m = build_model(cfg, sd).eval()
    b = synthetic.make_batch(2, 512, seed=15, feat_dim=1024, tc=512)
    b["labels"] = b["input_ids"].clone()
    kw = dict(input_ids=b["input_ids"].cuda(), token_type_ids=b["token_type_ids"].cuda(), labels=b["labels"].cuda(),
              emotion_labels=b["emotion_labels"].cuda(), caption_ids=b["caption_ids"].cuda(), imgs=b["imgs"].cuda(),
              auds=b["auds"].cuda())
    with torch.no_grad():
        out = m(**kw)
        o = O.forward(sd, cfg, b["input_ids"], b["token_type_ids"], b["labels"], b["emotion_labels"], b["imgs"],
                      b["auds"], b["caption_ids"])
    d_lm = abs(out.lm_loss.item() - o["lm_loss"].item())
    r_log = rel(out.logits, o["logits"])
    print("config 5 (medium, 24 layers): lm loss |d| %.2e, logits rel %.2e" % (d_lm, r_log))
    assert d_lm < LOSS_TOL
    assert r_log < LOGITS_REL_TOL
    assert rel(out.emotion_logits, o["emotion_logits"]) < 2e-2

"""CPU tests: the oracle restatement against the committed golden fixtures (generated from
the unmodified reference by oracle/make_golden.py) and, when the reference tree is present
(build container only), directly against the reference."""
import os

import numpy as np
import pytest
import torch

from oracle import ergm_oracle as O
from oracle import ref_shim, synthetic

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def tiny_cfg():
    return O.OracleConfig(vocab_size=1024, n_positions=256, n_embd=128, n_layer=2, n_head=2)


@pytest.mark.parametrize("mode", ["caption", "nocaption"])
def test_oracle_matches_tiny_golden(mode):
    g = np.load(os.path.join(GOLD, "tiny.npz"))
    cfg = tiny_cfg()
    sd = {k: v.clone().requires_grad_(True) for k, v in O.init_state_dict(cfg, seed=3, perturb=True).items()
          if k != "lm_head.weight"}
    sd["lm_head.weight"] = sd["transformer.wte.weight"]
    b = synthetic.make_batch(3, 48, seed=11, vocab=cfg.vocab_size, feat_dim=cfg.n_embd, tc=40)
    cap = torch.from_numpy(g[mode + "/caption_ids"]) if mode == "caption" else None
    o = O.forward(sd, cfg, b["input_ids"], b["token_type_ids"], b["labels"], b["emotion_labels"],
                  b["imgs"], b["auds"], cap)
    np.testing.assert_allclose(o["logits"].detach().numpy(), g[mode + "/logits"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(o["loss"].item(), g[mode + "/loss"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(o["emotion_logits"].detach().numpy(), g[mode + "/emotion_logits"], atol=2e-6)
    np.testing.assert_allclose(o["past_key_values"][0][0].detach().numpy(), g[mode + "/present0_k"], atol=2e-6)
    o["loss"].backward()
    for k in ("transformer.wte.weight", "transformer.h.1.mlp.c_fc.weight", "transformer.h.0.attn.c_attn.bias",
              "transformer.ln_f.weight", "emotion_head.weight"):
        np.testing.assert_allclose(sd[k].grad.numpy(), g["%s/grad/%s" % (mode, k)], rtol=0, atol=1e-6)


def test_oracle_generation_matches_golden():
    g = np.load(os.path.join(GOLD, "tiny_generate.npz"))
    cfg = tiny_cfg()
    cfg.initializer_range = 0.2
    sd = O.init_state_dict(cfg, seed=5, perturb=True)
    b = synthetic.make_batch(4, 24, seed=21, vocab=cfg.vocab_size, feat_dim=cfg.n_embd, ragged=False)
    with torch.no_grad():
        ids = O.greedy_generate_cached(sd, cfg, b["input_ids"], b["token_type_ids"], 12,
                                       sp2_id=cfg.vocab_size - 1, eos_id=-1)
        ids2 = O.greedy_generate_recompute(sd, cfg, b["input_ids"], b["token_type_ids"], 12,
                                           sp2_id=cfg.vocab_size - 1, eos_id=-1)
    assert np.array_equal(ids.numpy(), g["greedy_ids"])
    assert np.array_equal(ids2.numpy(), g["greedy_ids"])


def test_oracle_small_gv1_nocaption_golden():
    """BASELINE config 1 (GPT-2 small, B=4, T=128) against the reference-generated fixture."""
    g = np.load(os.path.join(GOLD, "small_gv1.npz"))
    cfg = O.OracleConfig()
    sd = O.init_state_dict(cfg, seed=0, perturb=True)
    b = synthetic.gv1_inputs()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        o = O.forward(sd, cfg, b["input_ids"], b["token_type_ids"], b["labels"], b["emotion_labels"],
                      b["imgs"], b["auds"], b["caption_ids"])
    assert abs(o["loss"].item() - float(g["caption/loss"])) < 1e-5
    np.testing.assert_allclose(o["logits"][3, 127].numpy(), g["caption/logits_b3_t127"], atol=2e-5)
    np.testing.assert_allclose(o["emotion_logits"].numpy(), g["caption/emotion_logits"], atol=2e-5)
    assert abs(o["logits"].double().sum().item() - float(g["caption/logits_sum"])) < 0.5


def test_top_p_filter_shift_rule():
    """main.py:263-265: the token that crosses top_p is kept (mask shifted right by one)."""
    p = torch.tensor([[0.5, 0.3, 0.15, 0.05]])
    out = O.top_p_filter_reference(p, 0.7)
    assert torch.allclose(out, torch.tensor([[0.625, 0.375, 0.0, 0.0]]))


def test_adamw_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(1000)
    g = torch.randn(1000)
    pt = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pt], lr=2e-5)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    q = p.clone()
    for step in range(1, 4):
        pt.grad = g.clone()
        opt.step()
        q, m, v = O.adamw_step(q, g, m, v, step, 2e-5)
    assert torch.allclose(q, pt.detach(), atol=1e-7)


def test_poly_schedule_matches_transformers():
    from transformers import get_polynomial_decay_schedule_with_warmup
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=2e-5)
    sched = get_polynomial_decay_schedule_with_warmup(opt, 10, 100, power=2)
    for step in range(0, 100, 7):
        want = O.poly_decay_lr(step, 2e-5, 10, 100)
        got = 2e-5 * sched.lr_lambdas[0](step)
        assert abs(want - got) < 1e-12


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
def test_oracle_bit_equal_to_reference_block():
    """Block-level parity vs the reference's own GPT2Block with arbitrary encoder length."""
    from transformers import GPT2Config
    mod = ref_shim.load()
    cfg = tiny_cfg()
    sd = O.init_state_dict(cfg, seed=9, perturb=True)
    hf = GPT2Config(vocab_size=1024, n_positions=256, n_embd=128, n_layer=2, n_head=2,
                    attn_pdrop=0, resid_pdrop=0, embd_pdrop=0)
    blk = mod.GPT2Block(hf, layer_idx=0).eval()
    blk.load_state_dict({k[len("transformer.h.0."):]: v for k, v in sd.items() if k.startswith("transformer.h.0.")},
                        strict=False)
    x = torch.randn(2, 17, 128)
    enc = torch.randn(2, 9, 128)
    with torch.no_grad():
        want = blk(x, encoder_hidden_states=enc, encoder_attention_mask=torch.zeros(2, 1, 1, 9))[0]
        got, _ = O.block(sd, 0, cfg, x, enc, enc_mask=torch.zeros(2, 1, 1, 9))
    assert torch.equal(want, got)


def test_modality_pool_proj_extension_restates_offline_pooling():
    """A3 extension oracle: time-mean (feature_extraction.py:63,69) then Linear; with an identity
    projection it must reproduce the reference's pooled-feature layout exactly."""
    g = torch.Generator().manual_seed(5)
    vis, aud = torch.randn(3, 197, 16, generator=g), torch.randn(3, 113, 16, generator=g)
    sd = {"visual_proj.weight": torch.eye(16), "visual_proj.bias": torch.zeros(16),
          "audio_proj.weight": torch.eye(16), "audio_proj.bias": torch.zeros(16)}
    v, a = O.modality_pool_proj(sd, vis, aud)
    assert v.shape == (3, 1, 16) and a.shape == (3, 16)
    assert torch.allclose(v[:, 0], vis.mean(1), atol=1e-7) and torch.allclose(a, aud.mean(1), atol=1e-7)
    cfg = O.OracleConfig(vocab_size=64, n_positions=32, n_embd=32, n_layer=1, n_head=2, visual_dim=16, audio_dim=16)
    keys = dict(O.param_shapes(cfg))
    assert keys["visual_proj.weight"] == (32, 16) and keys["audio_proj.bias"] == (32,)
    assert "visual_proj.weight" not in dict(O.param_shapes(O.OracleConfig(vocab_size=64, n_embd=32, n_layer=1, n_head=2)))

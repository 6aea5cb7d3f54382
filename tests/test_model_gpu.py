"""GPU parity of the drop-in model (ergm_b200.model.GPT2LMHeadModel, through the C ABI) against
the oracle restatement (oracle/ergm_oracle.py, fp32 on CPU) and the reference-generated golden
fixtures in tests/golden/ (oracle/make_golden.py).

Tolerances are those of BASELINE.json north_star for the bf16-operand / fp32-accumulate mode:
logits within 1e-2 relative (norm-wise) error, loss within 1e-3."""
import os

import numpy as np
import pytest
import torch

from oracle import ergm_oracle as O
from oracle import synthetic

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
LOGITS_REL_TOL = 1e-2
LOSS_TOL = 1e-3


def tiny_cfg():
    return O.OracleConfig(vocab_size=1024, n_positions=256, n_embd=128, n_layer=2, n_head=2)


def build_model(cfg, sd, dropout=0.0):
    from transformers import GPT2Config
    from ergm_b200.model import GPT2LMHeadModel
    hf = GPT2Config(vocab_size=cfg.vocab_size, n_positions=cfg.n_positions, n_embd=cfg.n_embd, n_layer=cfg.n_layer,
                    n_head=cfg.n_head, attn_pdrop=dropout, resid_pdrop=dropout, embd_pdrop=dropout,
                    initializer_range=cfg.initializer_range)
    if getattr(cfg, "visual_dim", None):
        hf.ergm_visual_dim, hf.ergm_audio_dim = cfg.visual_dim, cfg.audio_dim
    m = GPT2LMHeadModel(hf)
    m.load_state_dict(sd, strict=True)
    return m.to("cuda")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def cuda_batch(b, caption=True, fusion=True):
    kw = dict(input_ids=b["input_ids"].cuda(), token_type_ids=b["token_type_ids"].cuda(),
              labels=b["labels"].cuda(), emotion_labels=b["emotion_labels"].cuda())
    if caption:
        kw["caption_ids"] = b["caption_ids"].cuda()
    if fusion:
        kw["imgs"] = b["imgs"].cuda()
        kw["auds"] = b["auds"].cuda()
    return kw


@pytest.mark.parametrize("mode", ["caption", "nocaption"])
def test_tiny_forward_backward_vs_golden(cuda_device, mode):
    g = np.load(os.path.join(GOLD, "tiny.npz"))
    cfg = tiny_cfg()
    sd = O.init_state_dict(cfg, seed=3, perturb=True)
    m = build_model(cfg, sd).train()  # dropout p = 0: train mode exercises the save / backward path
    b = synthetic.make_batch(3, 48, seed=11, vocab=cfg.vocab_size, feat_dim=cfg.n_embd, tc=40)
    kw = cuda_batch(b, caption=(mode == "caption"))
    if mode == "caption":
        kw["caption_ids"] = torch.from_numpy(g[mode + "/caption_ids"]).cuda()
    out = m(**kw)
    assert abs(out.loss.item() - float(g[mode + "/loss"])) < LOSS_TOL
    assert rel(out.logits, torch.from_numpy(g[mode + "/logits"])) < LOGITS_REL_TOL
    assert rel(out.emotion_logits, torch.from_numpy(g[mode + "/emotion_logits"])) < LOGITS_REL_TOL
    k0 = out.past_key_values[0][0]
    assert tuple(k0.shape) == g[mode + "/present0_k"].shape
    assert rel(k0, torch.from_numpy(g[mode + "/present0_k"])) < LOGITS_REL_TOL
    out.loss.backward()
    worst = 0.0
    for name, p in m.named_parameters():
        key = "%s/grad/%s" % (mode, name)
        if key not in g.files:
            assert p.grad is None or p.grad.abs().max().item() == 0.0, name
            continue
        ref = torch.from_numpy(g[key])
        r = rel(p.grad, ref)
        worst = max(worst, r)
        # The fixture uses N(0, 0.02) biases, so every caption key is "bias + small": the cross-attention
        # query-side gradient is a difference of nearly equal bf16 products and carries more rounding noise.
        tol = 6e-2 if ("crossattention" in name or "ln_cross_attn" in name) else 3e-2
        assert r < tol, (name, r)
    print("worst grad rel err", worst)


def test_tiny_vs_oracle_fp64_direct(cuda_device):
    """Same check against the oracle evaluated in fp64 on this host (no fixture involved), with
    ragged captions (Tc != T, boundary decision 2) and no fusion."""
    cfg = tiny_cfg()
    sd = O.init_state_dict(cfg, seed=4, perturb=True)
    m = build_model(cfg, sd).eval()
    b = synthetic.make_batch(2, 37, seed=5, vocab=cfg.vocab_size, feat_dim=cfg.n_embd, tc=19)
    kw = cuda_batch(b, fusion=False)
    with torch.no_grad():
        out = m(**kw)
    sd64 = {k: v.double() for k, v in sd.items()}
    o = O.forward(sd64, cfg, b["input_ids"], b["token_type_ids"], b["labels"], b["emotion_labels"],
                  caption_ids=b["caption_ids"])
    assert abs(out.loss.item() - o["loss"].item()) < LOSS_TOL
    assert rel(out.logits, o["logits"]) < LOGITS_REL_TOL
    assert abs(out.lm_loss.item() - o["lm_loss"].item()) < LOSS_TOL


def test_emotion_only_and_lm_only_losses(cuda_device):
    cfg = tiny_cfg()
    sd = O.init_state_dict(cfg, seed=4, perturb=True)
    m = build_model(cfg, sd).eval()
    b = synthetic.make_batch(2, 32, seed=6, vocab=cfg.vocab_size, feat_dim=cfg.n_embd)
    with torch.no_grad():
        lm = m(input_ids=b["input_ids"].cuda(), token_type_ids=b["token_type_ids"].cuda(), labels=b["labels"].cuda())
        em = m(input_ids=b["input_ids"].cuda(), token_type_ids=b["token_type_ids"].cuda(),
               emotion_labels=b["emotion_labels"].cuda())
        no = m(input_ids=b["input_ids"].cuda(), token_type_ids=b["token_type_ids"].cuda())
    o = O.forward(sd, cfg, b["input_ids"], b["token_type_ids"], b["labels"], b["emotion_labels"])
    assert abs(lm.loss.item() - o["lm_loss"].item()) < LOSS_TOL      # model.py:714-718
    assert abs(em.loss.item() - o["emotion_loss"].item()) < LOSS_TOL  # model.py:719-721
    assert no.loss is None
    tup = m(input_ids=b["input_ids"].cuda(), token_type_ids=b["token_type_ids"].cuda(), labels=b["labels"].cuda(),
            return_dict=False)
    assert len(tup) == 4 and tup[1].shape == (2, 32, cfg.vocab_size)  # (loss, logits, emotion_logits, presents)


def test_small_gv1_config1_vs_golden(cuda_device):
    """BASELINE config 1: GPT-2 small, B=4, T=128, caption mode + fusion, against the fixture the
    unmodified reference produced (tests/golden/small_gv1.npz)."""
    g = np.load(os.path.join(GOLD, "small_gv1.npz"))
    cfg = O.OracleConfig()
    sd = O.init_state_dict(cfg, seed=0, perturb=True)
    m = build_model(cfg, sd).eval()
    b = synthetic.gv1_inputs()
    for mode in ("caption", "nocaption"):
        kw = cuda_batch(b, caption=(mode == "caption"))
        with torch.no_grad():
            out = m(**kw)
        # LM loss: mean over 252 tokens -> 1e-3.  The emotion CE is a mean over only B = 4 samples of a
        # 768-long dot product of bf16-rounded activations: its noise floor in bf16-operand mode is ~1e-3 by
        # itself, so the summed loss gets 2.5e-3 here (the fp32 mode test pins it to 1e-4).
        print(mode, "loss", out.loss.item(), float(g[mode + "/loss"]), "lm", out.lm_loss.item(), float(g[mode + "/lm_loss"]))
        assert abs(out.lm_loss.item() - float(g[mode + "/lm_loss"])) < LOSS_TOL
        assert abs(out.loss.item() - float(g[mode + "/loss"])) < 2.5e-3, (mode, out.loss.item())
        lg = out.logits
        assert rel(lg[3, 127], torch.from_numpy(g[mode + "/logits_b3_t127"])) < LOGITS_REL_TOL
        assert rel(lg[0, 0], torch.from_numpy(g[mode + "/logits_b0_t0"])) < LOGITS_REL_TOL
        assert rel(lg[1, ::8, ::64], torch.from_numpy(g[mode + "/logits_b1_stride"])) < LOGITS_REL_TOL
        assert abs(lg.double().norm().item() / float(g[mode + "/logits_norm"]) - 1) < 1e-3
        assert rel(out.emotion_logits, torch.from_numpy(g[mode + "/emotion_logits"])) < LOGITS_REL_TOL
        # greedy next-token agreement where the reference's top-1/top-2 margin is not within bf16 noise
        am = lg.argmax(-1).cpu().numpy()
        safe = g[mode + "/top_margin"] > 0.05
        assert (am[safe] == g[mode + "/argmax"][safe]).mean() > 0.999


def test_dropout_training_statistics(cuda_device):
    """Dropout cannot be bit-matched to torch's RNG stream: check that train-mode losses are
    finite, differ call to call, stay near the eval loss, and that backward runs."""
    cfg = tiny_cfg()
    sd = O.init_state_dict(cfg, seed=3, perturb=True)
    m = build_model(cfg, sd, dropout=0.1)
    b = synthetic.make_batch(4, 64, seed=8, vocab=cfg.vocab_size, feat_dim=cfg.n_embd)
    kw = cuda_batch(b)
    m.eval()
    with torch.no_grad():
        base = m(**kw).loss.item()
    m.train()
    l1 = m(**kw)
    l1.loss.backward()
    g1 = m.transformer.h[0].mlp.c_fc.weight.grad.clone()
    m.zero_grad()
    l2 = m(**kw)
    l2.loss.backward()
    assert l1.loss.item() != l2.loss.item()
    assert abs(l1.loss.item() - base) < 0.5 and abs(l2.loss.item() - base) < 0.5
    assert torch.isfinite(g1).all() and not torch.equal(g1, m.transformer.h[0].mlp.c_fc.weight.grad)


def test_grad_accumulation_and_adamw_step(cuda_device):
    """Two backward passes without zero_grad accumulate; FusedAdamW equals torch.optim.AdamW on the
    same gradients (main.py:68,153-155)."""
    from ergm_b200.optim import FusedAdamW
    cfg = tiny_cfg()
    sd = O.init_state_dict(cfg, seed=3, perturb=True)
    m = build_model(cfg, sd).train()
    b = synthetic.make_batch(2, 32, seed=9, vocab=cfg.vocab_size, feat_dim=cfg.n_embd)
    kw = cuda_batch(b)
    m(**kw).loss.backward()
    g1 = {n: p.grad.clone() for n, p in m.named_parameters()}
    m(**kw).loss.backward()
    for n, p in m.named_parameters():
        assert torch.allclose(p.grad, 2 * g1[n], rtol=1e-3, atol=1e-6), n
    m.zero_grad()
    m(**kw).loss.backward()
    ref_p = {n: p.detach().clone().requires_grad_(True) for n, p in m.named_parameters()}
    for n, p in m.named_parameters():
        ref_p[n].grad = p.grad.clone()
    topt = torch.optim.AdamW(list(ref_p.values()), lr=1e-3)
    topt.step()
    FusedAdamW(m, lr=1e-3).step()
    for n, p in m.named_parameters():
        assert torch.allclose(p.detach(), ref_p[n].detach(), rtol=0, atol=2e-6), n
    # the bf16 shadow the GEMMs read was refreshed by the optimiser kernel
    st = m.engine.store
    assert torch.equal(st.shadow_view("transformer.wte.weight"), m.transformer.wte.weight.detach().bfloat16())


def test_fp32_mode_logits_1e4_and_loss(cuda_device):
    """north_star fp32 mode: logits within 1e-4 relative, loss within 1e-4 (split-operand tcgen05
    GEMMs, fp32 attention), against the oracle in fp64."""
    cfg = tiny_cfg()
    sd = O.init_state_dict(cfg, seed=3, perturb=True)
    m = build_model(cfg, sd).eval()
    m.ergm_precision = "fp32"
    b = synthetic.make_batch(3, 48, seed=11, vocab=cfg.vocab_size, feat_dim=cfg.n_embd, tc=29)
    kw = cuda_batch(b)
    with torch.no_grad():
        out = m(**kw)
    sd64 = {k: v.double() for k, v in sd.items()}
    o = O.forward(sd64, cfg, b["input_ids"], b["token_type_ids"], b["labels"], b["emotion_labels"],
                  b["imgs"].double(), b["auds"].double(), b["caption_ids"])
    r = rel(out.logits, o["logits"])
    print("fp32-mode logits rel err %.3e, loss diff %.3e" % (r, abs(out.loss.item() - o["loss"].item())))
    assert r < 1e-4
    assert abs(out.loss.item() - o["loss"].item()) < 1e-4
    assert rel(out.emotion_logits, o["emotion_logits"]) < 1e-4


def test_fp32_mode_small_gv1_vs_golden(cuda_device):
    """BASELINE config 1 in fp32 mode against the reference fixture: loss within 1e-4, logits 1e-4,
    arg-max map identical wherever the reference's own top-1/top-2 margin exceeds 1e-4."""
    g = np.load(os.path.join(GOLD, "small_gv1.npz"))
    cfg = O.OracleConfig()
    sd = O.init_state_dict(cfg, seed=0, perturb=True)
    m = build_model(cfg, sd).eval()
    m.ergm_precision = "fp32"
    b = synthetic.gv1_inputs()
    with torch.no_grad():
        out = m(**cuda_batch(b))
    assert abs(out.loss.item() - float(g["caption/loss"])) < 1e-4, out.loss.item()
    lg = out.logits
    assert rel(lg[3, 127], torch.from_numpy(g["caption/logits_b3_t127"])) < 1e-4
    assert rel(lg[1, ::8, ::64], torch.from_numpy(g["caption/logits_b1_stride"])) < 1e-4
    am = lg.argmax(-1).cpu().numpy()
    safe = g["caption/top_margin"] > 1e-4
    assert (am[safe] == g["caption/argmax"][safe]).all()
    assert safe.mean() > 0.99


def test_medium_width_long_context_step(cuda_device):
    """BASELINE config 5 shape family: GPT-2-medium width (H=1024, 16 heads), T=512 (four 128-key
    blocks per head: exercises the multi-block causal / cross attention paths), reduced to 2 layers so
    the CPU oracle finishes in seconds.  Forward + backward against the oracle."""
    cfg = O.OracleConfig(vocab_size=2048, n_positions=512, n_embd=1024, n_layer=2, n_head=16)
    sd = O.init_state_dict(cfg, seed=13, perturb=True)
    m = build_model(cfg, sd).train()
    b = synthetic.make_batch(2, 512, seed=14, vocab=cfg.vocab_size, feat_dim=cfg.n_embd, tc=512)
    b["labels"] = b["input_ids"].clone()  # score every position: ~1000 targets keep the mean's bf16 noise < 1e-3
    kw = cuda_batch(b)
    out = m(**kw)
    out.loss.backward()
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "lm_head.weight"}
    sdo["lm_head.weight"] = sdo["transformer.wte.weight"]
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    o = O.forward(sdo, cfg, b["input_ids"], b["token_type_ids"], b["labels"], b["emotion_labels"], b["imgs"],
                  b["auds"], b["caption_ids"])
    o["loss"].backward()
    assert abs(out.lm_loss.item() - o["lm_loss"].item()) < LOSS_TOL
    assert rel(out.logits, o["logits"]) < LOGITS_REL_TOL
    for name in ("transformer.h.0.attn.c_attn.weight", "transformer.h.1.mlp.c_fc.weight", "transformer.wpe.weight",
                 "transformer.h.0.ln_1.weight", "transformer.h.1.attn.c_proj.bias"):
        p = dict(m.named_parameters())[name]
        assert rel(p.grad, sdo[name].grad) < 3e-2, (name, rel(p.grad, sdo[name].grad))


def test_modality_sequence_projection_extension(cuda_device):
    """A3 extension (768-wide features cannot feed a 1024-wide backbone in the
    reference): raw audio [B,113,768] / 4-key-frame visual [B,788,768] sequences are mean-pooled and
    projected 768 -> 1024 on the device.  Forward, loss and the projection gradients against the
    oracle's restatement of feature_extraction.py:63,69 + Linear (parity unpinned by the reference:
    the extension has no reference implementation)."""
    cfg = O.OracleConfig(vocab_size=2048, n_positions=128, n_embd=1024, n_layer=2, n_head=16, visual_dim=768,
                         audio_dim=768)
    sd = O.init_state_dict(cfg, seed=21, perturb=True)
    m = build_model(cfg, sd).train()
    b = synthetic.make_batch(4, 64, seed=22, vocab=cfg.vocab_size, feat_dim=768, kf=4)
    b["labels"] = b["input_ids"].clone()
    kw = cuda_batch(b, fusion=False)
    kw["imgs"], kw["auds"] = b["vis_seq"].cuda(), b["aud_seq"].cuda()
    out = m(**kw)
    out.loss.backward()
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "lm_head.weight"}
    sdo["lm_head.weight"] = sdo["transformer.wte.weight"]
    o = O.forward(sdo, cfg, b["input_ids"], b["token_type_ids"], b["labels"], b["emotion_labels"], b["vis_seq"],
                  b["aud_seq"], b["caption_ids"])
    o["loss"].backward()
    assert abs(out.lm_loss.item() - o["lm_loss"].item()) < LOSS_TOL
    assert rel(out.logits, o["logits"]) < LOGITS_REL_TOL
    # the fused rows (positions 0 and 1) must carry the projected features: compare logits there alone
    assert rel(out.logits[:, :2], o["logits"][:, :2]) < LOGITS_REL_TOL
    for name in ("visual_proj.weight", "visual_proj.bias", "audio_proj.weight", "audio_proj.bias"):
        p = dict(m.named_parameters())[name]
        assert p.grad is not None and rel(p.grad, sdo[name].grad) < 4e-2, (name, rel(p.grad, sdo[name].grad))
    # without the projection parameters the reference layout is untouched
    assert "visual_proj.weight" not in build_model(O.OracleConfig(vocab_size=256, n_positions=64, n_embd=128, n_layer=1,
                                                                   n_head=2),
                                                   O.init_state_dict(O.OracleConfig(vocab_size=256, n_positions=64,
                                                                                    n_embd=128, n_layer=1, n_head=2),
                                                                     seed=1)).state_dict()


@pytest.mark.parametrize("B,T,Tc", [(1, 17, 5), (3, 33, 1), (2, 130, 257), (5, 128, 129), (1, 1, 3)])
def test_odd_shapes_forward_backward(cuda_device, B, T, Tc):
    """Edge shapes: sequence / caption lengths that are not multiples of any tile (17, 33, 130 crosses the
    128-row attention block, caption length 257 = two full key blocks + 1, a single token), batch 1.
    Forward, loss and a few gradients against the oracle."""
    cfg = O.OracleConfig(vocab_size=515, n_positions=300, n_embd=128, n_layer=2, n_head=2)
    sd = O.init_state_dict(cfg, seed=17, perturb=True)
    m = build_model(cfg, sd).train()
    g = torch.Generator().manual_seed(B * 100 + T)
    ids = torch.randint(0, cfg.vocab_size, (B, T), generator=g)
    tt = torch.randint(cfg.vocab_size - 2, cfg.vocab_size, (B, T), generator=g)
    lab = ids.clone()
    if T > 2:
        lab[:, : T // 3] = -100
    emo = torch.randint(0, 7, (B,), generator=g)
    cap = torch.randint(0, cfg.vocab_size, (B, Tc), generator=g)
    imgs, auds = torch.randn(B, 1, 128, generator=g), torch.randn(B, 128, generator=g)
    if T < 2:
        imgs = auds = None  # model.py:497-498 indexes positions 0 and 1
    kw = dict(input_ids=ids.cuda(), token_type_ids=tt.cuda(), labels=lab.cuda(), emotion_labels=emo.cuda(),
              caption_ids=cap.cuda())
    if imgs is not None:
        kw.update(imgs=imgs.cuda(), auds=auds.cuda())
    out = m(**kw)
    out.loss.backward()
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "lm_head.weight"}
    sdo["lm_head.weight"] = sdo["transformer.wte.weight"]
    o = O.forward(sdo, cfg, ids, tt, lab, emo, imgs, auds, cap)
    o["loss"].backward()
    assert rel(out.logits, o["logits"]) < LOGITS_REL_TOL
    if T > 1:  # T == 1: the shifted LM loss has no target (mean over zero labels is nan in the reference too)
        assert abs(out.loss.item() - o["loss"].item()) < 5e-3
    assert rel(out.emotion_logits, o["emotion_logits"]) < 2e-2
    if T > 1:
        for name in ("transformer.h.0.attn.c_attn.weight", "transformer.h.1.crossattention.c_attn.weight",
                     "transformer.wpe.weight", "transformer.h.1.ln_2.bias"):
            p = dict(m.named_parameters())[name]
            assert rel(p.grad, sdo[name].grad) < 6e-2, (name, rel(p.grad, sdo[name].grad))


def test_optimizer_checkpoint_is_torch_adamw_compatible(cuda_device):
    """N4: the reference checkpoints torch.optim.AdamW.state_dict() (main.py:103-110,184-196).
    (a) a torch AdamW state dict loads into FusedAdamW with identical moments / step / hyper-parameters;
    (b) save -> fresh model + fresh FusedAdamW -> load resumes on the same trajectory as the uninterrupted run;
    (c) FusedAdamW.state_dict() loads back into torch's AdamW.
    (A torch step and a fused step agree to 1e-7 in fp32, which is enough to flip single bf16 roundings of the
    weight shadow, so trajectories are compared fused-vs-fused; the moments are compared exactly.)"""
    from ergm_b200.optim import FusedAdamW
    cfg = O.OracleConfig(vocab_size=512, n_positions=64, n_embd=128, n_layer=2, n_head=2)
    sd = O.init_state_dict(cfg, seed=9, perturb=True)
    b = synthetic.make_batch(3, 32, seed=10, vocab=cfg.vocab_size, feat_dim=cfg.n_embd)
    kw = cuda_batch(b)

    def grad_step(m):
        for p in m.parameters():
            p.grad = None
        m(**kw).loss.backward()

    # (a) torch -> fused
    ref = build_model(cfg, sd).train()
    opt_ref = FusedAdamW(ref, lr=1e-3)
    m = build_model(cfg, sd).train()
    topt = torch.optim.AdamW(m.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    grad_step(ref)
    grad_step(m)
    opt_ref.step()
    topt.step()
    fopt = FusedAdamW(m, lr=123.0)           # wrong on purpose: everything must come from the checkpoint
    fopt.load_state_dict(topt.state_dict())
    assert fopt.step_count == 1 and abs(fopt.lr - 1e-3) < 1e-12 and fopt.weight_decay == 0.01
    assert (fopt.state["m"] - opt_ref.state["m"]).abs().max().item() < 1e-6
    assert (fopt.state["v"] - opt_ref.state["v"]).abs().max().item() < 1e-6
    assert max((p.detach() - q.detach()).abs().max().item() for p, q in zip(m.parameters(), ref.parameters())) < 1e-6
    # (b) fused -> checkpoint -> fresh model / optimiser -> same trajectory
    ckpt = {"model_state_dict": {k: v.detach().cpu().clone() for k, v in ref.state_dict().items()},
            "optim_state_dict": opt_ref.state_dict()}
    resumed = build_model(cfg, ckpt["model_state_dict"]).train()
    ropt = FusedAdamW(resumed, lr=7.0)
    ropt.load_state_dict(ckpt["optim_state_dict"])
    for _ in range(2):
        grad_step(ref)
        opt_ref.step()
        grad_step(resumed)
        ropt.step()
    worst = max((p.detach() - q.detach()).abs().max().item() for p, q in zip(resumed.parameters(), ref.parameters()))
    assert worst < 1e-5, worst
    # (c) fused -> torch
    t2 = torch.optim.AdamW(ref.parameters(), lr=5.0)
    t2.load_state_dict(opt_ref.state_dict())
    assert abs(t2.param_groups[0]["lr"] - 1e-3) < 1e-12
    st0 = t2.state[t2.param_groups[0]["params"][0]]
    assert int(st0["step"]) == 3 and st0["exp_avg"].shape == next(ref.parameters()).shape

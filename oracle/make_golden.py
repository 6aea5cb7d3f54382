"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/ref_shim.py) —
run in the build container only:   python -m oracle.make_golden

For every fixture it first asserts that the oracle restatement (oracle/ergm_oracle.py) is
bit-identical to the reference on the same inputs, then stores the reference's outputs (or
slices / checksums of them when the tensor is large).  Weights come from
ergm_oracle.init_state_dict (deterministic per-tensor generators), inputs from
oracle/synthetic.py, so the GPU box can regenerate both without the reference tree.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ergm_oracle as O  # noqa: E402
from oracle import ref_shim, synthetic  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def tiny_cfg():
    return O.OracleConfig(vocab_size=1024, n_positions=256, n_embd=128, n_layer=2, n_head=2)


def run_reference(ref, batch, caption=True, fusion=True, labels=True):
    kw = dict(input_ids=batch["input_ids"], token_type_ids=batch["token_type_ids"])
    if labels:
        kw.update(labels=batch["labels"], emotion_labels=batch["emotion_labels"])
    if caption:
        kw.update(caption_ids=batch["caption_ids"])
    if fusion:
        kw.update(imgs=batch["imgs"], auds=batch["auds"])
    return ref(**kw)


def fixture_tiny():
    """Tiny model, full tensors: logits, losses, emotion logits, every parameter gradient."""
    cfg = tiny_cfg()
    sd = O.init_state_dict(cfg, seed=3, perturb=True)
    batch = synthetic.make_batch(3, 48, seed=11, vocab=cfg.vocab_size, feat_dim=cfg.n_embd, tc=40)
    out = {}
    for mode, caption in (("caption", True), ("nocaption", False)):
        # the reference forces Tc == T (model.py:461): give it T-long captions in caption mode
        b = dict(batch)
        if caption:
            b["caption_ids"] = synthetic.make_batch(3, 48, seed=12, vocab=cfg.vocab_size, feat_dim=cfg.n_embd)["caption_ids"]
        ref = ref_shim.build_reference_model(cfg, sd, no_caption_guard=not caption)
        ref.zero_grad()
        r = run_reference(ref, b, caption=caption)
        r.loss.backward()
        sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "lm_head.weight"}
        sdo["lm_head.weight"] = sdo["transformer.wte.weight"]
        o = O.forward(sdo, cfg, b["input_ids"], b["token_type_ids"], b["labels"], b["emotion_labels"],
                      b["imgs"], b["auds"], b["caption_ids"] if caption else None)
        o["loss"].backward()
        assert torch.equal(o["logits"], r.logits), "oracle != reference (logits, %s)" % mode
        assert torch.equal(o["loss"], r.loss)
        assert torch.equal(o["emotion_logits"], r.emotion_logits)
        refp = dict(ref.named_parameters())
        for k, v in sdo.items():
            if k == "lm_head.weight":
                continue
            g = refp[k].grad
            if g is None:
                assert v.grad is None or float(v.grad.abs().max()) == 0.0, k
                continue
            assert torch.allclose(v.grad, g, rtol=0, atol=1e-7), ("grad mismatch", k, float((v.grad - g).abs().max()))
            out["%s/grad/%s" % (mode, k)] = _np(g)
        out["%s/logits" % mode] = _np(r.logits)
        out["%s/loss" % mode] = _np(r.loss)
        out["%s/emotion_logits" % mode] = _np(r.emotion_logits)
        out["%s/caption_ids" % mode] = _np(b["caption_ids"])
        # KV-cache surface: present k/v of layer 0 and last layer
        out["%s/present0_k" % mode] = _np(r.past_key_values[0][0])
        out["%s/presentL_v" % mode] = _np(r.past_key_values[-1][1])
    np.savez_compressed(os.path.join(OUT, "tiny.npz"), **out)
    print("tiny.npz: %d arrays" % len(out))


def fixture_tiny_generate():
    """Greedy decoding: reference full-recompute loop (main.py:255-257 with argmax) and the
    reference KV-cache path, no-caption mode, larger init range so ids do not collapse."""
    cfg = tiny_cfg()
    cfg.initializer_range = 0.2
    sd = O.init_state_dict(cfg, seed=5, perturb=True)
    batch = synthetic.make_batch(4, 24, seed=21, vocab=cfg.vocab_size, feat_dim=cfg.n_embd, ragged=False)
    ref = ref_shim.build_reference_model(cfg, sd, no_caption_guard=True)
    ids, tt = batch["input_ids"].clone(), batch["token_type_ids"].clone()
    new = []
    with torch.no_grad():
        for _ in range(12):
            lg = ref(input_ids=ids, token_type_ids=tt).logits[:, -1, :]
            nxt = lg.argmax(-1)
            new.append(nxt)
            ids = torch.cat([ids, nxt[:, None]], 1)
            tt = torch.cat([tt, torch.full((4, 1), cfg.vocab_size - 1)], 1)
        ref_ids = torch.stack(new, 1)
        # reference KV-cache path
        r = ref(input_ids=batch["input_ids"], token_type_ids=batch["token_type_ids"], use_cache=True)
        past, lg = r.past_key_values, r.logits[:, -1, :]
        new2 = []
        for _ in range(12):
            nxt = lg.argmax(-1)
            new2.append(nxt)
            r = ref(input_ids=nxt[:, None], token_type_ids=torch.full((4, 1), cfg.vocab_size - 1),
                    past_key_values=past, use_cache=True)
            past, lg = r.past_key_values, r.logits[:, -1, :]
        cached_ids = torch.stack(new2, 1)
        o_ids = O.greedy_generate_recompute(sd, cfg, batch["input_ids"], batch["token_type_ids"], 12,
                                            sp2_id=cfg.vocab_size - 1, eos_id=-1)
        o_ids2 = O.greedy_generate_cached(sd, cfg, batch["input_ids"], batch["token_type_ids"], 12,
                                          sp2_id=cfg.vocab_size - 1, eos_id=-1)
    assert torch.equal(ref_ids, cached_ids), "reference cached != recompute"
    assert torch.equal(ref_ids, o_ids) and torch.equal(ref_ids, o_ids2), "oracle generation != reference"
    assert len(set(ref_ids.flatten().tolist())) > 8, "degenerate greedy sequence"
    np.savez_compressed(os.path.join(OUT, "tiny_generate.npz"), greedy_ids=_np(ref_ids))
    print("tiny_generate.npz", ref_ids[0].tolist())


def fixture_small_gv1():
    """GPT-2 small (152.8 M params) on the GV-1 inputs of SURVEY.md §8(c): BASELINE config 1.
    Stores losses, emotion logits, logits slices and checksums (logits are 103 MB)."""
    cfg = O.OracleConfig()
    sd = O.init_state_dict(cfg, seed=0, perturb=True)
    b = synthetic.gv1_inputs()
    out = {}
    with torch.no_grad():
        for mode, caption in (("caption", True), ("nocaption", False)):
            ref = ref_shim.build_reference_model(cfg, sd, no_caption_guard=not caption)
            r = run_reference(ref, b, caption=caption)
            r_lm = run_reference(ref, dict(b, emotion_labels=None), caption=caption) if False else None
            o = O.forward(sd, cfg, b["input_ids"], b["token_type_ids"], b["labels"], b["emotion_labels"],
                          b["imgs"], b["auds"], b["caption_ids"] if caption else None)
            assert torch.equal(o["logits"], r.logits), "oracle != reference (%s)" % mode
            assert torch.equal(o["loss"], r.loss)
            lg = r.logits
            out["%s/loss" % mode] = _np(r.loss)
            out["%s/lm_loss" % mode] = _np(o["lm_loss"])
            out["%s/emotion_logits" % mode] = _np(r.emotion_logits)
            out["%s/logits_b0_t0" % mode] = _np(lg[0, 0])
            out["%s/logits_b3_t127" % mode] = _np(lg[3, 127])
            out["%s/logits_b1_stride" % mode] = _np(lg[1, ::8, ::64])
            out["%s/logits_sum" % mode] = np.float64(lg.double().sum().item())
            out["%s/logits_abs_sum" % mode] = np.float64(lg.double().abs().sum().item())
            out["%s/logits_norm" % mode] = np.float64(lg.double().norm().item())
            out["%s/argmax" % mode] = _np(lg.argmax(-1))
            top2 = lg.topk(2, -1).values
            out["%s/top_margin" % mode] = _np(top2[..., 0] - top2[..., 1])
            out["%s/hidden_b2" % mode] = _np(o["hidden"][2])
            print(mode, "loss", float(r.loss), "lm", float(o["lm_loss"]))
    np.savez_compressed(os.path.join(OUT, "small_gv1.npz"), **out)


def check_survey_gv1():
    """Re-derives the GV-1 known answers quoted in SURVEY.md §8(c) with the reference's own
    init (torch.manual_seed(0)); informational — depends on the image's init RNG order."""
    from transformers import GPT2Config
    mod = ref_shim.load()
    torch.manual_seed(0)
    m = mod.GPT2LMHeadModel(GPT2Config(vocab_size=50260)).eval()
    b = synthetic.gv1_inputs()
    with torch.no_grad():
        r = run_reference(m, b)
    print("GV-1 reference loss %.9f (SURVEY: 12.970413208)" % float(r.loss))
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    with torch.no_grad():
        o = O.forward(sd, O.OracleConfig(), b["input_ids"], b["token_type_ids"], b["labels"], b["emotion_labels"],
                      b["imgs"], b["auds"], b["caption_ids"])
    assert torch.equal(o["logits"], r.logits)
    np.savez_compressed(os.path.join(OUT, "survey_gv1.npz"), loss=_np(r.loss), lm_loss=_np(o["lm_loss"]),
                        logits_0_0_4=_np(r.logits[0, 0, :4]), emotion_logits_0=_np(r.emotion_logits[0]))


if __name__ == "__main__":
    assert ref_shim.available(), "reference tree not present"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    fixture_tiny()
    fixture_tiny_generate()
    fixture_small_gv1()
    check_survey_gv1()

"""Packed variable-length batches (the reference's collate, custom_dataset.py:102-132, right-pads
every sample and model.py computes every pad position).

Oracle: the reference arithmetic called WITH the right-padded attention_mask (oracle/ergm_oracle.py's attention_mask
path = model.py:478-482): the packed path must reproduce it at every real position - logits, LM loss, every gradient -
and in the emotion head, which reads position T-1 (model.py:700): a padded sample keeps ONE extra row for that
position.  (main.py passes no mask; its pad positions attend to earlier pads, which only changes what the emotion
head sees - that is why packing is opt-in, `model.ergm_packed`.)
"""
import pytest
import torch

from oracle import ergm_oracle as O
from ergm_b200 import synthetic

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _build(cfg, sd, packed=True):
    from transformers import GPT2Config
    from ergm_b200.model import GPT2LMHeadModel
    hf = GPT2Config(vocab_size=cfg.vocab_size, n_positions=cfg.n_positions, n_embd=cfg.n_embd, n_layer=cfg.n_layer,
                    n_head=cfg.n_head, attn_pdrop=0.0, resid_pdrop=0.0, embd_pdrop=0.0)
    m = GPT2LMHeadModel(hf)
    m.load_state_dict(sd)
    m = m.to("cuda").train()
    m.ergm_packed = packed
    return m


def test_pack_plan(cuda_device):
    from ergm_b200 import ops
    B, T = 5, 40
    lens = torch.tensor([40, 1, 17, 39, 23], dtype=torch.int32, device="cuda")
    pk = ops.Pack(B, T, torch.device("cuda")).plan(lens)
    n_b = [40, 2, 18, 40, 24]   # +1 row (position T-1) for every padded sample
    cu = [0]
    for n in n_b:
        cu.append(cu[-1] + n)
    assert pk.cu.cpu().tolist() == cu and int(pk.n_rows.item()) == cu[-1]
    assert pk.kv_lens.cpu().tolist() == [40, 1, 17, 39, 23]
    rb, rt = pk.row_b.cpu().tolist(), pk.row_t.cpu().tolist()
    for b in range(B):
        L = int(lens[b])
        want_t = list(range(L)) + ([T - 1] if L < T else [])
        assert rt[cu[b]:cu[b + 1]] == want_t and set(rb[cu[b]:cu[b + 1]]) == {b}


@pytest.mark.parametrize("causal", [True, False])
def test_attention_packed_equals_padded(cuda_device, causal):
    """attn_fwd / attn_bwd on packed rows == the padded call with kv_lens, at every real row (and the extra row)."""
    from ergm_b200 import ops
    B, nh, T, H = 3, 2, 200, 128
    Tk = T if causal else 150
    lens = torch.tensor([200, 77, 130], dtype=torch.int32, device="cuda")
    pk = ops.Pack(B, T, torch.device("cuda")).plan(lens)
    n = int(pk.n_rows.item())
    g = torch.Generator(device="cuda").manual_seed(1)
    qkv_pad = torch.randn(B * T, 3 * H, device="cuda", generator=g).bfloat16()
    dout_pad = torch.randn(B * T, H, device="cuda", generator=g).bfloat16()
    src = (pk.row_b[:n].long() * T + pk.row_t[:n].long())
    qkv_pk = torch.zeros(B * T, 3 * H, device="cuda", dtype=torch.bfloat16)
    dout_pk = torch.zeros(B * T, H, device="cuda", dtype=torch.bfloat16)
    qkv_pk[:n] = qkv_pad[src]
    dout_pk[:n] = dout_pad[src]
    kv = torch.randn(B * Tk, 2 * H, device="cuda", generator=g).bfloat16()   # cross-attention keys / values (not packed)

    def run(qkv, dout, pack):
        out = torch.zeros(B * T, H, device="cuda", dtype=torch.bfloat16)
        o32 = torch.zeros(B * T, H, device="cuda")
        lse = torch.zeros(B, nh, T, device="cuda")
        delta = torch.zeros(B, nh, T, device="cuda")
        dq = torch.zeros(B * T, H, device="cuda", dtype=torch.bfloat16)
        if causal:
            dkv = torch.zeros(B * T, 3 * H, device="cuda", dtype=torch.bfloat16)
            kw = dict(B=B, nh=nh, Tq=T, Tk=T, q_col0=0, k_col0=H, v_col0=2 * H, causal=True,
                      kv_lens=None if pack is not None else lens, pack=pack, pack_kv=pack is not None)
            ops.attn_fwd(qkv, qkv, qkv, out, lse, out_f32=o32, **kw)
            ops.attn_bwd(qkv, qkv, qkv, out, dout, lse, delta, dq, dkv, dkv, dk_col0=H, dv_col0=2 * H, out_f32=o32, **kw)
        else:
            dkv = torch.zeros(B * Tk, 2 * H, device="cuda", dtype=torch.bfloat16)
            kw = dict(B=B, nh=nh, Tq=T, Tk=Tk, q_col0=0, k_col0=0, v_col0=H, causal=False, pack=pack)
            ops.attn_fwd(qkv, kv, kv, out, lse, out_f32=o32, **kw)
            ops.attn_bwd(qkv, kv, kv, out, dout, lse, delta, dq, dkv, dkv, dk_col0=0, dv_col0=H, out_f32=o32, **kw)
        return out, dq, dkv

    o_pad, dq_pad, dkv_pad = run(qkv_pad, dout_pad, None)
    o_pk, dq_pk, dkv_pk = run(qkv_pk, dout_pk, pk)
    if causal:
        # the padded call computes pad queries too (they see real keys only): position T-1 is the packed extra row
        assert rel(o_pk[:n], o_pad[src]) < 1e-6
        assert rel(dq_pk[:n], dq_pad[src]) < 2e-2
        # dK / dV of the real keys: the padded call also receives contributions of ITS pad queries (rows the packed
        # layout does not have), so compare through a padded call whose pad-row dO is zero
        real = (pk.row_t[:n].long() < lens[pk.row_b[:n].long()].long())
        dmask = torch.zeros(B * T, 1, device="cuda", dtype=torch.bfloat16)
        dmask[src] = 1
        _, dq2, dkv2 = run(qkv_pad, dout_pad * dmask, None)
        assert rel(dkv_pk[:n][real][:, H:], dkv2[src][real][:, H:]) < 2e-2
        assert rel(dq_pk[:n], dq2[src]) < 2e-2
    else:
        assert rel(o_pk[:n], o_pad[src]) < 1e-6
        dmask = torch.zeros(B * T, 1, device="cuda", dtype=torch.bfloat16)
        dmask[src] = 1
        _, dq2, dkv2 = run(qkv_pad, dout_pad * dmask, None)
        assert rel(dq_pk[:n], dq2[src]) < 2e-2
        assert rel(dkv_pk, dkv2) < 2e-2


@pytest.mark.parametrize("caption", [True, False])
def test_packed_step_equals_masked_oracle(cuda_device, caption):
    cfg = O.OracleConfig(vocab_size=1024, n_positions=256, n_embd=128, n_layer=2, n_head=2)
    sd = O.init_state_dict(cfg, seed=61, perturb=True)
    B, T = 6, 160   # sequences cross the 128-row attention block
    b = synthetic.make_batch(B, T, seed=62, vocab=1024, feat_dim=128, tc=70)
    lens = (b["token_type_ids"] != 1020).sum(1)             # pad type id = eos = vocab - 4
    b["input_ids"][0, :] = b["input_ids"][0, 0]             # one sample without padding: every position real
    b["token_type_ids"][0, :] = 1023
    lens[0] = T
    mask = (torch.arange(T)[None] < lens[:, None]).long()
    m = _build(cfg, sd)
    kw = dict(input_ids=b["input_ids"].cuda(), token_type_ids=b["token_type_ids"].cuda(), labels=b["labels"].cuda(),
              emotion_labels=b["emotion_labels"].cuda(), imgs=b["imgs"].cuda(), auds=b["auds"].cuda(),
              attention_mask=mask.cuda())
    if caption:
        kw["caption_ids"] = b["caption_ids"].cuda()
    out = m(**kw)
    n_rows = int(m.engine.saved["pack"].n_rows.item())
    assert n_rows == int(lens.sum()) + int((lens < T).sum()) and n_rows < B * T
    logits = out.logits
    out.loss.backward()
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "lm_head.weight"}
    sdo["lm_head.weight"] = sdo["transformer.wte.weight"]
    o = O.forward(sdo, cfg, b["input_ids"], b["token_type_ids"], b["labels"], b["emotion_labels"], b["imgs"], b["auds"],
                  b["caption_ids"] if caption else None, attention_mask=mask)
    o["loss"].backward()
    assert abs(out.lm_loss.item() - o["lm_loss"].item()) < 1e-3
    assert abs(out.loss.item() - o["loss"].item()) < 2.5e-3
    assert rel(out.emotion_logits, o["emotion_logits"]) < 1e-2      # the emotion head's position T-1 (pad for 5 of 6 samples)
    real = mask.bool()
    assert rel(logits.cpu()[real], o["logits"][real]) < 1e-2
    params = dict(m.named_parameters())
    worst = ("", 0.0)
    for n, p in params.items():
        if p.grad is None:
            assert not caption and ("crossattention" in n or "ln_cross_attn" in n), n
            continue
        r = rel(p.grad, sdo[n].grad)
        worst = max(worst, (n, r), key=lambda t: t[1])
        tol = 6e-2 if ("crossattention" in n or "ln_cross_attn" in n) else 3e-2
        assert r < tol, (n, r)
    print("packed step (%s): %d of %d rows computed, worst gradient rel err %.2e (%s)"
          % ("caption" if caption else "no caption", n_rows, B * T, worst[1], worst[0]))


def test_packed_graphed_train_step(cuda_device):
    """GraphedTrainStep with `seq_lens` in the batch: the packed path inside one CUDA graph; the SAME graph serves batches
    with different lengths (run-time row count on the device) and follows the eager packed model step by step."""
    from ergm_b200.optim import FusedAdamW
    from ergm_b200.trainer import GraphedTrainStep
    cfg = O.OracleConfig(vocab_size=1024, n_positions=128, n_embd=128, n_layer=2, n_head=2)
    sd = O.init_state_dict(cfg, seed=63, perturb=True)
    mg, me = _build(cfg, sd), _build(cfg, sd)
    step = GraphedTrainStep(mg, FusedAdamW(mg, lr=1e-3))
    eopt = FusedAdamW(me, lr=1e-3)
    for it in range(4):
        b = synthetic.make_batch(4, 96, seed=70 + it, vocab=1024, feat_dim=128)
        lens = (b["token_type_ids"] != 1020).sum(1).to(torch.int32)
        host = {k: b[k] for k in ("input_ids", "token_type_ids", "labels", "emotion_labels", "caption_ids", "auds")}
        host["imgs"] = b["imgs"][:, 0].contiguous()
        host["seq_lens"] = lens
        host = {k: v.pin_memory() for k, v in host.items()}
        lg = step(host)
        mask = (torch.arange(96)[None] < lens[:, None]).long()
        eopt.zero_grad()
        out = me(input_ids=b["input_ids"].cuda(), token_type_ids=b["token_type_ids"].cuda(), labels=b["labels"].cuda(),
                 emotion_labels=b["emotion_labels"].cuda(), caption_ids=b["caption_ids"].cuda(), imgs=b["imgs"].cuda(),
                 auds=b["auds"].cuda(), attention_mask=mask.cuda())
        out.loss.backward()
        eopt.step()
        assert abs(lg - out.loss.item()) < 2e-3, (it, lg, out.loss.item())
    assert len(step.graphs) == 1     # one captured graph served four different row counts
    w_g = mg.transformer.h[1].mlp.c_fc.weight.detach()
    w_e = me.transformer.h[1].mlp.c_fc.weight.detach()
    assert (w_g - w_e).abs().max().item() < 5e-3

"""The drop-in boundary, driven the way ERGM's own driver drives it.

`compat/model.py` is imported under the module name `model`, exactly what `from model import *` in
/root/reference/src/main.py:22 resolves to, and the test replays main.py's call sequence against it:
construct -> resize_token_embeddings (main.py:62-64) -> torch.optim.AdamW + polynomial schedule (:68, :93-95)
-> the training step with its CE recompute from outputs.logits (:137-169) -> the one-forward-per-token nucleus
sampling loop (:253-282).  The same sequence runs on the oracle (fp32, CPU) and the two are compared.
Also: GPT2Model / GPT2Block / GPT2Attention / GPT2MLP forwards against the oracle's block functions.
"""
import importlib
import os
import sys

import pytest
import torch
import torch.nn.functional as F
from torch import nn

from oracle import ergm_oracle as O
from ergm_b200 import synthetic

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def import_as_model():
    """`import model` with compat/ first on the path = what main.py:22 would get."""
    compat = os.path.join(ROOT, "compat")
    sys.modules.pop("model", None)
    sys.path.insert(0, compat)
    try:
        return importlib.import_module("model")
    finally:
        sys.path.remove(compat)


def _oracle_state(m):
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    sd["lm_head.weight"] = sd["transformer.wte.weight"]
    return sd


def test_main_py_training_and_sampling_sequence(cuda_device):
    model = import_as_model()
    for name in ("GPT2LMHeadModel", "GPT2Model", "CausalLMOutputWithEmotionClassification", "torch", "nn", "F", "math", "os"):
        assert hasattr(model, name), name   # what `import *` leaks today
    from transformers import GPT2Config, get_polynomial_decay_schedule_with_warmup
    base_vocab, vocab = 1021, 1024          # tokenizer.add_special_tokens adds <bos>, <sp1>, <sp2> (main.py:47-54)
    eos_id, bos_id, sp1_id, sp2_id = 1020, 1021, 1022, 1023
    torch.manual_seed(0)
    hf = GPT2Config(vocab_size=base_vocab, n_positions=128, n_embd=128, n_layer=2, n_head=2, attn_pdrop=0.0,
                    resid_pdrop=0.0, embd_pdrop=0.0)
    # main.py:62-64 (from_pretrained needs the hub: constructed from the config instead)
    m = model.GPT2LMHeadModel(hf).to("cuda")
    m.resize_token_embeddings(vocab)
    assert m.transformer.wte.weight.shape[0] == vocab and m.lm_head.weight is m.transformer.wte.weight
    max_len = min(64, m.config.n_ctx)
    assert max_len == 64
    sd0 = _oracle_state(m)
    cfg = O.OracleConfig(vocab_size=vocab, n_positions=128, n_embd=128, n_layer=2, n_head=2)
    # oracle twin: same weights, same optimiser class, same schedule, on the CPU in fp32
    osd = {k: v.clone().requires_grad_(True) for k, v in sd0.items() if k != "lm_head.weight"}
    osd["lm_head.weight"] = osd["transformer.wte.weight"]
    lr, total = 1e-3, 10
    optim = torch.optim.AdamW(m.parameters(), lr=lr)                                   # main.py:68
    sched = get_polynomial_decay_schedule_with_warmup(optim, num_warmup_steps=2, num_training_steps=total, power=2)  # :93-95
    o_optim = torch.optim.AdamW([v for k, v in osd.items() if k != "lm_head.weight"], lr=lr)
    o_sched = get_polynomial_decay_schedule_with_warmup(o_optim, num_warmup_steps=2, num_training_steps=total, power=2)
    m.train()
    for it in range(4):
        b = synthetic.make_batch(4, 48, seed=40 + it, vocab=vocab, feat_dim=128)
        input_ids, token_type_ids, lm_labels = b["input_ids"].cuda(), b["token_type_ids"].cuda(), b["labels"].cuda()
        emotion_labels = torch.LongTensor(b["emotion_labels"].tolist()).cuda()
        # ---- main.py:147-156 ----
        outputs = m(input_ids=input_ids, token_type_ids=token_type_ids, labels=lm_labels, emotion_labels=emotion_labels)
        loss = outputs.loss
        optim.zero_grad()
        loss.backward()
        optim.step()
        sched.step()
        loss_item = loss.item()
        # ---- main.py:160-169 ----
        with torch.no_grad():
            shift_logits = outputs.logits[..., :-1, :].contiguous()
            shift_labels = lm_labels[..., 1:].contiguous()
            lm_loss = nn.CrossEntropyLoss()(shift_logits.view(-1, shift_logits.size(-1)), shift_labels.view(-1)).item()
            preds = torch.argmax(outputs.emotion_logits, dim=-1)
        # ---- oracle twin ----
        o = O.forward(osd, cfg, b["input_ids"], b["token_type_ids"], b["labels"], b["emotion_labels"])
        o_optim.zero_grad()
        o["loss"].backward()
        o_optim.step()
        o_sched.step()
        assert abs(loss_item - o["loss"].item()) < 2.5e-3, (it, loss_item, o["loss"].item())
        assert abs(lm_loss - o["lm_loss"].item()) < 1e-3, (it, lm_loss, o["lm_loss"].item())
        assert abs(lm_loss - outputs.lm_loss.item()) < 1e-3   # the recompute equals the fused kernel's own number (N4)
        assert preds.shape == (4,)
        assert abs(optim.param_groups[0]["lr"] - o_optim.param_groups[0]["lr"]) < 1e-12
    # after 4 AdamW steps (lr 1e-3: every element moves ~1e-3 per step, sign-driven) the weights still agree
    new = _oracle_state(m)
    moved = rel(new["transformer.h.1.mlp.c_fc.weight"], sd0["transformer.h.1.mlp.c_fc.weight"])
    drift = rel(new["transformer.h.1.mlp.c_fc.weight"], osd["transformer.h.1.mlp.c_fc.weight"].detach())
    print("4 torch.optim.AdamW steps: weights moved %.3e, differ from the oracle twin by %.3e" % (moved, drift))
    assert moved > 5 * drift and drift < 2e-2
    # cross-attention parameters took no part (no caption_ids): torch leaves them untouched
    assert m.transformer.h[0].crossattention.c_attn.weight.grad is None
    assert torch.equal(new["transformer.h.0.crossattention.c_attn.weight"], sd0["transformer.h.0.crossattention.c_attn.weight"])

    # ---- main.py:253-282: nucleus_sampling, one full forward per new token, batch 1 ----
    m.eval()
    top_p = 0.8
    sd1 = {k: v.detach() for k, v in osd.items()}

    def nucleus_filter(next_token_logits):  # main.py:258-269 verbatim in structure
        probs = F.softmax(next_token_logits, dim=-1)
        sorted_probs, sorted_idxs = torch.sort(probs, descending=True)
        cumsum_probs = torch.cumsum(sorted_probs, dim=-1)
        idx_remove = cumsum_probs > top_p
        idx_remove[:, 1:] = idx_remove[:, :-1].clone()
        idx_remove[:, 0] = False
        sorted_probs[idx_remove] = 0.0
        sorted_probs /= torch.sum(sorted_probs, dim=-1, keepdim=True)
        return torch.zeros(probs.shape, device=probs.device).scatter_(-1, sorted_idxs, sorted_probs)

    b = synthetic.make_batch(1, 24, seed=77, vocab=vocab, feat_dim=128, ragged=False)
    input_ids, token_type_ids = b["input_ids"].cuda(), b["token_type_ids"].cuda()
    o_ids, o_tt = b["input_ids"].clone(), b["token_type_ids"].clone()
    input_len = input_ids.shape[1]
    with torch.no_grad():
        for pos in range(input_len, input_len + 8):
            outputs = m(input_ids=input_ids, token_type_ids=token_type_ids)
            probs = nucleus_filter(outputs.logits[:, pos - 1, :])
            o_probs = nucleus_filter(O.forward(sd1, cfg, o_ids, o_tt)["logits"][:, pos - 1, :])
            # same support up to tokens whose probability sits at the nucleus boundary, same distribution
            sup, o_sup = probs.cpu() > 0, o_probs > 0
            assert (sup ^ o_sup).sum().item() <= max(2, int(0.02 * o_sup.sum().item())), (pos, sup.sum(), o_sup.sum())
            assert (probs.cpu() - o_probs).abs().sum().item() < 0.05
            idx = torch.argmax(o_probs, dim=-1, keepdim=True)      # shared deterministic draw (multinomial in main.py:270)
            idx_item = idx.squeeze(-1).squeeze(-1).item()
            if idx_item == eos_id:
                break
            input_ids = torch.cat((input_ids, idx.cuda()), dim=-1)
            token_type_ids = torch.cat((token_type_ids, torch.LongTensor([[sp2_id]]).cuda()), dim=-1)
            o_ids = torch.cat((o_ids, idx), dim=-1)
            o_tt = torch.cat((o_tt, torch.LongTensor([[sp2_id]])), dim=-1)
            assert input_ids.shape == token_type_ids.shape


def test_missing_resize_raises_index_error(cuda_device):
    """Forgetting resize_token_embeddings for the speaker tokens: the reference raises IndexError / a device assert;
    here the kernels flag the row (no out-of-bounds access in forward or backward) and the host raises."""
    from transformers import GPT2Config
    from ergm_b200.model import GPT2LMHeadModel
    from ergm_b200 import ops
    hf = GPT2Config(vocab_size=1020, n_positions=64, n_embd=128, n_layer=1, n_head=2, attn_pdrop=0.0, resid_pdrop=0.0,
                    embd_pdrop=0.0)
    m = GPT2LMHeadModel(hf).to("cuda").train()
    b = synthetic.make_batch(2, 32, seed=3, vocab=1024, feat_dim=128)   # token types 1022 / 1023 >= vocab
    out = m(input_ids=b["input_ids"].cuda().clamp(max=1019), token_type_ids=b["token_type_ids"].cuda(),
            labels=b["labels"].cuda().clamp(max=1019))
    out.loss.backward()     # must not scatter outside the gradient tables
    torch.cuda.synchronize()
    assert torch.isfinite(m.transformer.wte.weight.grad).all()
    with pytest.raises(IndexError):
        ops.check_err_flag(torch.device("cuda", torch.cuda.current_device()))


def test_gpt2model_forward_and_past(cuda_device):
    """GPT2Model.forward (model.py:420-596): last_hidden_state + legacy past, through the LM model's backbone and as a
    stand-alone backbone."""
    from transformers import GPT2Config
    from ergm_b200.model import GPT2LMHeadModel, GPT2Model
    cfg = O.OracleConfig(vocab_size=1024, n_positions=128, n_embd=128, n_layer=2, n_head=2)
    sd = O.init_state_dict(cfg, seed=31, perturb=True)
    hf = GPT2Config(vocab_size=1024, n_positions=128, n_embd=128, n_layer=2, n_head=2)
    m = GPT2LMHeadModel(hf)
    m.load_state_dict(sd)
    m = m.to("cuda").eval()
    b = synthetic.make_batch(3, 40, seed=32, vocab=1024, feat_dim=128, ragged=False, tc=21)
    ids, tt, cap = b["input_ids"].cuda(), b["token_type_ids"].cuda(), b["caption_ids"].cuda()
    with torch.no_grad():
        h_ref, presents = O.backbone(sd, cfg, b["input_ids"], b["token_type_ids"], b["imgs"], b["auds"], b["caption_ids"])
    out = m.transformer(input_ids=ids, token_type_ids=tt, imgs=b["imgs"].cuda(), auds=b["auds"].cuda(), caption_ids=cap)
    assert tuple(out.last_hidden_state.shape) == (3, 40, 128)
    assert rel(out.last_hidden_state, h_ref) < 1e-2
    k0, v0 = out.past_key_values[0]
    assert tuple(k0.shape) == (3, 2, 40, 64) and rel(k0, presents[0][0]) < 1e-2 and rel(v0, presents[0][1]) < 1e-2
    # incremental: 30 tokens, then 10 more through past_key_values
    o1 = m.transformer(input_ids=ids[:, :30], token_type_ids=tt[:, :30], caption_ids=cap, use_cache=True)
    o2 = m.transformer(input_ids=ids[:, 30:], token_type_ids=tt[:, 30:], caption_ids=cap, past_key_values=o1.past_key_values)
    with torch.no_grad():
        h_nofuse, _ = O.backbone(sd, cfg, b["input_ids"], b["token_type_ids"], None, None, b["caption_ids"])
    assert rel(o2.last_hidden_state, h_nofuse[:, 30:]) < 1e-2
    tup = m.transformer(input_ids=ids, token_type_ids=tt, return_dict=False)
    assert isinstance(tup, tuple) and tup[0].shape == (3, 40, 128)
    # stand-alone backbone with the same weights
    bare = GPT2Model(hf)
    bare.load_state_dict({k[len("transformer."):]: v for k, v in sd.items() if k.startswith("transformer.")})
    bare = bare.to("cuda").eval()
    ob = bare(input_ids=ids, token_type_ids=tt, imgs=b["imgs"].cuda(), auds=b["auds"].cuda(), caption_ids=cap)
    assert rel(ob.last_hidden_state, h_ref) < 1e-2


def test_block_attention_mlp_forwards(cuda_device):
    """GPT2Block / GPT2Attention / GPT2MLP called directly (model.py:200-251, 262-267, 286-341) against the oracle's
    block functions."""
    from transformers import GPT2Config
    from ergm_b200.model import GPT2LMHeadModel
    cfg = O.OracleConfig(vocab_size=256, n_positions=64, n_embd=128, n_layer=1, n_head=2)
    sd = O.init_state_dict(cfg, seed=33, perturb=True)
    hf = GPT2Config(vocab_size=256, n_positions=64, n_embd=128, n_layer=1, n_head=2)
    m = GPT2LMHeadModel(hf)
    m.load_state_dict(sd)
    m = m.to("cuda").eval()
    blk = m.transformer.h[0]
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 20, 128, generator=g)
    enc = torch.randn(2, 9, 128, generator=g)
    p = "transformer.h.0."
    with torch.no_grad():
        want_mlp = O.mlp(sd, p + "mlp.", x)
        want_attn, (wk, wv) = O.self_attention(sd, p + "attn.", cfg, x)
        want_cross = O.cross_attention(sd, p + "crossattention.", cfg, x, enc)
        want_blk, _ = O.block(sd, 0, cfg, x, enc)
        want_blk_nc, _ = O.block(sd, 0, cfg, x, None)
    xc, ec = x.cuda(), enc.cuda()
    assert rel(blk.mlp(xc), want_mlp) < 1e-2
    a, present = blk.attn(xc, use_cache=True)
    assert rel(a, want_attn) < 1e-2 and rel(present[0], wk) < 1e-2 and rel(present[1], wv) < 1e-2
    assert blk.attn(xc)[1] is None
    assert rel(blk.crossattention(xc, encoder_hidden_states=ec)[0], want_cross) < 1e-2
    outs = blk(xc, encoder_hidden_states=ec, use_cache=True)
    assert len(outs) == 2 and rel(outs[0], want_blk) < 1e-2
    outs = blk(xc)
    assert len(outs) == 1 and rel(outs[0], want_blk_nc) < 1e-2
    # incremental self-attention through layer_past equals the full call
    a1, p1 = blk.attn(xc[:, :12], use_cache=True)
    a2, p2 = blk.attn(xc[:, 12:], layer_past=p1, use_cache=True)
    assert rel(torch.cat([a1, a2], 1), want_attn) < 1e-2 and tuple(p2[0].shape) == (2, 2, 20, 64)
    with pytest.raises(ValueError):
        blk.attn(xc, encoder_hidden_states=ec)   # no q_attn on a self-attention module (model.py:212-216)


def test_fused_adamw_skips_parameters_without_gradient(cuda_device):
    """torch.optim.AdamW skips parameters whose .grad is None: without caption_ids (how main.py:147 calls the model)
    the cross-attention tensors must not be weight-decayed, get no optimiser state, and the exported state dict equals
    torch's in which parameters it lists.  Eager FusedAdamW and the graph-captured train step both."""
    from transformers import GPT2Config
    from ergm_b200.model import GPT2LMHeadModel
    from ergm_b200.optim import FusedAdamW
    from ergm_b200.trainer import GraphedTrainStep
    cfg = O.OracleConfig(vocab_size=512, n_positions=64, n_embd=128, n_layer=2, n_head=2)
    sd = O.init_state_dict(cfg, seed=41, perturb=True)
    hf = GPT2Config(vocab_size=512, n_positions=64, n_embd=128, n_layer=2, n_head=2, attn_pdrop=0.0, resid_pdrop=0.0,
                    embd_pdrop=0.0)

    def build():
        m = GPT2LMHeadModel(hf)
        m.load_state_dict(sd)
        return m.to("cuda").train()

    b = synthetic.make_batch(3, 32, seed=42, vocab=512, feat_dim=128)
    kw = dict(input_ids=b["input_ids"].cuda(), token_type_ids=b["token_type_ids"].cuda(), labels=b["labels"].cuda(),
              emotion_labels=b["emotion_labels"].cuda())
    m, t = build(), build()
    fopt = FusedAdamW(m, lr=1e-2)
    topt = torch.optim.AdamW(t.parameters(), lr=1e-2)
    cross = "transformer.h.0.crossattention.c_attn.weight"
    pm, pt = dict(m.named_parameters()), dict(t.named_parameters())
    for _ in range(2):
        fopt.zero_grad()
        m(**kw).loss.backward()
        # both optimisers see the SAME gradients (Adam turns the sign of a noise-level gradient - e.g. the key bias,
        # whose true gradient is zero - into a full +-lr step, so two separately computed backwards cannot be compared)
        for n in pm:
            pt[n].grad = pm[n].grad.clone() if pm[n].grad is not None else None
        assert pm[cross].grad is None
        fopt.step()
        topt.step()
    assert torch.equal(pm[cross].detach().cpu(), sd[cross])            # untouched: no decay without a gradient
    assert torch.equal(pt[cross].detach().cpu(), sd[cross])
    worst = max((pm[n].detach() - pt[n].detach()).abs().max().item() for n in pm)
    assert worst < 5e-6, worst
    fs, ts = fopt.state_dict(), topt.state_dict()
    assert sorted(fs["state"].keys()) == sorted(ts["state"].keys())    # no state for the skipped parameters
    assert len(fs["state"]) < len(fs["param_groups"][0]["params"])
    # graph-captured step: same rule, staged before the backward from the batch keys
    g = build()
    step = GraphedTrainStep(g, FusedAdamW(g, lr=1e-2))
    host = {k: v.cpu().pin_memory() for k, v in kw.items()}
    for _ in range(3):
        step(host)
    assert torch.equal(dict(g.named_parameters())[cross].detach().cpu(), sd[cross])
    assert not torch.equal(g.transformer.h[0].mlp.c_fc.weight.detach().cpu(), sd["transformer.h.0.mlp.c_fc.weight"])


def test_summed_losses_need_separate_backwards(cuda_device):
    """One set of saved activations per model: backpropagating through an OLDER training forward raises a clear
    error instead of silently using the later forward's activations."""
    from transformers import GPT2Config
    from ergm_b200.model import GPT2LMHeadModel
    hf = GPT2Config(vocab_size=512, n_positions=64, n_embd=128, n_layer=1, n_head=2, attn_pdrop=0.0, resid_pdrop=0.0,
                    embd_pdrop=0.0)
    m = GPT2LMHeadModel(hf).to("cuda").train()
    b = synthetic.make_batch(2, 32, seed=43, vocab=512, feat_dim=128)
    kw = dict(input_ids=b["input_ids"].cuda(), token_type_ids=b["token_type_ids"].cuda(), labels=b["labels"].cuda())
    la = m(**kw).loss
    lb = m(**kw).loss
    with pytest.raises(RuntimeError, match="ONE training forward|saved activations"):
        (la + lb).backward()
    # the supported pattern: backward each loss before the next forward (gradients accumulate)
    m.zero_grad()
    m(**kw).loss.backward()
    g1 = m.transformer.wpe.weight.grad.clone()
    m(**kw).loss.backward()
    assert torch.allclose(m.transformer.wpe.weight.grad, 2 * g1, rtol=1e-3, atol=1e-7)

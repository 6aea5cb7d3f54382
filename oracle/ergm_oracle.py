"""CPU restatement of the ERGM multimodal GPT-2 hot path — TEST INFRASTRUCTURE ONLY.

This file is the parity oracle for ergm_b200.  It restates, op for op, the arithmetic of
/root/reference/src/model.py (a fork of HuggingFace modeling_gpt2.py 4.26) as plain
functional PyTorch over a state-dict, so that it can travel to the GPU box (where
/root/reference does not exist) and be evaluated in fp32 or fp64 on the CPU.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import it.  The product path (ergm_b200/) never does.

Pinning: the reference ships no tests or golden vectors.  The oracle is
pinned instead against the reference itself: oracle/make_golden.py imports the unmodified
reference (oracle/ref_shim.py) in the build container, checks that this restatement is
bit-identical to it on seeded inputs, and commits known-answer fixtures under tests/golden/.

Third-party arithmetic restated here (absent from /root/reference, pinned there by
requirements.txt:212,218 to torch==1.13.1 / transformers==4.26.1):
  * transformers.pytorch_utils.Conv1D.forward  : addmm(bias, x.view(-1,K), W[K,N])
  * transformers.activations.NewGELUActivation : 0.5x(1+tanh(sqrt(2/pi)(x+0.044715x^3)))
  * torch.nn.LayerNorm / CrossEntropyLoss / softmax : the installed torch CPU kernels.
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

NUM_EMOTIONS = 7  # model.py:607


class OracleConfig:
    """The GPT2Config fields the path reads (model.py:64-104, 271-284, 383-396)."""

    def __init__(self, vocab_size=50260, n_positions=1024, n_embd=768, n_layer=12, n_head=12,
                 n_inner=None, layer_norm_epsilon=1e-5, initializer_range=0.02, visual_dim=None, audio_dim=None):
        self.vocab_size = vocab_size
        self.n_positions = n_positions
        self.n_embd = n_embd
        self.n_layer = n_layer
        self.n_head = n_head
        self.n_inner = n_inner if n_inner is not None else 4 * n_embd
        self.layer_norm_epsilon = layer_norm_epsilon
        self.initializer_range = initializer_range
        self.visual_dim, self.audio_dim = visual_dim, audio_dim  # A3 extension (None = reference layout)

    @property
    def head_dim(self):
        return self.n_embd // self.n_head


# --------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------
def param_shapes(cfg: OracleConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """State-dict keys and shapes of reference GPT2LMHeadModel, in module
    registration order; lm_head.weight is tied to transformer.wte.weight (model.py:600)."""
    H, I = cfg.n_embd, cfg.n_inner
    out = [("transformer.wte.weight", (cfg.vocab_size, H)), ("transformer.wpe.weight", (cfg.n_positions, H))]
    for i in range(cfg.n_layer):
        p = "transformer.h.%d." % i
        out += [
            (p + "ln_1.weight", (H,)), (p + "ln_1.bias", (H,)),
            (p + "attn.c_attn.weight", (H, 3 * H)), (p + "attn.c_attn.bias", (3 * H,)),
            (p + "attn.c_proj.weight", (H, H)), (p + "attn.c_proj.bias", (H,)),
            (p + "ln_2.weight", (H,)), (p + "ln_2.bias", (H,)),
            (p + "crossattention.c_attn.weight", (H, 2 * H)), (p + "crossattention.c_attn.bias", (2 * H,)),
            (p + "crossattention.q_attn.weight", (H, H)), (p + "crossattention.q_attn.bias", (H,)),
            (p + "crossattention.c_proj.weight", (H, H)), (p + "crossattention.c_proj.bias", (H,)),
            (p + "ln_cross_attn.weight", (H,)), (p + "ln_cross_attn.bias", (H,)),
            (p + "mlp.c_fc.weight", (H, I)), (p + "mlp.c_fc.bias", (I,)),
            (p + "mlp.c_proj.weight", (I, H)), (p + "mlp.c_proj.bias", (H,)),
        ]
    out += [("transformer.ln_f.weight", (H,)), ("transformer.ln_f.bias", (H,)),
            ("emotion_head.weight", (NUM_EMOTIONS, H))]
    if cfg.visual_dim:  # A3 extension parameters (absent from the reference state dict)
        out += [("visual_proj.weight", (H, cfg.visual_dim)), ("visual_proj.bias", (H,)),
                ("audio_proj.weight", (H, cfg.audio_dim)), ("audio_proj.bias", (H,))]
    return out


def init_state_dict(cfg: OracleConfig, seed: int = 0, perturb: bool = False,
                    dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random-init weights with the distributions of reference _init_weights (model.py:359-375):
    N(0, initializer_range) for Linear / Conv1D / Embedding weights, N(0, range/sqrt(2L)) for
    every c_proj.weight, LayerNorm (1, 0), zero biases.  Each tensor draws from its own
    torch.Generator seeded by (seed, index) so the result does not depend on module
    construction order and is reproducible on any host.  perturb=True additionally makes
    biases and LayerNorm affine parameters non-trivial so that tests exercise them."""
    sd = {}
    for idx, (name, shape) in enumerate(param_shapes(cfg)):
        g = torch.Generator().manual_seed(seed * 1000003 + idx)
        leaf = name.rsplit(".", 2)[-2] if name.count(".") >= 2 else ""
        is_ln = ".ln_" in name or "ln_f" in name
        if is_ln and name.endswith("weight"):
            t = torch.ones(shape)
            if perturb:
                t = t + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("bias"):
            t = torch.zeros(shape)
            if perturb:
                t = cfg.initializer_range * torch.randn(shape, generator=g)
        else:
            std = cfg.initializer_range
            if leaf == "c_proj":
                std = cfg.initializer_range / math.sqrt(2 * cfg.n_layer)
            t = torch.randn(shape, generator=g) * std
        sd[name] = t.to(dtype)
    sd["lm_head.weight"] = sd["transformer.wte.weight"]
    return sd


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def conv1d(x, w, b):
    """transformers Conv1D.forward: addmm(bias, x.view(-1, K), W[K, N])."""
    size_out = x.size()[:-1] + (w.shape[1],)
    return torch.addmm(b, x.reshape(-1, x.size(-1)), w).view(size_out)


def gelu_new(x):
    """transformers NewGELUActivation.forward (ACT2FN['gelu_new'], model.py:259)."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def split_heads(t, nh, hd):
    """model.py:190-193"""
    return t.view(t.size()[:-1] + (nh, hd)).permute(0, 2, 1, 3)


def merge_heads(t, nh, hd):
    """model.py:195-198"""
    t = t.permute(0, 2, 1, 3).contiguous()
    return t.view(t.size()[:-2] + (nh * hd,))


def attn_core(q, k, v, causal: bool, attention_mask=None):
    """GPT2Attention._attn, model.py:119-148 (dropout omitted: parity runs use eval / p=0)."""
    w = torch.matmul(q, k.transpose(-1, -2))
    w = w / torch.full([], v.size(-1) ** 0.5, dtype=w.dtype)
    if causal:
        ql, kl = q.size(-2), k.size(-2)
        mask = torch.tril(torch.ones((kl, kl), dtype=torch.bool))[kl - ql:kl, :kl]
        w = torch.where(mask, w, torch.full([], torch.finfo(w.dtype).min, dtype=w.dtype))
    if attention_mask is not None:
        w = w + attention_mask
    w = F.softmax(w, dim=-1)
    return torch.matmul(w, v)


def self_attention(sd, pfx, cfg, x, layer_past=None, attention_mask=None):
    """GPT2Attention.forward, self branch (model.py:222-246). Returns (out, (k, v))."""
    nh, hd, H = cfg.n_head, cfg.head_dim, cfg.n_embd
    q, k, v = conv1d(x, sd[pfx + "c_attn.weight"], sd[pfx + "c_attn.bias"]).split(H, dim=2)
    q, k, v = split_heads(q, nh, hd), split_heads(k, nh, hd), split_heads(v, nh, hd)
    if layer_past is not None:
        k = torch.cat((layer_past[0], k), dim=-2)
        v = torch.cat((layer_past[1], v), dim=-2)
    o = attn_core(q, k, v, True, attention_mask)
    o = conv1d(merge_heads(o, nh, hd), sd[pfx + "c_proj.weight"], sd[pfx + "c_proj.bias"])
    return o, (k, v)


def cross_attention(sd, pfx, cfg, x, enc, enc_mask=None):
    """GPT2Attention.forward, cross branch (model.py:211-220, 241-245): q from x, k/v from the
    caption embeddings, no causal mask, additive (zero) encoder mask."""
    nh, hd, H = cfg.n_head, cfg.head_dim, cfg.n_embd
    q = conv1d(x, sd[pfx + "q_attn.weight"], sd[pfx + "q_attn.bias"])
    k, v = conv1d(enc, sd[pfx + "c_attn.weight"], sd[pfx + "c_attn.bias"]).split(H, dim=2)
    q, k, v = split_heads(q, nh, hd), split_heads(k, nh, hd), split_heads(v, nh, hd)
    o = attn_core(q, k, v, False, enc_mask)
    return conv1d(merge_heads(o, nh, hd), sd[pfx + "c_proj.weight"], sd[pfx + "c_proj.bias"])


def mlp(sd, pfx, x):
    """GPT2MLP.forward, model.py:262-267"""
    h = gelu_new(conv1d(x, sd[pfx + "c_fc.weight"], sd[pfx + "c_fc.bias"]))
    return conv1d(h, sd[pfx + "c_proj.weight"], sd[pfx + "c_proj.bias"])


def layer_norm(sd, pfx, cfg, x):
    return F.layer_norm(x, (cfg.n_embd,), sd[pfx + "weight"], sd[pfx + "bias"], cfg.layer_norm_epsilon)


def block(sd, i, cfg, x, enc, layer_past=None, attention_mask=None, enc_mask=None):
    """GPT2Block.forward, model.py:286-341"""
    p = "transformer.h.%d." % i
    a, present = self_attention(sd, p + "attn.", cfg, layer_norm(sd, p + "ln_1.", cfg, x), layer_past, attention_mask)
    x = a + x
    if enc is not None:
        c = cross_attention(sd, p + "crossattention.", cfg, layer_norm(sd, p + "ln_cross_attn.", cfg, x), enc, enc_mask)
        x = x + c
    m = mlp(sd, p + "mlp.", layer_norm(sd, p + "ln_2.", cfg, x))
    return x + m, present


# --------------------------------------------------------------------------------------
# model
# --------------------------------------------------------------------------------------
def modality_pool_proj(sd, vis_seq, aud_seq):
    """A3 EXTENSION — not in the reference model (parity unpinned for this function; SURVEY.md §8 A3,
    Appendix A D7).  Restates the reference's OFFLINE pooling, feature_extraction.py:63 (audio
    `last_hidden_state.mean(dim=1)`) and :69 (visual `.mean(dim=1)`), followed by a learned
    Linear(D -> H) per modality (`visual_proj`, `audio_proj`) so that 768-wide features can feed a
    1024-wide backbone.  Returns (imgs [B,1,H], auds [B,H]) in the layout model.py:497-498 indexes."""
    v = F.linear(vis_seq.mean(dim=1), sd["visual_proj.weight"], sd["visual_proj.bias"])
    a = F.linear(aud_seq.mean(dim=1), sd["audio_proj.weight"], sd["audio_proj.bias"])
    return v[:, None, :], a


def backbone(sd, cfg, input_ids, token_type_ids=None, imgs=None, auds=None, caption_ids=None,
             past_key_values=None, attention_mask=None, position_ids=None):
    """GPT2Model.forward, model.py:420-596.  caption_ids=None skips cross-attention (the
    reference crashes there, model.py:521; boundary decision (1) of SURVEY.md §8b — with the
    one-line guard the reference is bit-equal to stock HF GPT-2)."""
    wte, wpe = sd["transformer.wte.weight"], sd["transformer.wpe.weight"]
    B, T = input_ids.shape
    inputs_embeds = F.embedding(input_ids, wte)  # :459
    enc = None
    if caption_ids is not None:
        # reference does caption_ids.view(-1, T) (:461), forcing Tc == T; arbitrary Tc is a
        # documented extension (block-level reference accepts any encoder length)
        enc = F.embedding(caption_ids.view(B, -1), wte)  # :460-463
    past_len = 0 if past_key_values is None else past_key_values[0][0].size(-2)
    if position_ids is None:
        position_ids = torch.arange(past_len, T + past_len, dtype=torch.long).unsqueeze(0)  # :474-476
    add_mask = None
    if attention_mask is not None:  # :478-482
        am = attention_mask.view(B, -1)[:, None, None, :].to(wte.dtype)
        add_mask = (1.0 - am) * torch.finfo(wte.dtype).min
    enc_mask = None
    if enc is not None:  # :484-489: all-ones mask inverted -> additive zeros
        enc_mask = torch.zeros(B, 1, 1, enc.shape[1], dtype=wte.dtype)
    if imgs is not None and "visual_proj.weight" in sd:
        imgs, auds = modality_pool_proj(sd, imgs, auds)
    if imgs is not None:  # :495-498, in place on the wte output, before wpe / token types
        inputs_embeds = inputs_embeds.clone()
        for i in range(B):
            inputs_embeds[i, 0] = inputs_embeds[i, 0] + imgs[i][0]
            inputs_embeds[i, 1] = inputs_embeds[i, 1] + auds[i].unsqueeze(0)
    h = inputs_embeds + F.embedding(position_ids, wpe)  # :500-501
    if token_type_ids is not None:
        h = h + F.embedding(token_type_ids, wte)  # :502-504 (types looked up in the WORD table)
    presents = []
    for i in range(cfg.n_layer):
        lp = None if past_key_values is None else past_key_values[i]
        h, present = block(sd, i, cfg, h, enc, lp, add_mask, enc_mask)
        presents.append(present)
    h = layer_norm(sd, "transformer.ln_f.", cfg, h)  # :578
    return h, tuple(presents)


def forward(sd, cfg, input_ids, token_type_ids=None, labels=None, emotion_labels=None, imgs=None,
            auds=None, caption_ids=None, past_key_values=None, attention_mask=None,
            position_ids=None):
    """GPT2LMHeadModel.forward, model.py:654-737. Returns dict(loss, lm_loss, emotion_loss,
    logits, emotion_logits, past_key_values, hidden)."""
    h, presents = backbone(sd, cfg, input_ids, token_type_ids, imgs, auds, caption_ids,
                           past_key_values, attention_mask, position_ids)
    logits = F.linear(h, sd["transformer.wte.weight"])  # :698, tied head
    emo_logits = F.linear(h[:, -1, :], sd["emotion_head.weight"])  # :700-701
    lm_loss = emo_loss = loss = None
    if labels is not None:  # :705-708 / :715-718
        sl = logits[..., :-1, :].contiguous()
        tl = labels[..., 1:].contiguous()
        lm_loss = F.cross_entropy(sl.view(-1, sl.size(-1)), tl.view(-1))
    if emotion_labels is not None:  # :710-711 / :720-721
        emo_loss = F.cross_entropy(emo_logits.view(-1, NUM_EMOTIONS), emotion_labels.view(-1))
    if lm_loss is not None and emo_loss is not None:
        loss = lm_loss + emo_loss  # :713
    elif lm_loss is not None:
        loss = lm_loss
    elif emo_loss is not None:
        loss = emo_loss
    return dict(loss=loss, lm_loss=lm_loss, emotion_loss=emo_loss, logits=logits,
                emotion_logits=emo_logits, past_key_values=presents, hidden=h)


# --------------------------------------------------------------------------------------
# generation (main.py:253-282 pattern and the model's own KV-cache surface)
# --------------------------------------------------------------------------------------
def top_p_filter_reference(probs, top_p):
    """Nucleus filter exactly as main.py:258-269 (note the shift-right-by-one of the mask,
    :263-265): returns re-normalised probabilities in vocabulary order."""
    sorted_probs, sorted_idx = torch.sort(probs, dim=-1, descending=True)
    cum = torch.cumsum(sorted_probs, dim=-1)
    remove = cum > top_p
    remove[..., 1:] = remove[..., :-1].clone()
    remove[..., 0] = False
    sorted_probs = sorted_probs.masked_fill(remove, 0.0)
    sorted_probs = sorted_probs / sorted_probs.sum(dim=-1, keepdim=True)
    out = torch.zeros_like(probs)
    out.scatter_(-1, sorted_idx, sorted_probs)
    return out


def greedy_generate_recompute(sd, cfg, input_ids, token_type_ids, max_new_tokens, sp2_id=50259,
                              eos_id=50256, caption_ids=None, imgs=None, auds=None):
    """The reference decode loop (main.py:255-279) with argmax instead of multinomial: one FULL
    forward per new token, next-token logits read at the last position, the new token gets
    speaker type sp2.  Batch rows are independent; all rows run max_new_tokens steps (finished
    rows keep emitting eos)."""
    ids, tt = input_ids.clone(), token_type_ids.clone()
    B = ids.shape[0]
    done = torch.zeros(B, dtype=torch.bool)
    out = []
    for _ in range(max_new_tokens):
        cap = caption_ids
        if cap is not None and cap.shape[1] != ids.shape[1]:
            pass  # arbitrary Tc extension; reference would need Tc == T
        r = forward(sd, cfg, ids, tt, imgs=imgs, auds=auds, caption_ids=cap)
        nxt = r["logits"][:, -1, :].argmax(-1)
        nxt = torch.where(done, torch.full_like(nxt, eos_id), nxt)
        done |= nxt == eos_id
        out.append(nxt)
        ids = torch.cat([ids, nxt[:, None]], 1)
        tt = torch.cat([tt, torch.full((B, 1), sp2_id, dtype=tt.dtype)], 1)
    return torch.stack(out, 1)


def greedy_generate_cached(sd, cfg, input_ids, token_type_ids, max_new_tokens, sp2_id=50259,
                           eos_id=50256, caption_ids=None, imgs=None, auds=None):
    """Same decode through the model's KV-cache surface (model.py:228-236, 469-476): prefill
    once, then one-token steps with past_key_values.  Multimodal fusion is applied on the
    prefill only (boundary decision (3))."""
    B = input_ids.shape[0]
    r = forward(sd, cfg, input_ids, token_type_ids, imgs=imgs, auds=auds, caption_ids=caption_ids)
    past = r["past_key_values"]
    done = torch.zeros(B, dtype=torch.bool)
    out = []
    logits = r["logits"][:, -1, :]
    for step in range(max_new_tokens):
        nxt = logits.argmax(-1)
        nxt = torch.where(done, torch.full_like(nxt, eos_id), nxt)
        done |= nxt == eos_id
        out.append(nxt)
        if step + 1 == max_new_tokens:
            break
        tt = torch.full((B, 1), sp2_id, dtype=token_type_ids.dtype)
        r = forward(sd, cfg, nxt[:, None], tt, caption_ids=caption_ids, past_key_values=past)
        past = r["past_key_values"]
        logits = r["logits"][:, -1, :]
    return torch.stack(out, 1)


# --------------------------------------------------------------------------------------
# north_star extension: pooled + projected feature sequences (unpinned by the reference)
# --------------------------------------------------------------------------------------
def pool_project(seq, weight=None, bias=None):
    """Mean over time/patches (feature_extraction.py:63,69), then optional Linear(D -> H).
    With weight=None and D == H this is exactly the pre-pooled tensor model.py:497-498 adds."""
    pooled = seq.mean(dim=1)
    if weight is not None:
        pooled = F.linear(pooled, weight, bias)
    return pooled


# --------------------------------------------------------------------------------------
# optimiser row: torch.optim.AdamW semantics, main.py:68,155
# --------------------------------------------------------------------------------------
def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01):
    """One AdamW update, same arithmetic order as torch.optim.AdamW (single tensor path)."""
    p = p * (1 - lr * weight_decay)
    m = m + (g - m) * (1 - beta1)  # lerp
    v = v * beta2 + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def poly_decay_lr(step, base_lr, warmup_steps, total_steps, lr_end=1e-7, power=2.0):
    """transformers get_polynomial_decay_schedule_with_warmup lambda (main.py:93-95)."""
    if step < warmup_steps:
        return base_lr * float(step) / float(max(1, warmup_steps))
    if step > total_steps:
        return lr_end
    lr_range = base_lr - lr_end
    decay_steps = total_steps - warmup_steps
    pct_remaining = 1 - (step - warmup_steps) / decay_steps
    return lr_range * pct_remaining ** power + lr_end

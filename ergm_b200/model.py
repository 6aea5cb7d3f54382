"""Drop-in replacement for /root/reference/src/model.py on B200.

Exports the names ERGM's driver imports with `from model import *` (main.py:22):
GPT2LMHeadModel, GPT2Model, GPT2Block, GPT2Attention, GPT2MLP,
CausalLMOutputWithEmotionClassification.  Constructor, forward keyword surface, output
fields, state_dict keys / shapes and loss definition follow the reference (file:line cited
inline); the arithmetic runs in the hand-written sm_100a kernels of libergm_b200.so through
ergm_b200.engine — there is no PyTorch-eager or CPU fallback, a forward on a non-CUDA model
raises.

Documented boundary decisions:
  1. caption_ids=None skips cross-attention (the reference raises UnboundLocalError at
     model.py:521; with the one-line guard it equals stock HF GPT-2).
  2. caption_ids may have any length Tc (the reference's .view(-1, T) at :461 forces Tc == T).
  3. imgs/auds fusion (:495-498) is applied only when there is no past (prefill).
  4. The emotion head reads the last position, padded or not (:700) — kept.
  5. .logits / .past_key_values are materialised lazily from device buffers that stay valid
     until the next forward of the same model.
"""
import math
import os
from typing import Optional, Tuple

import torch
from torch import nn
from torch.nn import functional as F  # noqa: F401  (re-exported like the reference's `import *`)
from transformers.modeling_utils import PreTrainedModel
from transformers.models.gpt2.configuration_gpt2 import GPT2Config
from transformers.pytorch_utils import Conv1D

from . import _lib as L
from . import blocks
from .engine import Engine

NUM_EMOTIONS = 7  # model.py:607


class CausalLMOutputWithEmotionClassification:
    """Same fields as the reference dataclass (model.py:48-60); `logits` and `past_key_values`
    are produced on first access from the engine's device buffers."""

    _fields = ("loss", "logits", "emotion_logits", "past_key_values", "hidden_states", "attentions",
               "cross_attentions")

    def __init__(self, loss=None, logits_fn=None, emotion_logits=None, past_fn=None, lm_loss=None,
                 emotion_loss=None):
        self.loss = loss
        self._logits_fn = logits_fn
        self._logits = None
        self.emotion_logits = emotion_logits
        self._past_fn = past_fn
        self._past = None
        self.hidden_states = None
        self.attentions = None
        self.cross_attentions = None
        self.lm_loss = lm_loss
        self.emotion_loss = emotion_loss

    @property
    def logits(self):
        if self._logits is None and self._logits_fn is not None:
            self._logits = self._logits_fn()
        return self._logits

    @property
    def past_key_values(self):
        if self._past is None and self._past_fn is not None:
            self._past = self._past_fn()
        return self._past

    def keys(self):
        return [k for k in self._fields if getattr(self, k) is not None]

    def to_tuple(self):
        return tuple(getattr(self, k) for k in self.keys())

    def __getitem__(self, k):
        if isinstance(k, str):
            return getattr(self, k)
        return self.to_tuple()[k]

    def __iter__(self):
        return iter(self.to_tuple())

    def __len__(self):
        return len(self.keys())


class BaseModelOutputWithPastAndCrossAttentions:
    def __init__(self, last_hidden_state, past_fn=None):
        self.last_hidden_state = last_hidden_state
        self._past_fn = past_fn
        self._past = None
        self.hidden_states = self.attentions = self.cross_attentions = None

    @property
    def past_key_values(self):
        if self._past is None and self._past_fn is not None:
            self._past = self._past_fn()
        return self._past

    def __getitem__(self, i):
        return (self.last_hidden_state, self.past_key_values)[i]


# ----------------------------------------------------------------------------------------------
# parameter containers: same module tree / names / shapes as the reference
# ----------------------------------------------------------------------------------------------
class GPT2Attention(nn.Module):
    """Parameters of model.py:64-104; the arithmetic (:119-251) lives in ergm_attn_fwd/bwd and
    ergm_gemm_bf16."""

    def __init__(self, config, is_cross_attention=False, layer_idx=None):
        super().__init__()
        self.embed_dim = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = self.embed_dim // self.num_heads
        if self.head_dim * self.num_heads != self.embed_dim:
            raise ValueError(
                f"`embed_dim` must be divisible by num_heads (got `embed_dim`: {self.embed_dim} and `num_heads`:"
                f" {self.num_heads}).")
        self.is_cross_attention = is_cross_attention
        self.layer_idx = layer_idx
        if is_cross_attention:
            self.c_attn = Conv1D(2 * self.embed_dim, self.embed_dim)
            self.q_attn = Conv1D(self.embed_dim, self.embed_dim)
        else:
            self.c_attn = Conv1D(3 * self.embed_dim, self.embed_dim)
        self.c_proj = Conv1D(self.embed_dim, self.embed_dim)

    forward = blocks.attention_forward  # model.py:200-251 (inference surface; see ergm_b200/blocks.py)


class GPT2MLP(nn.Module):
    def __init__(self, intermediate_size, config):
        super().__init__()
        self.c_fc = Conv1D(intermediate_size, config.hidden_size)
        self.c_proj = Conv1D(config.hidden_size, intermediate_size)

    forward = blocks.mlp_forward  # model.py:262-267


class GPT2Block(nn.Module):
    def __init__(self, config, layer_idx=None):
        super().__init__()
        hidden_size = config.hidden_size
        inner_dim = config.n_inner if config.n_inner is not None else 4 * hidden_size
        config.add_cross_attention = True  # model.py:275: every block owns a cross-attention
        self.ln_1 = nn.LayerNorm(hidden_size, eps=config.layer_norm_epsilon)
        self.attn = GPT2Attention(config, layer_idx=layer_idx)
        self.ln_2 = nn.LayerNorm(hidden_size, eps=config.layer_norm_epsilon)
        self.crossattention = GPT2Attention(config, is_cross_attention=True, layer_idx=layer_idx)
        self.ln_cross_attn = nn.LayerNorm(hidden_size, eps=config.layer_norm_epsilon)
        self.mlp = GPT2MLP(inner_dim, config)

    forward = blocks.block_forward  # model.py:286-341


class GPT2PreTrainedModel(PreTrainedModel):
    config_class = GPT2Config
    base_model_prefix = "transformer"
    _no_split_modules = ["GPT2Block"]
    _skip_keys_device_placement = "past_key_values"

    def _init_weights(self, module):
        """model.py:359-375"""
        if isinstance(module, (nn.Linear, Conv1D)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        for name, p in module.named_parameters():
            if name == "c_proj.weight":
                p.data.normal_(mean=0.0, std=(self.config.initializer_range / math.sqrt(2 * self.config.n_layer)))


class GPT2Model(GPT2PreTrainedModel):
    def __init__(self, config):
        super().__init__(config)
        self.embed_dim = config.hidden_size
        self.wte = nn.Embedding(config.vocab_size, self.embed_dim)
        self.wpe = nn.Embedding(config.max_position_embeddings, self.embed_dim)
        self.h = nn.ModuleList([GPT2Block(config, layer_idx=i) for i in range(config.num_hidden_layers)])
        self.ln_f = nn.LayerNorm(self.embed_dim, eps=config.layer_norm_epsilon)
        self.post_init()

        self._owner = None    # weakref to the GPT2LMHeadModel this backbone belongs to (shares its engine)
        self._engine = None   # stand-alone backbone: its own engine (parameter names get the "transformer." prefix)

    def get_input_embeddings(self):
        return self.wte

    def set_input_embeddings(self, new_embeddings):
        self.wte = new_embeddings

    @property
    def engine(self):
        owner = self._owner() if self._owner is not None else None
        if owner is not None:
            return owner.engine
        if self._engine is None:
            dev = self.wte.weight.device
            if dev.type != "cuda":
                raise L.ErgmError("ergm_b200 has no CPU path: call .to('cuda') on the model first")
            with torch.cuda.device(dev):
                self._engine = Engine(self, prefix="transformer.")
        return self._engine

    @torch.no_grad()
    def forward(self, input_ids=None, past_key_values=None, attention_mask=None, token_type_ids=None,
                position_ids=None, head_mask=None, inputs_embeds=None, encoder_hidden_states=None,
                encoder_attention_mask=None, use_cache=None, output_attentions=None, output_hidden_states=None,
                return_dict=None, imgs=None, auds=None, caption_ids=None):
        """GPT2Model.forward, model.py:420-596: embeddings + fusion, L blocks, ln_f.  Returns
        BaseModelOutputWithPastAndCrossAttentions(last_hidden_state [B, T, H] fp32, past_key_values) or the tuple
        form.  Inference surface (no autograd graph): the training path is GPT2LMHeadModel.forward."""
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if input_ids is None:
            if inputs_embeds is not None:
                raise L.ErgmError("inputs_embeds is not supported: the embedding stage is fused (ergm_embed_fuse_fwd)")
            raise ValueError("You have to specify either input_ids or inputs_embeds")
        _get_head_mask_check(head_mask)
        if output_attentions or output_hidden_states:
            raise L.ErgmError("output_attentions / output_hidden_states are not available: attention probabilities "
                              "are never materialised by the fused kernels")
        if encoder_hidden_states is not None:
            raise L.ErgmError("pass caption_ids (model.py:460-463 embeds them with wte), not encoder_hidden_states")
        return_dict = return_dict if return_dict is not None else getattr(self.config, "return_dict", True)
        use_cache = use_cache if use_cache is not None else self.config.use_cache
        eng = self.engine
        dev = eng.device
        input_ids = input_ids.reshape(-1, input_ids.shape[-1]).to(dev, torch.int64).contiguous()
        B, T = input_ids.shape
        tt = token_type_ids.reshape(B, T).to(dev, torch.int64).contiguous() if token_type_ids is not None else None
        cap = caption_ids.reshape(B, -1).to(dev, torch.int64).contiguous() if caption_ids is not None else None
        pos = None
        if position_ids is not None:
            position_ids = position_ids.reshape(-1, T)
            if position_ids.shape[0] != 1 and not bool((position_ids == position_ids[:1]).all().item()):
                raise L.ErgmError("per-sample position_ids are not supported (model.py:474-476 uses a shared arange)")
            pos = position_ids[0].to(dev, torch.int64).contiguous()
        with torch.cuda.device(dev):
            legacy, past_len = None, 0
            if past_key_values is not None:
                from .generation import legacy_past_to_rows
                legacy = legacy_past_to_rows(past_key_values, B, eng.H, dev)
                past_len = legacy[0].shape[1]
            kv_lens = None
            if attention_mask is not None:
                kv_lens = _kv_lens_from_mask(attention_mask.to(dev), B, past_len + T)
            out = eng.forward(input_ids, tt, None, None, imgs if legacy is None else None,
                              auds if legacy is None else None, cap, pos, kv_lens=kv_lens, training=False, save=False,
                              heads=False, legacy_past=legacy)
            from . import ops
            hidden = torch.empty(B * T, eng.H, dtype=torch.float32, device=dev)
            ops.ln_fwd(out["x_final"], eng.p("transformer.ln_f.weight"), eng.p("transformer.ln_f.bias"), None, hidden,
                       None, None, self.config.layer_norm_epsilon)  # model.py:578
            hidden = hidden.view(B, T, eng.H)
            past_fn = _past_fn(out["kv_present"], B, eng.H, eng.nh, legacy is not None) if use_cache else None
        ret = BaseModelOutputWithPastAndCrossAttentions(hidden, past_fn)
        if not return_dict:
            return (hidden,) + ((ret.past_key_values,) if use_cache else ())
        return ret


def _kv_lens_from_mask(attention_mask, B, total):
    """Right-padded 0/1 attention mask [B, past + T] (model.py:478-482) -> per-sequence key counts."""
    am = attention_mask.reshape(B, -1)
    if am.shape[1] != total:
        raise ValueError("attention_mask must cover past + current tokens (model.py:478-482)")
    ok = bool(((am[:, 1:] <= am[:, :-1]).all() & (am[:, 0] > 0).all()).item())
    if not ok:
        raise L.ErgmError("only right-padded attention masks (ones then zeros) are supported")
    return am.to(torch.int64).sum(1).to(torch.int32)


def _past_fn(kv_bufs, B, H, nh, legacy):
    """Lazy legacy-tuple view of the K/V the forward produced: L x (k, v), each [B, nh, ctx, 64] fp32
    (model.py:228-236).  kv_bufs: per layer the [B*T, 3H] qkv matrix, or (legacy) the [B, ctx, 2H] K|V rows."""
    def fn():
        res = []
        for buf in kv_bufs:
            if legacy:
                tk = buf.shape[1]
                k = buf[:, :, :H].view(B, tk, nh, 64).permute(0, 2, 1, 3).float()
                v = buf[:, :, H:].view(B, tk, nh, 64).permute(0, 2, 1, 3).float()
            else:
                v5 = buf.view(B, -1, 3, nh, 64)
                k, v = v5[:, :, 1].permute(0, 2, 1, 3).float(), v5[:, :, 2].permute(0, 2, 1, 3).float()
            res.append((k, v))
        return tuple(res)
    return fn


def _get_head_mask_check(head_mask):
    if head_mask is not None:
        raise L.ErgmError("head_mask is not supported by the fused attention kernels (main.py never passes it)")


class _LossFn(torch.autograd.Function):
    """Connects the hand-written backward (Engine.backward) to loss.backward().  `anchor` is a
    real parameter used only so that autograd schedules this node; parameter gradients are
    written straight into the flat gradient buffer and bound to p.grad by the engine."""

    @staticmethod
    def forward(ctx, anchor, engine, losses, on_backward, forward_id):
        ctx.engine = engine
        ctx.on_backward = on_backward
        ctx.forward_id = forward_id
        return losses[0].clone()

    @staticmethod
    def backward(ctx, g):
        eng = ctx.engine
        g = g.detach().reshape(1).to(torch.float32).contiguous()
        accumulate = eng.store.grads_live()
        sv = eng.saved
        if sv is None or sv.get("id") != ctx.forward_id:
            eng.backward(g, forward_id=ctx.forward_id)  # raises the explanatory error
        if ctx.on_backward is not None:
            ctx.on_backward(g, accumulate)
        else:
            eng.backward(g, accumulate=accumulate, forward_id=ctx.forward_id)
        return None, None, None, None, None


class GPT2LMHeadModel(GPT2PreTrainedModel):
    _tied_weights_keys = {"lm_head.weight": "transformer.wte.weight"}

    def __init__(self, config):
        super().__init__(config)
        if not hasattr(config, "n_ctx"):
            config.n_ctx = config.n_positions  # main.py:64 reads config.n_ctx
        self.transformer = GPT2Model(config)
        import weakref
        self.transformer._owner = weakref.ref(self)
        self.lm_head = nn.Linear(config.n_embd, config.vocab_size, bias=False)
        self.num_emotions = NUM_EMOTIONS
        self.emotion_head = nn.Linear(config.n_embd, self.num_emotions, bias=False)
        # A3 extension (absent from the reference): with
        # config.ergm_visual_dim / ergm_audio_dim set, `imgs` / `auds` are raw feature SEQUENCES
        # [B, T, D] that are mean-pooled (feature_extraction.py:63,69) and projected D -> n_embd
        # on the device.  Without them the reference layout (pooled n_embd-wide vectors) applies
        # and state_dict() has exactly the reference's keys.
        vd, ad = getattr(config, "ergm_visual_dim", None), getattr(config, "ergm_audio_dim", None)
        if bool(vd) != bool(ad):
            raise ValueError("set both config.ergm_visual_dim and config.ergm_audio_dim (model.py:495-498 fuses both)")
        if vd:
            self.visual_proj = nn.Linear(int(vd), config.n_embd)
            self.audio_proj = nn.Linear(int(ad), config.n_embd)
        self.model_parallel = False
        self.device_map = None
        self.post_init()
        self.lm_head.weight = self.transformer.wte.weight  # model.py:600,605 (tied)
        self._engine = None
        self._dp = None  # set by ergm_b200.parallel.DataParallel
        self.fp32_logits = False
        # label-sparse LM head (engine.forward): with labels, head + CE + their backward run only on the rows whose
        # shifted label is not -100; identical loss and gradients, outputs.logits computed on first access
        self.ergm_sparse_lm_head = True
        # packed variable-length batches: with a right-padded attention_mask, compute only the real
        # positions of every sample (plus position T-1, which the emotion head reads).  Results equal the reference
        # called WITH that attention_mask at every real position and in the emotion head; main.py passes no mask
        # (its pad positions attend to earlier pads), so this is opt-in.
        self.ergm_packed = False
        # "bf16" = bf16 tensor-core operands, fp32 accumulation / residual / statistics (throughput mode);
        # "fp32" = split-operand fp32-accurate products (forward only; logits within 1e-4 of the reference)
        self.ergm_precision = "bf16"

    # -- HF plumbing ---------------------------------------------------------------------
    def get_output_embeddings(self):
        return self.lm_head

    def set_output_embeddings(self, new_embeddings):
        self.lm_head = new_embeddings

    def get_input_embeddings(self):
        return self.transformer.wte

    def set_input_embeddings(self, new_embeddings):
        self.transformer.wte = new_embeddings

    def prepare_inputs_for_generation(self, input_ids, past_key_values=None, inputs_embeds=None, **kwargs):
        """model.py:620-652"""
        token_type_ids = kwargs.get("token_type_ids", None)
        if past_key_values:
            input_ids = input_ids[:, -1].unsqueeze(-1)
            if token_type_ids is not None:
                token_type_ids = token_type_ids[:, -1].unsqueeze(-1)
        attention_mask = kwargs.get("attention_mask", None)
        position_ids = kwargs.get("position_ids", None)
        if attention_mask is not None and position_ids is None:
            position_ids = attention_mask.long().cumsum(-1) - 1
            position_ids.masked_fill_(attention_mask == 0, 1)
            if past_key_values:
                position_ids = position_ids[:, -1].unsqueeze(-1)
        else:
            position_ids = None
        model_inputs = {"input_ids": input_ids}
        model_inputs.update({"past_key_values": past_key_values, "use_cache": kwargs.get("use_cache"),
                             "position_ids": position_ids, "attention_mask": attention_mask,
                             "token_type_ids": token_type_ids})
        return model_inputs

    @staticmethod
    def _reorder_cache(past_key_values, beam_idx):
        """model.py:739-746"""
        return tuple(tuple(ps.index_select(0, beam_idx.to(ps.device)) for ps in layer_past)
                     for layer_past in past_key_values)

    # -- engine --------------------------------------------------------------------------
    @property
    def engine(self):
        if self._engine is None:
            dev = self.transformer.wte.weight.device
            if dev.type != "cuda":
                raise L.ErgmError("ergm_b200 has no CPU path: call .to('cuda') on the model first")
            if self.lm_head.weight is not self.transformer.wte.weight:
                self.lm_head.weight = self.transformer.wte.weight
            with torch.cuda.device(dev):
                self._engine = Engine(self)
        return self._engine

    def _kv_lens_from_mask(self, attention_mask, B, total):
        return _kv_lens_from_mask(attention_mask, B, total)

    def forward(self, input_ids: Optional[torch.LongTensor] = None,
                past_key_values: Optional[Tuple[Tuple[torch.Tensor]]] = None,
                attention_mask: Optional[torch.FloatTensor] = None,
                token_type_ids: Optional[torch.LongTensor] = None,
                position_ids: Optional[torch.LongTensor] = None,
                head_mask: Optional[torch.FloatTensor] = None,
                inputs_embeds: Optional[torch.FloatTensor] = None,
                encoder_hidden_states: Optional[torch.Tensor] = None,
                encoder_attention_mask: Optional[torch.FloatTensor] = None,
                labels: Optional[torch.LongTensor] = None,
                emotion_labels: Optional[torch.LongTensor] = None,
                use_cache: Optional[bool] = None,
                output_attentions: Optional[bool] = None,
                output_hidden_states: Optional[bool] = None,
                return_dict: Optional[bool] = None,
                imgs=None, auds=None, caption_ids: Optional[torch.LongTensor] = None):
        """Same keyword surface as model.py:654-672."""
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if input_ids is None:
            if inputs_embeds is not None:
                raise L.ErgmError("inputs_embeds is not supported: the embedding stage is fused (ergm_embed_fuse_fwd)")
            raise ValueError("You have to specify either input_ids or inputs_embeds")
        _get_head_mask_check(head_mask)
        if output_attentions or output_hidden_states:
            raise L.ErgmError("output_attentions / output_hidden_states are not available: attention probabilities "
                              "are never materialised by the fused kernels")
        return_dict = return_dict if return_dict is not None else getattr(self.config, "return_dict", True)
        use_cache = use_cache if use_cache is not None else self.config.use_cache
        eng = self.engine
        dev = eng.device
        input_ids = input_ids.reshape(-1, input_ids.shape[-1])
        B, T = input_ids.shape

        def to_dev(t):
            if t is None:
                return None
            t = t.to(device=dev, dtype=torch.int64)
            return t.contiguous()

        input_ids = to_dev(input_ids)
        token_type_ids = to_dev(token_type_ids.reshape(B, T)) if token_type_ids is not None else None
        labels = to_dev(labels.reshape(B, T)) if labels is not None else None
        emotion_labels = to_dev(emotion_labels.reshape(-1)) if emotion_labels is not None else None
        caption_ids = to_dev(caption_ids.reshape(B, -1)) if caption_ids is not None else None
        pos = None
        if position_ids is not None:
            position_ids = position_ids.reshape(-1, T)
            if position_ids.shape[0] != 1:
                if not bool((position_ids == position_ids[:1]).all().item()):
                    raise L.ErgmError("per-sample position_ids are not supported (model.py:474-476 uses a shared arange)")
            pos = to_dev(position_ids[0])
        training = self.training and torch.is_grad_enabled()
        with torch.cuda.device(dev):
            if past_key_values is not None:
                return self._forward_with_legacy_past(input_ids, token_type_ids, pos, past_key_values,
                                                      attention_mask, caption_ids, use_cache, return_dict)
            kv_lens = None
            pack = None
            if attention_mask is not None:
                kv_lens = self._kv_lens_from_mask(attention_mask.to(dev), B, T)
                if self.ergm_packed and self.ergm_precision != "fp32":
                    pack = eng.get_pack(B, T).plan(kv_lens.contiguous())
                    kv_lens = None
                    use_cache = False
            save = training and (labels is not None or emotion_labels is not None)
            if self.ergm_precision == "fp32":
                if save:
                    raise L.ErgmError("ergm_precision='fp32' is a forward-only verification mode: call it under "
                                      "torch.no_grad() / model.eval()")
                out = eng.forward_fp32(input_ids, token_type_ids, labels, emotion_labels, imgs, auds, caption_ids,
                                       pos, kv_lens=kv_lens)
                use_cache = False
            else:
                # with labels the LM head runs label-sparse and .logits (all positions) is computed on first access
                out = eng.forward(input_ids, token_type_ids, labels, emotion_labels, imgs, auds, caption_ids, pos,
                                  past_len=0, kv_lens=kv_lens, training=self.training, save=save,
                                  want_logits=labels is None, logits_fp32=self.fp32_logits, pack=pack)
            loss = lm_loss = emo_loss = None
            if labels is not None or emotion_labels is not None:
                if self._dp is not None:
                    self._dp.reduce_loss_sums(out["loss_sums"])
                losses = eng.finalize_loss(out)
                if save:
                    on_bwd = self._dp.backward if self._dp is not None else None
                    loss = _LossFn.apply(self.transformer.wte.weight, eng, losses, on_bwd, out["forward_id"])
                else:
                    loss = losses[0].clone()
                lm_loss = losses[1].clone() if labels is not None else None
                emo_loss = losses[2].clone() if emotion_labels is not None else None
            V = eng.V
            logits_buf = out["logits"]
            logits_src = out.get("logits_src")
            fwd_serial = eng.forward_serial

            def logits_fn():
                buf = logits_buf
                if buf is None:
                    # lazily: ln_f output of THIS forward is valid until the next forward of the model
                    if eng.forward_serial != fwd_serial:
                        raise L.ErgmError("outputs.logits must be read before the next forward of the same model "
                                          "(the label-sparse LM head computes all-position logits on demand)")
                    with torch.cuda.device(dev):
                        buf = eng.full_logits(*logits_src)
                if pack is not None:
                    # packed rows -> the padded [B, T, V] layout the caller expects (pad positions: zeros)
                    n = int(pack.n_rows.item())
                    dst = (pack.row_b[:n].long() * T + pack.row_t[:n].long())
                    full = torch.zeros(B * T, V, dtype=torch.float32, device=buf.device)
                    full[dst] = buf[:n, :V].to(torch.float32)
                    return full.view(B, T, V)
                return buf[:, :V].to(torch.float32).view(B, T, V)

            past_fn = _past_fn(out["kv_present"], B, eng.H, eng.nh, False) if use_cache else None

            emo_logits = out["emotion_logits"].clone()
        from . import ops
        if os.environ.get("ERGM_CHECK_INDICES", "0") == "1":
            ops.check_err_flag(dev)   # synchronous (debugging)
        else:
            ops.poll_err_flag(dev)    # asynchronous: raises at the next forward
        ret = CausalLMOutputWithEmotionClassification(loss=loss, logits_fn=logits_fn, emotion_logits=emo_logits,
                                                      past_fn=past_fn, lm_loss=lm_loss,
                                                      emotion_loss=emo_loss)
        if not return_dict:
            outp = (ret.logits, ret.emotion_logits) + ((ret.past_key_values,) if use_cache else ())
            return ((loss,) + outp) if loss is not None else outp
        return ret

    # -- legacy tuple-cache surface (model.py:228-236, 469-476) ----------------------------
    def _forward_with_legacy_past(self, input_ids, token_type_ids, pos, past_key_values, attention_mask,
                                  caption_ids, use_cache, return_dict):
        from .generation import legacy_cached_forward
        return legacy_cached_forward(self, input_ids, token_type_ids, pos, past_key_values, attention_mask,
                                     caption_ids, use_cache, return_dict)

    @torch.no_grad()
    def generate(self, input_ids, token_type_ids=None, max_new_tokens=64, do_sample=False, top_k=0, top_p=1.0,
                 temperature=1.0, eos_token_id=None, sp2_id=None, imgs=None, auds=None, caption_ids=None,
                 prompt_lens=None, seed=0):
        """KV-cached batched response generation (replaces the full-recompute loop of
        main.py:253-282; transformers 4.26's GenerationMixin.generate no longer exists in 5.x)."""
        from .generation import generate
        return generate(self, input_ids, token_type_ids, max_new_tokens=max_new_tokens, do_sample=do_sample,
                        top_k=top_k, top_p=top_p, temperature=temperature, eos_token_id=eos_token_id,
                        sp2_id=sp2_id, imgs=imgs, auds=auds, caption_ids=caption_ids, prompt_lens=prompt_lens,
                        seed=seed)


__all__ = ["GPT2LMHeadModel", "GPT2Model", "GPT2Block", "GPT2Attention", "GPT2MLP", "GPT2PreTrainedModel",
           "CausalLMOutputWithEmotionClassification", "GPT2Config", "torch", "nn", "F", "math", "os"]

"""Module-level forwards of the exported building blocks (GPT2Attention / GPT2MLP / GPT2Block).

The reference makes these classes callable (model.py:200-251, 262-267, 286-341) and `from model import *`
re-exports them, so code that composes a block by hand keeps working against the B200 build.  They are
INFERENCE surfaces: each call is a short chain of the same C-ABI kernels the fused engine uses (bf16
tensor-core GEMMs with fp32 accumulation, fused attention, LayerNorm), no autograd graph is recorded —
training goes through GPT2LMHeadModel.forward, whose backward is hand-written (ergm_b200.engine).
Dropout modules of the reference (attn / resid) are identity here, exactly as in its eval mode.
"""
import torch

from . import _lib as L
from . import ops

K_MAJOR, MN_MAJOR = L.ERGM_MAJOR_K, L.ERGM_MAJOR_MN


def _need_cuda(t, what):
    if t.device.type != "cuda":
        raise L.ErgmError("%s: ergm_b200 has no CPU path (tensor on %s)" % (what, t.device))


def _shadow(module, name):
    """bf16 copy of an fp32 parameter for the tensor cores, cached per (storage, version)."""
    p = getattr(module, name)
    cache = module.__dict__.setdefault("_ergm_shadow", {})
    key = (p.data_ptr(), p._version, tuple(p.shape))
    hit = cache.get(name)
    if hit is not None and hit[0] == key:
        return hit[1]
    src = p.detach()
    if src.dtype != torch.float32 or not src.is_contiguous():
        src = src.float().contiguous()
    out = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    ops.cast_f32_bf16(src, out)
    cache[name] = (key, out)
    return out


def _rows_bf16(x):
    """[..., K] any float dtype -> contiguous bf16 [rows, K] (the GEMM A operand) through ergm_cast_f32_bf16."""
    x2 = x.reshape(-1, x.shape[-1])
    if x2.dtype == torch.bfloat16 and x2.is_contiguous():
        return x2
    x2 = x2.float().contiguous()
    out = torch.empty(x2.shape, dtype=torch.bfloat16, device=x2.device)
    ops.cast_f32_bf16(x2, out)
    return out


def conv1d(module, x2_bf16, out_dtype=torch.float32, gelu=False, residual=None):
    """transformers Conv1D (addmm(bias, x, W[K, N])) on [rows, K] bf16 rows -> [rows, N]; `residual` (fp32
    [rows, N]) is added in the GEMM epilogue (the block's `attn_output + residual`, model.py:309,328,334)."""
    w = _shadow(module, "weight")
    K, N = w.shape
    rows = x2_bf16.shape[0]
    out = torch.empty(rows, N, dtype=out_dtype, device=x2_bf16.device)
    ops.gemm(x2_bf16, w, out, M=rows, N=N, K=K, a_major=K_MAJOR, b_major=MN_MAJOR, bias=module.bias.detach(),
             residual=residual, epilogue=L.EPI_GELU if gelu else 0)
    return out


def layer_norm_rows(ln, x, want_bf16=True):
    """nn.LayerNorm over the last dim of an fp32 [rows, H] matrix -> bf16 (GEMM operand) or fp32."""
    rows, H = x.shape
    y = torch.empty(rows, H, dtype=torch.bfloat16 if want_bf16 else torch.float32, device=x.device)
    ops.ln_fwd(x, ln.weight.detach(), ln.bias.detach(), y if want_bf16 else None, None if want_bf16 else y, None, None,
               ln.eps)
    return y


@torch.no_grad()
def mlp_forward(self, hidden_states, _residual=None):
    """GPT2MLP.forward, model.py:262-267: c_proj(gelu_new(c_fc(x))) (dropout = identity)."""
    _need_cuda(hidden_states, "GPT2MLP.forward")
    with torch.cuda.device(hidden_states.device):
        g = conv1d(self.c_fc, _rows_bf16(hidden_states), out_dtype=torch.bfloat16, gelu=True)
        y = conv1d(self.c_proj, g, residual=_residual)
    y = y.view(hidden_states.shape[:-1] + (y.shape[-1],))
    return y if _residual is not None else y.to(hidden_states.dtype)


def _check_masks(attention_mask, head_mask, output_attentions):
    if head_mask is not None:
        raise L.ErgmError("head_mask is not supported by the fused attention kernels")
    if output_attentions:
        raise L.ErgmError("output_attentions: attention probabilities are never materialised by the fused kernels")
    if attention_mask is not None and bool((attention_mask != 0).any()):
        raise L.ErgmError("module-level forwards take no additive attention_mask (pass right-padded batches through "
                          "GPT2LMHeadModel / GPT2Model, which turn the mask into per-sequence key lengths)")


def _split_heads(t2, B, T, nh):
    return t2.view(B, T, nh, 64).permute(0, 2, 1, 3)


@torch.no_grad()
def attention_forward(self, hidden_states, layer_past=None, attention_mask=None, head_mask=None,
                      encoder_hidden_states=None, encoder_attention_mask=None, use_cache=False,
                      output_attentions=False, _residual=None):
    """GPT2Attention.forward, model.py:200-251.  Returns (attn_output, present) like the reference: present =
    (key, value) as [B, nh, ctx, 64] when use_cache, else None."""
    _need_cuda(hidden_states, "GPT2Attention.forward")
    _check_masks(attention_mask, head_mask, output_attentions)
    _check_masks(encoder_attention_mask, None, False)
    if self.head_dim != 64:
        raise L.ErgmError("ergm_b200 attention kernels need head_dim == 64 (got %d)" % self.head_dim)
    B, T, H = hidden_states.shape
    nh = self.num_heads
    dev = hidden_states.device
    bf16 = torch.bfloat16
    with torch.cuda.device(dev):
        x = _rows_bf16(hidden_states)
        ctx = torch.empty(B * T, H, dtype=bf16, device=dev)
        lse = torch.empty(B, nh, T, dtype=torch.float32, device=dev)
        if encoder_hidden_states is not None:
            if not hasattr(self, "q_attn"):
                raise ValueError(
                    "If class is used as cross attention, the weights `q_attn` have to be defined. "
                    "Please make sure to instantiate class with `GPT2Attention(..., is_cross_attention=True)`.")
            Tk = encoder_hidden_states.shape[1]
            q = conv1d(self.q_attn, x, out_dtype=bf16)
            kv = conv1d(self.c_attn, _rows_bf16(encoder_hidden_states), out_dtype=bf16)  # [B*Tk, 2H]
            if layer_past is not None:
                raise L.ErgmError("layer_past with cross-attention is not meaningful (model.py:211-220)")
            ops.attn_fwd(q, kv, kv, ctx, lse, B=B, nh=nh, Tq=T, Tk=Tk, q_col0=0, k_col0=0, v_col0=H, causal=False)
            k2, v2 = kv[:, :H], kv[:, H:]
        else:
            qkv = conv1d(self.c_attn, x, out_dtype=bf16)  # [B*T, 3H]
            Tk = T
            if layer_past is not None:
                pk, pv = layer_past
                past = pk.shape[-2]
                pkv = torch.cat([pk.permute(0, 2, 1, 3).reshape(B, past, H), pv.permute(0, 2, 1, 3).reshape(B, past, H)],
                                dim=-1).to(bf16)
                kvf = torch.cat([pkv, qkv.view(B, T, 3 * H)[:, :, H:]], dim=1).contiguous()  # model.py:228-231
                Tk = past + T
                kv2d = kvf.view(B * Tk, 2 * H)
                ops.attn_fwd(qkv, kv2d, kv2d, ctx, lse, B=B, nh=nh, Tq=T, Tk=Tk, q_col0=0, k_col0=0, v_col0=H,
                             causal=True, causal_off=past)
                k2, v2 = kv2d[:, :H], kv2d[:, H:]
            else:
                ops.attn_fwd(qkv, qkv, qkv, ctx, lse, B=B, nh=nh, Tq=T, Tk=T, q_col0=0, k_col0=H, v_col0=2 * H, causal=True)
                k2, v2 = qkv[:, H:2 * H], qkv[:, 2 * H:]
        out = conv1d(self.c_proj, ctx, residual=_residual).view(B, T, H)
        if _residual is None:
            out = out.to(hidden_states.dtype)
        present = None
        if use_cache is True:
            present = (_split_heads(k2, B, Tk, nh).float(), _split_heads(v2, B, Tk, nh).float())
    return out, present


@torch.no_grad()
def block_forward(self, hidden_states, layer_past=None, attention_mask=None, head_mask=None,
                  encoder_hidden_states=None, encoder_attention_mask=None, use_cache=False, output_attentions=False):
    """GPT2Block.forward, model.py:286-341: x + attn(ln_1 x); (+ cross(ln_cross x, enc)); + mlp(ln_2 x).
    Returns (hidden_states, present) when use_cache else (hidden_states,)."""
    _need_cuda(hidden_states, "GPT2Block.forward")
    B, T, H = hidden_states.shape
    with torch.cuda.device(hidden_states.device):
        # the three residual adds ride in the c_proj GEMM epilogues, the residual stream stays fp32
        x = hidden_states.reshape(B * T, H).float().contiguous()
        x, present = attention_forward(self.attn, layer_norm_rows(self.ln_1, x).view(B, T, H), layer_past=layer_past,
                                       attention_mask=attention_mask, head_mask=head_mask, use_cache=use_cache,
                                       output_attentions=output_attentions, _residual=x)
        x = x.view(B * T, H)
        if encoder_hidden_states is not None:
            x, _ = attention_forward(self.crossattention, layer_norm_rows(self.ln_cross_attn, x).view(B, T, H),
                                     head_mask=head_mask, encoder_hidden_states=encoder_hidden_states,
                                     encoder_attention_mask=encoder_attention_mask, output_attentions=output_attentions,
                                     _residual=x)
            x = x.view(B * T, H)
        x = mlp_forward(self.mlp, layer_norm_rows(self.ln_2, x).view(B, T, H), _residual=x)
        x = x.view(B, T, H).to(hidden_states.dtype)
    return (x, present) if use_cache else (x,)

"""Flat parameter / gradient / bf16-shadow storage for the ERGM model.

All nn.Parameters of the model are re-bound to views of ONE flat fp32 buffer (so the fused
AdamW, the bf16 weight-shadow refresh and the gradient all-reduce are single launches over
contiguous memory), with a same-layout fp32 gradient buffer and a bf16 shadow that the
tensor-core GEMMs read.  state_dict() keys and shapes are untouched: the reference's
checkpoints (main.py:98-110,184-196) load 1:1.
"""
import torch

from . import ops

ALIGN = 64  # elements; keeps every tensor 16-byte aligned in both the fp32 and bf16 buffers


class ParamStore:
    def __init__(self, module, prefix=""):
        self.module = module
        self.prefix = prefix  # a stand-alone GPT2Model is stored under the names it has inside GPT2LMHeadModel
        self.entries = {}   # name -> (offset, numel, shape)
        self.flat = self.grad = self.shadow = None
        self._ptrs = None
        self._versions = None
        self.shadow_fresh = False
        # bumped whenever the parameter VALUES may have changed (fused optimiser step, detected in-place edits,
        # rebuild): derived weight formats (decode slabs, fp32-mode split operands) key their caches on it,
        # because the fused AdamW writes the flat buffer directly and never touches torch's version counters
        self.weights_epoch = 0
        self.build()

    # ------------------------------------------------------------------
    def _unique_named_params(self):
        seen = set()
        for name, p in self.module.named_parameters(remove_duplicate=False):
            if id(p) in seen:
                continue
            seen.add(id(p))
            yield self.prefix + name, p

    def build(self):
        params = list(self._unique_named_params())
        if not params:
            raise RuntimeError("model has no parameters")
        device = params[0][1].device
        if device.type != "cuda":
            raise RuntimeError("ergm_b200 runs on CUDA only (no CPU fallback): move the model to a B200 first")
        off = 0
        self.entries = {}
        for name, p in params:
            if p.dtype != torch.float32:
                raise RuntimeError("parameter %s must be fp32 (master weights), got %s" % (name, p.dtype))
            self.entries[name] = (off, p.numel(), tuple(p.shape))
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.total = off
        flat = torch.zeros(off, dtype=torch.float32, device=device)
        grad = torch.zeros(off, dtype=torch.float32, device=device)
        had_grad = False
        with torch.no_grad():
            for name, p in params:
                o, n, shape = self.entries[name]
                flat[o:o + n].copy_(p.data.reshape(-1))
                if p.grad is not None:
                    grad[o:o + n].copy_(p.grad.reshape(-1))
                    had_grad = True
                p.data = flat[o:o + n].view(shape)
                if p.grad is not None:
                    p.grad = grad[o:o + n].view(shape)
        self.flat, self.grad = flat, grad
        self.shadow = torch.empty(off, dtype=torch.bfloat16, device=device)
        self.params = dict(params)
        self._ptrs = [p.data_ptr() for _, p in params]
        self._plist = [p for _, p in params]
        self._versions = None
        self.shadow_fresh = False
        self.weights_epoch = getattr(self, "weights_epoch", 0) + 1
        self.device = device
        self.had_grad = had_grad

    def valid(self):
        """False when some parameter was re-bound behind our back (.to(), resize_token_embeddings,
        load_state_dict(assign=True) ...)."""
        plist = [p for _, p in self._unique_named_params()]
        if len(plist) != len(self._plist):
            return False
        for p, q, ptr in zip(plist, self._plist, self._ptrs):
            if p is not q or p.data_ptr() != ptr:
                return False
        return True

    def ensure(self):
        if not self.valid():
            self.build()

    # ------------------------------------------------------------------
    def view(self, name):
        o, n, shape = self.entries[name]
        return self.flat[o:o + n].view(shape)

    def grad_view(self, name):
        o, n, shape = self.entries[name]
        return self.grad[o:o + n].view(shape)

    def shadow_view(self, name):
        o, n, shape = self.entries[name]
        return self.shadow[o:o + n].view(shape)

    def refresh_shadow(self, force=False):
        """bf16(weights) for the GEMMs; re-cast only when a parameter changed (tensor version
        counters) unless the fused optimiser already wrote it."""
        versions = sum(p._version for p in self._plist)
        if force or not self.shadow_fresh or versions != self._versions:
            ops.cast_f32_bf16(self.flat, self.shadow)
            self._versions = versions
            self.shadow_fresh = True
            self.weights_epoch += 1

    def mark_shadow_fresh(self):
        """Called by the fused optimiser: it rewrote the fp32 weights AND their bf16 shadow."""
        self._versions = sum(p._version for p in self._plist)
        self.shadow_fresh = True
        self.weights_epoch += 1

    def grads_live(self):
        """True when the caller kept gradients (no zero_grad since the last backward): the next
        backward must accumulate instead of overwrite."""
        return all(p.grad is not None for p in self._plist) or (
            any(p.grad is not None for p in self._plist) and all(
                p.grad is not None for n, p in self.params.items()
                if "crossattention." not in n and "ln_cross_attn." not in n))

    def bind_grads(self, skip_substrings=()):
        """Points every p.grad at its slice of the flat gradient buffer.  Parameters whose name
        contains one of `skip_substrings` took no part in this backward: their .grad is left
        untouched (None after zero_grad), like autograd would."""
        for name, p in self.params.items():
            if skip_substrings and any(s in name for s in skip_substrings):
                continue
            o, n, shape = self.entries[name]
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + n].view(shape)

"""Fused flat AdamW + the polynomial-decay schedule of main.py:93-95.

Same arithmetic as torch.optim.AdamW (lr, betas (0.9, 0.999), eps 1e-8, weight_decay 0.01 on
ALL parameters including LayerNorm / bias, as main.py:68 configures it), applied by ONE kernel
over the flat parameter buffer; the kernel also writes the bf16 weight shadow the GEMMs read,
so no separate cast pass is needed after a step.  Hyper-parameters live in a small device
buffer refreshed by an async H2D copy from pinned memory, which keeps the step CUDA-graph
capturable while lr / bias corrections change every step.
"""
import torch

from . import ops


class FusedAdamW:
    def __init__(self, model, lr=2e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        self.model = model
        self.lr = lr
        self.betas = betas
        self.eps = eps
        self.weight_decay = weight_decay
        self.step_count = 0
        self.state = None
        self.param_groups = [{"lr": lr, "params": list(model.parameters())}]

    def _ensure(self):
        eng = self.model.engine
        eng.ensure_params()
        st = eng.store
        if self.state is None or self.state["n"] != st.total or self.state["flat_ptr"] != st.flat.data_ptr():
            dev = st.device
            self.state = dict(n=st.total, flat_ptr=st.flat.data_ptr(),
                              m=torch.zeros(st.total, device=dev), v=torch.zeros(st.total, device=dev),
                              hyper=torch.zeros(8, device=dev),
                              hyper_host=torch.zeros(8).pin_memory())
        return eng, st

    def set_lr(self, lr):
        self.lr = lr
        self.param_groups[0]["lr"] = lr

    def zero_grad(self, set_to_none=True):
        for p in self.model.parameters():
            p.grad = None

    def load_hyper(self):
        """Host part of a step: bump the step count and stage the hyper-parameters."""
        self._ensure()
        self.step_count += 1
        self.lr = self.param_groups[0]["lr"]
        b1, b2 = self.betas
        h = self.state["hyper_host"]
        h[0], h[1], h[2], h[3], h[4] = self.lr, b1, b2, self.eps, self.weight_decay
        h[5], h[6] = 1.0 - b1 ** self.step_count, 1.0 - b2 ** self.step_count
        self.state["hyper"].copy_(h, non_blocking=True)

    def apply(self, grad_scale=None, lo=0, hi=None, last=True):
        """Device part of a step (graph-capturable): one kernel over the flat buffers, or over the
        element range [lo, hi) of them (DataParallel updates the layer parameters while the
        embedding bucket is still being all-reduced).  `last`: this call completes the step."""
        eng, st = self._ensure()
        hi = st.total if hi is None else hi
        if hi > lo:
            ops.adamw_flat(st.flat[lo:hi], st.grad[lo:hi], self.state["m"][lo:hi], self.state["v"][lo:hi],
                           st.shadow[lo:hi], self.state["hyper"], grad_scale)
        if last:
            st.mark_shadow_fresh()

    @torch.no_grad()
    def step(self):
        self.load_hyper()
        self.apply()

    def state_dict(self):
        return dict(step=self.step_count, m=None if self.state is None else self.state["m"].clone(),
                    v=None if self.state is None else self.state["v"].clone(), lr=self.lr)

    def load_state_dict(self, sd):
        self._ensure()
        self.step_count = sd["step"]
        if sd["m"] is not None:
            self.state["m"].copy_(sd["m"])
            self.state["v"].copy_(sd["v"])


class PolynomialDecaySchedule:
    """transformers.get_polynomial_decay_schedule_with_warmup(power=2, lr_end=1e-7) — main.py:93-95."""

    def __init__(self, optimizer, num_warmup_steps, num_training_steps, lr_end=1e-7, power=2.0):
        self.opt = optimizer
        self.base_lr = optimizer.lr
        self.warm, self.total, self.lr_end, self.power = num_warmup_steps, num_training_steps, lr_end, power
        self.last_step = 0
        self.opt.set_lr(self.lr_at(0))

    def lr_at(self, step):
        if step < self.warm:
            return self.base_lr * float(step) / float(max(1, self.warm))
        if step > self.total:
            return self.lr_end
        rng = self.base_lr - self.lr_end
        pct = 1 - (step - self.warm) / (self.total - self.warm)
        return rng * pct ** self.power + self.lr_end

    def step(self):
        self.last_step += 1
        self.opt.set_lr(self.lr_at(self.last_step))

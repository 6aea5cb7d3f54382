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
    """torch.optim.AdamW semantics over the flat parameter store, including its treatment of parameters
    that took no part in the backward: a parameter whose .grad is None (the cross-attention / ln_cross_attn
    tensors when the model runs without caption_ids, which is how main.py:147 calls it) is skipped - no
    weight decay, no moment update, no step count - and has no entry in state_dict()."""

    MAX_GROUPS = 4  # distinct per-parameter step counts that can be live at once (normally 1)

    def __init__(self, model, lr=2e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        self.model = model
        self.lr = lr
        self.betas = betas
        self.eps = eps
        self.weight_decay = weight_decay
        self.step_count = 0          # number of step() calls (the schedule's clock)
        self.param_steps = {}        # parameter name -> number of updates it received
        self.state = None
        self.param_groups = [{"lr": lr, "params": list(model.parameters())}]
        self._plan = None            # [(lo, hi, hyper_row)] of the step being applied

    def _ensure(self):
        eng = self.model.engine
        eng.ensure_params()
        st = eng.store
        if self.state is None or self.state["n"] != st.total or self.state["flat_ptr"] != st.flat.data_ptr():
            dev = st.device
            self.state = dict(n=st.total, flat_ptr=st.flat.data_ptr(),
                              m=torch.zeros(st.total, device=dev), v=torch.zeros(st.total, device=dev),
                              hyper=torch.zeros(self.MAX_GROUPS, 8, device=dev),
                              hyper_host=torch.zeros(self.MAX_GROUPS, 8).pin_memory())
        return eng, st

    def set_lr(self, lr):
        self.lr = lr
        self.param_groups[0]["lr"] = lr

    def zero_grad(self, set_to_none=True):
        for p in self.model.parameters():
            p.grad = None

    def _active_names(self, st, skip):
        """skip=None: torch's rule (p.grad is not None).  skip=tuple of substrings: the caller knows which
        parameters the coming backward will not touch (GraphedTrainStep stages the hyper-parameters before the
        backward has run)."""
        if skip is None:
            return [n for n, p in st.params.items() if p.grad is not None]
        return [n for n in st.params if not any(s in n for s in skip)]

    def load_hyper(self, skip=None):
        """Host part of a step: bump the step counts of the participating parameters, merge them into
        contiguous ranges of the flat buffer and stage one hyper-parameter row per distinct step count."""
        eng, st = self._ensure()
        self.step_count += 1
        self.lr = self.param_groups[0]["lr"]
        b1, b2 = self.betas
        names = self._active_names(st, skip)
        groups = {}   # step -> hyper row
        plan = []
        h = self.state["hyper_host"]
        for n in names:
            k = self.param_steps.get(n, 0) + 1
            self.param_steps[n] = k
            row = groups.get(k)
            if row is None:
                row = len(groups)
                if row >= self.MAX_GROUPS:
                    raise RuntimeError("FusedAdamW: more than %d distinct per-parameter step counts are live"
                                       % self.MAX_GROUPS)
                groups[k] = row
                h[row, 0], h[row, 1], h[row, 2], h[row, 3], h[row, 4] = self.lr, b1, b2, self.eps, self.weight_decay
                h[row, 5], h[row, 6] = 1.0 - b1 ** k, 1.0 - b2 ** k
            o, numel, _ = st.entries[n]
            hi = o + (numel + 63) // 64 * 64   # the alignment padding is zero in every buffer: harmless to include
            if plan and plan[-1][1] == o and plan[-1][2] == row:
                plan[-1][1] = hi
            else:
                plan.append([o, hi, row])
        self._plan = plan
        self.state["hyper"].copy_(h, non_blocking=True)

    def apply(self, grad_scale=None, lo=0, hi=None, last=True, grads=None):
        """Device part of a step (graph-capturable): one kernel per contiguous range of participating
        parameters (ONE launch when every parameter has a gradient), restricted to the element range [lo, hi)
        (DataParallel updates the layer parameters while the embedding bucket is still being all-reduced).
        `last`: this call completes the step.  `grads`: a same-layout gradient buffer to read instead of the flat
        fp32 one (DataParallel's bf16 buckets)."""
        eng, st = self._ensure()
        hi = st.total if hi is None else hi
        if self._plan is None:
            raise RuntimeError("FusedAdamW.apply() without load_hyper()")
        g = st.grad if grads is None else grads
        for a, b, row in self._plan:
            a, b = max(a, lo), min(b, hi)
            if b > a:
                ops.adamw_flat(st.flat[a:b], g[a:b], self.state["m"][a:b], self.state["v"][a:b],
                               st.shadow[a:b], self.state["hyper"][row], grad_scale)
        if last:
            st.mark_shadow_fresh()

    @torch.no_grad()
    def step(self):
        self.load_hyper()
        self.apply()

    # -- checkpoint surface: the layout of torch.optim.AdamW.state_dict(), which is what the reference stores in
    #    ckpt["optim_state_dict"] (main.py:103-110, 184-196), so its checkpoints resume here and vice versa --------
    def _param_slices(self):
        eng, st = self._ensure()
        out = []
        for name, p in st.params.items():  # same order as model.parameters() (tied lm_head counted once)
            o, n, shape = st.entries[name]
            out.append((name, o, n, shape))
        return out

    def state_dict(self):
        slices = self._param_slices()
        state = {}
        for i, (name, o, n, shape) in enumerate(slices):
            k = self.param_steps.get(name, 0)
            if k > 0:  # torch keeps state only for parameters that were ever updated
                state[i] = {"step": torch.tensor(float(k)),
                            "exp_avg": self.state["m"][o:o + n].view(shape).clone(),
                            "exp_avg_sq": self.state["v"][o:o + n].view(shape).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "params": list(range(len(slices)))}
        if "initial_lr" in self.param_groups[0]:
            group["initial_lr"] = self.param_groups[0]["initial_lr"]
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        slices = self._param_slices()
        if "state" not in sd:  # legacy flat layout written by earlier versions of this class
            self.step_count = sd["step"]
            self.param_steps = {name: sd["step"] for name, _, _, _ in slices}
            if sd["m"] is not None:
                self.state["m"].copy_(sd["m"])
                self.state["v"].copy_(sd["v"])
            return
        groups = sd["param_groups"]
        ids = [i for g in groups for i in g["params"]]
        if len(ids) != len(slices):
            raise ValueError("optimizer state has %d parameters, the model has %d" % (len(ids), len(slices)))
        g0 = groups[0]
        self.lr, self.betas, self.eps = g0["lr"], tuple(g0["betas"]), g0["eps"]
        self.weight_decay = g0["weight_decay"]
        self.param_groups[0]["lr"] = self.lr
        if "initial_lr" in g0:
            self.param_groups[0]["initial_lr"] = g0["initial_lr"]
        self.state["m"].zero_()
        self.state["v"].zero_()
        self.param_steps = {}
        for pos, pid in enumerate(ids):
            ent = sd["state"].get(pid)
            if ent is None:
                continue
            name, o, n, shape = slices[pos]
            if tuple(ent["exp_avg"].shape) != tuple(shape):
                raise ValueError("optimizer state of parameter %d has shape %s, expected %s"
                                 % (pid, tuple(ent["exp_avg"].shape), tuple(shape)))
            self.state["m"][o:o + n].copy_(ent["exp_avg"].reshape(-1))
            self.state["v"][o:o + n].copy_(ent["exp_avg_sq"].reshape(-1))
            self.param_steps[name] = int(float(ent["step"]))
        self.step_count = max(self.param_steps.values()) if self.param_steps else 0


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

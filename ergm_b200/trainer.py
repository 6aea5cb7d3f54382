"""Graph-captured training step: the public entry point for the throughput path.

    step = GraphedTrainStep(model, optimizer)          # optimizer: ergm_b200.optim.FusedAdamW
    loss = step(batch_on_pinned_host)                  # python float (D2H of the loss)

One step = H2D copy of the batch from pinned host memory into static device buffers, then ONE
CUDA-graph replay of  forward -> hand-written backward (-> bucketed NCCL all-reduce under
DataParallel) -> fused AdamW (+ bf16 shadow refresh), then the D2H read of the loss.  The graph
is captured on the first call for a given batch shape; the arithmetic is exactly what
model(**batch).loss.backward(); optimizer.step() runs eagerly (same C-ABI calls, same order).
Learning-rate / bias-correction values and the dropout step counter live in device memory, so
the replayed kernels see fresh values every step.
"""
import os

import torch

from . import ops

# seq_lens (int32 [B], real tokens of every right-padded sample): selects the packed variable-length path
_KEYS = ("input_ids", "token_type_ids", "labels", "emotion_labels", "caption_ids", "imgs", "auds", "seq_lens")


class GraphedTrainStep:
    def __init__(self, model, optimizer, dp=None, use_graph=True):
        self.model = model
        self.opt = optimizer
        self.dp = dp if dp is not None else getattr(model, "_dp", None)
        # Under DataParallel the bucketed NCCL all-reduces are captured into the same graph (side stream
        # fork/join per bucket, see parallel.py); ERGM_DP_GRAPH=0 falls back to eager launches.  Call
        # close() before dist.destroy_process_group(): a live graph pins NCCL resources.
        if self.dp is not None and getattr(self.dp, "world", 1) > 1 and os.environ.get("ERGM_DP_GRAPH", "1") == "0":
            use_graph = False
        self.use_graph = use_graph
        # A3 extension: imgs / auds are raw feature sequences [B, T, D] (pooled + projected on the device)
        self.seq_features = hasattr(model, "visual_proj")
        self.graphs = {}
        self.graph_plans = {}
        self.static = {}
        eng = model.engine
        self.eng = eng
        self.rng_step = torch.zeros(1, dtype=torch.int64, device=eng.device)
        eng.set_rng_step_tensor(self.rng_step)
        self.one = torch.ones(1, dtype=torch.float32, device=eng.device)
        self.rows_hint = 0   # typical packed row count (from the first packed batch): steers GEMM tile shapes only
        self.loss_host = torch.zeros(5, dtype=torch.float32).pin_memory()
        self.err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.h2d_bytes = 0
        self.launches_per_step = 0

    # the device work of one step (identical in eager and captured mode)
    def _device_step(self, b):
        eng, model = self.eng, self.model
        ops.reset_launch_count()
        pack = None
        if b.get("seq_lens") is not None:
            Bq, Tq = b["input_ids"].shape
            pack = eng.get_pack(Bq, Tq).plan(b["seq_lens"])   # one small kernel: part of the captured step
            pack.rows_hint = self.rows_hint
        out = eng.forward(b["input_ids"], b.get("token_type_ids"), b.get("labels"), b.get("emotion_labels"),
                          b.get("imgs"), b.get("auds"), b.get("caption_ids"), None, training=model.training,
                          save=True, want_logits=False, logits_fp32=model.fp32_logits, pack=pack)
        if self.dp is not None:
            self.dp.reduce_loss_sums(out["loss_sums"])
        losses = eng.finalize_loss(out)
        if self.dp is not None and getattr(self.dp, "world", 1) > 1:
            # AdamW on the layer / head parameters runs while the embedding bucket (wte + wpe, 25 % of the
            # bytes, final only after the embedding backward) is still being all-reduced
            # (wte + wpe, 25 % of the bytes, and the A3 projections: final only after the embedding backward)
            self.dp.backward(self.one, False, defer_last=True)
            gb = self.dp.grad_bf16   # bf16 buckets (grad_dtype="bf16"): the optimiser reads them directly
            for lo, hi in self.dp.early_ranges():
                self.opt.apply(lo=lo, hi=hi, last=False, grads=gb)
            self.dp.finish()
            late = self.dp.deferred_ranges()
            for i, (lo, hi) in enumerate(late):
                self.opt.apply(lo=lo, hi=hi, last=(i == len(late) - 1), grads=gb)
        else:
            if self.dp is not None:
                self.dp.backward(self.one, False)
            else:
                eng.backward(self.one, accumulate=False)
            self.opt.apply()
        L = ops.launch_count()
        ops.rng_step_advance(self.rng_step, 1)
        self.launches_per_step = L + 1
        return losses

    def _stage(self, key, batch):
        st = {}
        nbytes = 0
        for k in _KEYS:
            v = batch.get(k)
            if v is None:
                continue
            if k == "imgs" and v.dim() == 3 and not self.seq_features:
                v = v[:, 0]
            dt = torch.float32 if k in ("imgs", "auds") else (torch.int32 if k == "seq_lens" else torch.int64)
            st[k] = torch.empty(tuple(v.shape), dtype=dt, device=self.eng.device)
            nbytes += st[k].numel() * st[k].element_size()
        self.static[key] = st
        self.h2d_bytes = nbytes
        return st

    def _key(self, batch):
        return tuple((k, tuple(batch[k].shape)) for k in _KEYS if batch.get(k) is not None)

    def copy_in(self, batch):
        key = self._key(batch)
        st = self.static.get(key) or self._stage(key, batch)
        for k, dst in st.items():
            src = batch[k]
            if k == "imgs" and src.dim() == 3 and not self.seq_features:
                src = src[:, 0]
            dst.copy_(src, non_blocking=True)
        return key, st

    def run_device(self, key, st):
        """Forward + backward + optimizer on the static buffers (graph replay after capture)."""
        # the optimiser needs to know BEFORE the backward which parameters it will touch (torch's "grad is None
        # -> skip" rule): without captions the cross-attention tensors take no part (model.py:311-329)
        self.opt.load_hyper(skip=() if "caption_ids" in st else ("crossattention.", "ln_cross_attn."))
        if not self.use_graph:
            self.losses = self._device_step(st)
            return
        g = self.graphs.get(key)
        if g is not None and self.graph_plans.get(key) != self.opt._plan:
            g = None  # the per-parameter step counts regrouped (mode switch mid-training): capture again
        if g is None:
            # warm-up eagerly once (allocates workspaces, NCCL channels), then capture
            self.losses = self._device_step(st)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()  # capture executes nothing: the hyper buffer is re-staged before each replay
            # thread_local: the NCCL watchdog thread polls its events while we capture
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self.losses = self._device_step(st)
            self.graphs[key] = g
            self.graph_plans[key] = [list(r) for r in self.opt._plan]
            return
        g.replay()
        # the replayed AdamW rewrote the weights and their bf16 shadow; the Python-side bookkeeping of the
        # captured call did not run again, so tell the parameter store explicitly
        self.eng.store.mark_shadow_fresh()

    def close(self):
        """Drops the captured graphs (needed before tearing down the process group) and unregisters this trainer's
        dropout step counter from the library."""
        torch.cuda.synchronize()
        self.graphs.clear()
        torch.cuda.synchronize()
        if self.eng.rng_step is self.rng_step:
            self.eng.set_rng_step_tensor(None)

    def __del__(self):
        try:
            if self.eng.rng_step is self.rng_step:
                self.eng.set_rng_step_tensor(None)
        except Exception:
            pass

    def __call__(self, batch):
        if batch.get("seq_lens") is not None and not self.rows_hint:
            lens = batch["seq_lens"].to(torch.int64)
            T = batch["input_ids"].shape[1]
            self.rows_hint = int(lens.clamp(max=T).sum() + (lens < T).sum())
        key, st = self.copy_in(batch)
        self.run_device(key, st)
        self.loss_host.copy_(self.losses, non_blocking=True)
        flag = ops.err_flag(self.eng.device)
        self.err_host.copy_(flag, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if int(self.err_host[0]) != 0:
            flag.zero_()
            raise IndexError("ergm_b200: index out of range in input_ids / token_type_ids / caption_ids / labels "
                             "(the reference raises IndexError there)")
        return float(self.loss_host[0])

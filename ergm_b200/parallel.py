"""Data-parallel training and per-GPU-sharded generation.

The reference is single-process / single-GPU (main.py:40-41); the path shards naturally over
samples, so the build adds exactly one exchange step per iteration: a bucketed all-reduce (sum)
of the flat fp32 gradient buffer over NCCL / NVLink, issued layer-group by layer-group while the
hand-written backward is still running, plus one 4-float all-reduce of
[lm_loss_sum, n_valid_tokens, emotion_loss_sum, n_samples] right after the forward so that every
rank scales its backward by 1 / (GLOBAL count): the N-GPU step then equals the single-GPU
reference on the concatenated batch (per-rank means would not, because the number of non -100
labels differs per rank).

One process per GPU (torchrun), backend nccl on GPUs and gloo for the CPU tests of the host
logic in this file.
"""
import torch
import torch.distributed as dist


def plan_buckets(entries, n_layer, bucket_bytes):
    """Splits the flat gradient buffer into all-reduce ranges in the order the backward pass
    completes them.  `entries`: {param name: (offset, numel, shape)} in flat-buffer order
    (wte, wpe, h.0 ... h.L-1, ln_f, emotion_head).  Returns a list of
    (trigger_layer, lo, hi): after the backward of `trigger_layer` (L-1 ... 0, or -1 for the
    embedding stage) the fp32 range [lo, hi) is final."""
    total_hi = max(o + ((n + 63) // 64) * 64 for o, n, _ in entries.values())
    layer_lo = [entries["transformer.h.%d.ln_1.weight" % l][0] for l in range(n_layer)]
    # Parameters whose gradients are only written by the EMBEDDING stage, at the very end of the backward
    # (engine.backward: embed_bwd, then the A3 projection wgrads): wte / wpe at the front of the buffer and,
    # when the model has them, visual_proj / audio_proj, registered after emotion_head at its tail.  They must
    # never ride in a bucket that a layer triggers - that bucket would be reduced before they exist.
    late = [entries[n][0] for n in entries if n.startswith(("visual_proj.", "audio_proj."))]
    head_hi = min(late) if late else total_hi
    if late and any(o >= head_hi and not n.startswith(("visual_proj.", "audio_proj.")) for n, (o, _, _) in entries.items()):
        raise RuntimeError("plan_buckets: a non-projection parameter is registered after the A3 projections")
    buckets = []
    hi = head_hi  # the head parameters (ln_f, emotion_head) sit after the last layer
    for l in reversed(range(n_layer)):
        lo = layer_lo[l]
        if (hi - lo) * 4 >= bucket_bytes or l == 0:
            buckets.append((l, lo, hi))
            hi = lo
    if not n_layer:
        buckets.append((-1, 0, total_hi))
        return buckets
    buckets.append((-1, 0, layer_lo[0]))  # wte / wpe: complete last
    if late:
        buckets.append((-1, head_hi, total_hi))
    return buckets


def global_loss_from_sums(sums, has_lm=True, has_emo=True):
    """loss = CE_lm + CE_emotion as means over the GLOBAL counts (model.py:704-721)."""
    lm = sums[0] / sums[1] if has_lm else 0.0
    em = sums[2] / sums[3] if has_emo else 0.0
    return lm + em


def shard_range(n, rank, world):
    """Contiguous split of n requests over `world` ranks (first ranks get the remainder)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class DataParallel:
    """Wraps an ergm_b200 GPT2LMHeadModel for one-process-per-GPU data parallelism."""

    def __init__(self, model, bucket_mb=128, process_group=None, broadcast_params=True, grad_dtype="fp32"):
        """grad_dtype "fp32": the flat fp32 gradient buffer is all-reduced in place (exact: the N-rank step equals
        the single-GPU step on the concatenated batch to fp32 rounding).  "bf16" (SURVEY 8e): every bucket is cast to
        bf16 as soon as it is final and reduced in bf16 - half the NVLink bytes; the fused optimiser reads the bf16
        buckets directly (GraphedTrainStep), the eager API copies them back into p.grad."""
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.model = model
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        eng = model.engine
        eng.ensure_params()
        if broadcast_params and self.world > 1:
            dist.broadcast(eng.store.flat, 0, group=process_group)
            eng.store.shadow_fresh = False
        self.buckets = plan_buckets(eng.store.entries, eng.L, self.bucket_bytes)
        if grad_dtype not in ("fp32", "bf16"):
            raise ValueError("grad_dtype must be 'fp32' or 'bf16'")
        self.grad_dtype = grad_dtype
        self.grad_bf16 = (torch.zeros(eng.store.total, dtype=torch.bfloat16, device=eng.store.device)
                          if grad_dtype == "bf16" and self.world > 1 else None)
        self._pending = []
        model._dp = self

    def reduce_loss_sums(self, sums):
        if self.world > 1:
            dist.all_reduce(sums, group=self.group)

    def _launch(self, layer):
        g = self.model.engine.store.grad
        for trig, lo, hi in self.buckets:
            if trig == layer and hi > lo:
                if self.grad_bf16 is not None:
                    from . import ops
                    ops.cast_f32_bf16(g[lo:hi], self.grad_bf16[lo:hi])   # the bucket is final on this stream
                    self._pending.append(dist.all_reduce(self.grad_bf16[lo:hi], group=self.group, async_op=True))
                else:
                    self._pending.append(dist.all_reduce(g[lo:hi], group=self.group, async_op=True))

    def backward(self, grad_loss, accumulate, defer_last=False):
        """Backward + bucketed all-reduce.  With defer_last the final (embedding) bucket is left in
        flight: call finish() before its range [0, split_point()) of the gradient buffer is read."""
        eng = self.model.engine
        self._pending = []
        if self.world > 1 and accumulate:
            raise RuntimeError("gradient accumulation across backward calls is not supported under DataParallel")
        eng.backward(grad_loss, accumulate=accumulate, on_layer_done=self._launch if self.world > 1 else None)
        keep = min(len(self.deferred_ranges()), len(self._pending)) if (defer_last and self.world > 1) else 0
        for w in self._pending[:len(self._pending) - keep]:
            w.wait()
        self._pending = self._pending[len(self._pending) - keep:]
        if self.grad_bf16 is not None and not defer_last:
            # eager API (loss.backward(); any optimiser on p.grad): hand the reduced bf16 gradients back in fp32
            eng.store.grad.copy_(self.grad_bf16)

    def deferred_ranges(self):
        """Element ranges of the flat buffers that become final only after the embedding stage (trigger -1):
        with defer_last their all-reduces are still in flight when backward() returns."""
        return [(lo, hi) for trig, lo, hi in self.buckets if trig == -1 and hi > lo]

    def early_ranges(self):
        """Complement of deferred_ranges(): reduced and final when backward(defer_last=True) returns."""
        total = self.model.engine.store.total
        out, cur = [], 0
        for lo, hi in sorted(self.deferred_ranges()):
            if lo > cur:
                out.append((cur, lo))
            cur = max(cur, hi)
        if cur < total:
            out.append((cur, total))
        return out

    def finish(self):
        for w in self._pending:
            w.wait()
        self._pending = []

    @torch.no_grad()
    def generate(self, input_ids, *args, **kw):
        """Per-GPU sharded batched generation: every rank decodes its contiguous slice of the
        request batch with its own paged-KV pool; the ids are all-gathered."""
        n = input_ids.shape[0]
        lo, hi = shard_range(n, self.rank, self.world)
        sl = slice(lo, hi)

        def cut(t):
            return t[sl] if (torch.is_tensor(t) and t.shape[0] == n) else t

        local = self.model.generate(cut(input_ids), *[cut(a) for a in args], **{k: cut(v) for k, v in kw.items()})
        return gather_rows(local, n, self.rank, self.world, self.group)


def gather_rows(local, n, rank, world, group=None):
    """all-gather of ragged row shards produced by shard_range back into [n, ...]."""
    if world == 1:
        return local
    sizes = [shard_range(n, r, world) for r in range(world)]
    maxr = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxr,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)], 0)

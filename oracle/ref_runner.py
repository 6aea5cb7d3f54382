"""Times the REFERENCE ITSELF for bench.py's baseline legs — BENCH INFRASTRUCTURE ONLY (never imported by
ergm_b200/).

  * host CPU: the unmodified /root/reference/src/model.py (oracle/_ref copy on the GPU box,
    imported through oracle/ref_shim.py) in fp32 on all host threads: the training step main.py:137-169 runs
    (forward, zero_grad, backward, torch.optim.AdamW.step; train mode, dropout 0.1), config-1 forward+loss and
    forward+backward, and generation both the main.py:253-282 way (batch 1, one FULL forward per token) and
    through the model's own KV-cache surface (model.py:228-236).
  * the same-box bar: the same module executed by stock torch eager ON the B200, fp32 and
    under torch.autocast(bfloat16), at configs 2 (training step) and 4 (batched cached greedy decode).

When no reference copy is available the oracle port (oracle/ergm_oracle.py, bit-identical restatement, eval
arithmetic) stands in and every result says kind = "port".
"""
import os
import statistics
import time

import torch

from . import ergm_oracle as O
from . import ref_shim

SP2 = 50259


def kind():
    return "reference" if ref_shim.available() else "port"


def _small_cfg(n_layer=12, n_embd=768, n_head=12):
    return O.OracleConfig(n_layer=n_layer, n_embd=n_embd, n_head=n_head)


def build_reference(device="cpu", dropout=0.1, caption=True, cfg=None, seed=0):
    """Reference GPT2LMHeadModel with random-init weights (seed 0) on `device`; caption=False execs the source
    with the one-line `caption_embeds = None` guard (the unpatched file raises without caption_ids)."""
    cfg = cfg or _small_cfg()
    sd = O.init_state_dict(cfg, seed=seed)
    m = ref_shim.build_reference_model(cfg, sd, no_caption_guard=not caption, dropout=dropout)
    m = m.to(device)
    m.lm_head.weight = m.transformer.wte.weight
    return m


def _batch(B, T, seed=1234, device="cpu"):
    from ergm_b200 import synthetic
    b = synthetic.make_batch(B, T, seed=seed)
    return {k: v.to(device) for k, v in b.items()}


def _fwd_kwargs(b, caption=True, labels=True):
    kw = dict(input_ids=b["input_ids"], token_type_ids=b["token_type_ids"], imgs=b["imgs"], auds=b["auds"])
    if labels:
        kw.update(labels=b["labels"], emotion_labels=b["emotion_labels"])
    if caption:
        kw["caption_ids"] = b["caption_ids"]
    return kw


# ------------------------------------------------------------------------------------------
# host CPU
# ------------------------------------------------------------------------------------------
def cpu_train_step(steps, warmup, B=2, T=256, threads=None, caption=True):
    """tokens/s of the training step on the host cores.  Returns dict(value, ms_per_step, cores, kind, sample)."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    if ref_shim.available():
        m = build_reference("cpu", dropout=0.1, caption=caption).train()
        opt = torch.optim.AdamW(m.parameters(), lr=2e-5)   # main.py:68
        b = _batch(B, T)
        kw = _fwd_kwargs(b, caption)

        def step():
            out = m(**kw)
            opt.zero_grad()
            out.loss.backward()
            opt.step()
        what = "reference model.py (unmodified, via shim) fwd + bwd + torch.optim.AdamW, train mode dropout 0.1"
    else:
        cfg = _small_cfg()
        sd = {k: v.clone().requires_grad_(True) for k, v in O.init_state_dict(cfg, seed=0).items() if k != "lm_head.weight"}
        sd["lm_head.weight"] = sd["transformer.wte.weight"]
        opt = torch.optim.AdamW([v for k, v in sd.items() if k != "lm_head.weight"], lr=2e-5)
        b = _batch(B, T)

        def step():
            opt.zero_grad()
            o = O.forward(sd, cfg, b["input_ids"], b["token_type_ids"], b["labels"], b["emotion_labels"],
                          b["imgs"], b["auds"], b["caption_ids"] if caption else None)
            o["loss"].backward()
            opt.step()
        what = "oracle port (dropout omitted) fwd + bwd + torch.optim.AdamW"
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    return dict(value=B * T / (ms / 1e3), ms_per_step=ms, cores=threads, kind=kind(),
                sample="%s, GPT-2 small %s mode, B=%d x T=%d per step (bounded sample of the B=32 step), %d timed steps"
                       % (what, "caption" if caption else "no-caption", B, T, steps))


def _median_time(fn, warm, reps):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts), min(ts)


def cpu_suite(threads=None, gen_new=16, budget_s=120.0):
    """BASELINE.md §3.4 workloads on the host cores (median of 5 after 2 warm-ups where time allows)."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    res = {"cores": threads, "kind": kind()}
    if not ref_shim.available():
        return res
    t_start = time.perf_counter()
    from ergm_b200 import synthetic
    g = synthetic.gv1_inputs()   # config 1: B=4, T=128
    m = build_reference("cpu", dropout=0.1, caption=True).eval()
    kw = dict(input_ids=g["input_ids"], token_type_ids=g["token_type_ids"], labels=g["labels"],
              emotion_labels=g["emotion_labels"], imgs=g["imgs"], auds=g["auds"], caption_ids=g["caption_ids"])
    with torch.no_grad():
        med, mn = _median_time(lambda: m(**kw), 2, 5)
    res["config1_fwd_loss"] = {"median_s": med, "min_s": mn, "tokens_per_s": 4 * 128 / med, "shape": "B=4 T=128 eval caption"}
    m.train()

    def fb():
        m.zero_grad()
        m(**kw).loss.backward()
    med, mn = _median_time(fb, 1, 3)
    res["config1_fwd_bwd"] = {"median_s": med, "min_s": mn, "tokens_per_s": 4 * 128 / med, "shape": "B=4 T=128 train dropout 0.1"}
    del m
    # generation (no-caption guard: main.py passes no caption_ids), prompt 128, greedy instead of multinomial
    mg = build_reference("cpu", dropout=0.0, caption=False).eval()
    b = _batch(1, 128, seed=99)
    ids0, tt0 = b["input_ids"], b["token_type_ids"]
    with torch.no_grad():
        # (a) main.py:253-282: batch 1, one full forward per new token
        ids, tt = ids0.clone(), tt0.clone()
        lat = []
        for _ in range(gen_new):
            if time.perf_counter() - t_start > budget_s:
                break
            t0 = time.perf_counter()
            nxt = mg(input_ids=ids, token_type_ids=tt).logits[:, -1, :].argmax(-1)
            ids = torch.cat([ids, nxt[:, None]], 1)
            tt = torch.cat([tt, torch.full((1, 1), SP2)], 1)
            lat.append(time.perf_counter() - t0)
        if lat:
            res["gen_recompute_b1"] = {"gen_tokens_per_s": 1.0 / statistics.median(lat), "p50_ms_per_token": 1e3 * statistics.median(lat),
                                       "new_tokens": len(lat), "how": "main.py:253-282 loop (batch 1, full forward per token, greedy)"}
        # (b) the model's own KV-cache surface (model.py:228-236), batch 1 and batch 8
        for B, new in ((1, 64), (8, 16)):
            if time.perf_counter() - t_start > budget_s:
                break
            bb = _batch(B, 128, seed=99)
            out = mg(input_ids=bb["input_ids"], token_type_ids=bb["token_type_ids"], use_cache=True)
            past, nxt = out.past_key_values, out.logits[:, -1, :].argmax(-1)
            lat = []
            for _ in range(new):
                t0 = time.perf_counter()
                out = mg(input_ids=nxt[:, None], token_type_ids=torch.full((B, 1), SP2), past_key_values=past, use_cache=True)
                past, nxt = out.past_key_values, out.logits[:, -1, :].argmax(-1)
                lat.append(time.perf_counter() - t0)
            p50 = statistics.median(lat)
            res["gen_kvcache_b%d" % B] = {"gen_tokens_per_s": B / p50, "p50_ms_per_step": 1e3 * p50, "new_tokens": new,
                                          "how": "reference past_key_values path, prompt 128, greedy"}
    return res


# ------------------------------------------------------------------------------------------
# torch eager on the B200 (the same-box bar)
# ------------------------------------------------------------------------------------------
def _cuda_ms(fn, warm, reps):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def gpu_eager_suite(device, B=32, T=256, gen_B=64, gen_prompt=128, gen_new=64, steps=5, medium=False):
    """Reference model in stock torch eager on the B200: training step at config 2 (fp32, autocast bf16, caption and
    no-caption) and KV-cached greedy decode at config 4 (uniform prompts: the reference has no ragged support)."""
    if not ref_shim.available():
        return {"unavailable": "no reference copy (oracle/_ref missing)"}
    res = {"kind": "reference", "engine": "torch %s eager" % torch.__version__}
    cfg = _small_cfg(24, 1024, 16) if medium else _small_cfg()
    for caption in (True, False):
        if medium and not caption:
            continue
        m = build_reference(device, dropout=0.1, caption=caption, cfg=cfg).train()
        opt = torch.optim.AdamW(m.parameters(), lr=2e-5)
        b = _batch(B, T, device=device)
        kw = _fwd_kwargs(b, caption)
        if medium:   # 768-wide features cannot be added to a 1024-wide stream (Appendix A D7): no fusion for the reference
            kw.pop("imgs"); kw.pop("auds")
        for mode in ("fp32", "autocast_bf16"):
            def step():
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "autocast_bf16")):
                    out = m(**kw)
                opt.zero_grad()
                out.loss.backward()
                opt.step()
            try:
                ms = _cuda_ms(step, 3, steps)
                res["train_%s_%s" % ("caption" if caption else "nocaption", mode)] = {
                    "ms_per_step": ms, "train_tokens_per_s": B * T / (ms / 1e3), "B": B, "T": T}
            except Exception as e:  # e.g. OOM: keep whatever was measured
                res["train_%s_%s" % ("caption" if caption else "nocaption", mode)] = {"error": "%s: %s" % (type(e).__name__, e)}
        del m, opt
        torch.cuda.empty_cache()
    if medium:
        return res
    mg = build_reference(device, dropout=0.0, caption=False).eval()
    b = _batch(gen_B, gen_prompt, seed=99, device=device)
    for mode in ("fp32", "autocast_bf16"):
        def gen():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "autocast_bf16")):
                out = mg(input_ids=b["input_ids"], token_type_ids=b["token_type_ids"], use_cache=True)
                past, nxt = out.past_key_values, out.logits[:, -1, :].argmax(-1)
                tt = torch.full((gen_B, 1), SP2, device=device)
                for _ in range(gen_new - 1):
                    out = mg(input_ids=nxt[:, None], token_type_ids=tt, past_key_values=past, use_cache=True)
                    past, nxt = out.past_key_values, out.logits[:, -1, :].argmax(-1)
        ms = _cuda_ms(gen, 1, 3)
        res["generate_kvcache_%s" % mode] = {"ms_total": ms, "gen_tokens_per_s": gen_B * gen_new / (ms / 1e3),
                                             "ms_per_token_step": ms / gen_new, "batch": gen_B, "prompt": gen_prompt,
                                             "new_tokens": gen_new}
    return res

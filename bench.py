#!/usr/bin/env python
"""Benchmark of the ERGM hot path on B200 — contract: see the task description / DESIGN.md §Measurement.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

step      = one teacher-forced training step (forward + backward + AdamW) of ERGM GPT-2 small on a
            synthetic MELD-shaped batch, B=32 x T=256 per GPU (BASELINE.json configs[1]; weak scaling).
value     = training tokens/s (B*T*N / step time) with inputs already resident in HBM (CUDA-graph replay).
e2e       = same metric through the public API ergm_b200.trainer.GraphedTrainStep with HOST (pinned)
            batches: H2D copy of the step's inputs + D2H read of the loss inside the timed region.
roofline  = dominant kernel (gemm_bf16_kernel, tcgen05): algorithmic GEMM FLOPs / CUDA-event time of
            those launches, against the measured bf16 peak in MEASURED_PEAKS.json.
cpu_baseline / --impl reference = the oracle port (oracle/ergm_oracle.py, a restatement bit-identical
            to the reference model) timed on the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "train_tokens_per_s"
UNIT = "tokens/s"
B_PER_GPU, SEQ = 32, 256
VOCAB = 50260


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def train_flops_per_token(H, L, V, T, Tc, caption=True):
    """Algorithmic FLOPs (multiply-add = 2) per token of one training step = 3 x forward
    (SURVEY.md §8d): GEMMs L*(24 or 32)H^2 + 2HV, attention L*(2TH [+4TcH])."""
    per_layer = (32 if caption else 24) * H * H
    attn = L * (2 * T * H + (4 * Tc * H if caption else 0))
    fwd = L * per_layer + 2 * H * V + attn
    return 3 * fwd, 3 * (L * per_layer + 2 * H * V)


def gemm_flops(info, dyn_count=None):
    """Algorithmic FLOPs of one launch; launches whose M / K extent is bounded at run time by a device-side count
    (label-sparse LM head) are counted with that count, not with their static capacity."""
    M, N, K = info[0], info[1], info[2]
    dyn = info[6] if len(info) > 6 else 0
    if dyn == 1 and dyn_count is not None:
        M = min(M, dyn_count)
    if dyn == 2 and dyn_count is not None:
        K = min(K, dyn_count)
    return 2.0 * M * N * K


def build(device, n_layer=12, n_embd=768, n_head=12, dropout=0.1, seed=0, feat_dim=None):
    from transformers import GPT2Config
    from ergm_b200.model import GPT2LMHeadModel
    torch.manual_seed(seed)
    cfg = GPT2Config(vocab_size=VOCAB, n_embd=n_embd, n_layer=n_layer, n_head=n_head,
                     attn_pdrop=dropout, resid_pdrop=dropout, embd_pdrop=dropout)
    if feat_dim:  # A3 extension: raw feature sequences pooled + projected feat_dim -> n_embd on the device
        cfg.ergm_visual_dim = cfg.ergm_audio_dim = feat_dim
    m = GPT2LMHeadModel(cfg).to(device)
    return m.train()


def host_batch(B, T, seed, pin=True, sequences=False, kf=1):
    from ergm_b200 import synthetic
    b = synthetic.make_batch(B, T, seed=seed, kf=kf)
    out = {k: b[k] for k in ("input_ids", "token_type_ids", "labels", "emotion_labels", "caption_ids", "imgs", "auds")}
    if sequences:  # [B, 197*kf, 768] key-frame features, [B, 113, 768] audio features (text_feature.py:44,49)
        out["imgs"], out["auds"] = b["vis_seq"].contiguous(), b["aud_seq"].contiguous()
    else:
        out["imgs"] = out["imgs"][:, 0].contiguous()
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


def run_reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (the unmodified model.py through
    oracle/ref_shim.py; the oracle port only if no copy of it is present) on all host threads.  Each step is a
    bounded sample (B=4) of the B=32 training step; rank 0 alone runs it under torchrun."""
    if rank != 0:
        return
    from oracle import ref_runner
    B = 4
    r = ref_runner.cpu_train_step(args.steps, args.warmup, B=B, T=SEQ)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ERGM GPT-2 small teacher-forced training step (fwd+bwd+AdamW), caption mode + "
                                   "img/aud fusion, dropout 0.10, synthetic MELD-shaped, CPU sample B=%d x T=%d of the "
                                   "B=%d x T=%d step" % (B, SEQ, B_PER_GPU, SEQ)},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_suite:
        # config-1 forward+loss / forward+backward and generation both ways, same host cores
        try:
            line["cpu_reference_suite"] = ref_runner.cpu_suite()
        except Exception as e:
            line["cpu_reference_suite"] = {"error": "%s: %s" % (type(e).__name__, e)}
    print(json.dumps(line), flush=True)


def run_reference_gpu_arm(args, rank, local):
    """--impl reference-gpu: the same-box bar of BASELINE.md 3.5 - the unmodified reference model executed by stock
    torch eager ON the B200 (fp32 and autocast bf16), configs 2 and 4.  Rank 0 only."""
    if rank != 0:
        return
    from oracle import ref_runner
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    medium = args.config == "medium"
    res = ref_runner.gpu_eager_suite(device, B=B_PER_GPU, T=SEQ, steps=max(3, min(args.steps, 10)), medium=medium)
    key = "train_caption_autocast_bf16"
    best = res.get(key, {})
    line = {"impl": "reference-gpu", "metric": METRIC, "value": best.get("train_tokens_per_s"), "unit": UNIT,
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": best.get("ms_per_step"),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 autocast (fp32 beside it)",
            "data": "synthetic",
            "config": {"workload": "reference ERGM GPT-2 %s in torch eager on the B200: training step B=%d x T=%d "
                                   "(caption mode, dropout 0.10, torch.optim.AdamW) and KV-cached greedy decode"
                                   % (args.config, B_PER_GPU, SEQ)},
            "reference_on_b200": res}
    print(json.dumps(line), flush=True)


def decode_bytes(eng, B, ctx, Tc=0):
    """Algorithmic HBM bytes of one decode step: bf16 weights once + the K/V each sequence reads;
    caption mode adds the q_attn / c_proj weights and the cached cross-attention K/V (Tc keys per sequence)."""
    H, L, V = eng.H, eng.L, eng.V
    w = 2 * (12 * L * H * H + V * H) + (2 * L * 2 * H * H if Tc else 0)
    kv = B * (2 * L * ctx * H * 2) + (B * 2 * L * Tc * H * 2 if Tc else 0)
    return w + kv


def bench_generation(model, device, B=64, prompt=128, new=64, reps=3):
    """BASELINE config 4: ragged prompts 64..128, 64 new tokens, paged KV; greedy (caption / no caption) and
    top-k = 50 sampling; per-token latency = graph replays of the decode step alone."""
    from ergm_b200 import synthetic
    from ergm_b200 import generation
    g = torch.Generator().manual_seed(7)
    b = synthetic.make_batch(B, prompt, seed=99, ragged=False)
    lens = torch.randint(prompt // 2, prompt + 1, (B,), generator=g)
    ids, tt, cap = b["input_ids"].to(device), b["token_type_ids"].to(device), b["caption_ids"].to(device)
    model.eval()
    res = {}
    for mode, capt, kw in (("nocaption", None, {}), ("caption", cap, {}),
                           ("nocaption_topk50", None, dict(do_sample=True, top_k=50, seed=1))):
        times = []
        for r in range(reps + 1):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = generation.generate(model, ids, tt, max_new_tokens=new, eos_token_id=None, sp2_id=50259,
                                      caption_ids=capt, prompt_lens=lens, **kw)
            e1.record()
            torch.cuda.synchronize()
            if r > 0:
                times.append(e0.elapsed_time(e1))
        ms = statistics.median(times)
        res[mode] = {"gen_tokens_per_s": B * new / (ms / 1e3), "ms_total": ms, "batch": B, "new_tokens": new}
    # per-token latency: time graph replays of the decode step alone
    ctx = float(lens.float().mean()) + new / 2
    for key, capt in (("decode_step", None), ("decode_step_caption", cap)):
        out, st = generation.generate(model, ids, tt, max_new_tokens=new, sp2_id=50259, prompt_lens=lens,
                                      caption_ids=capt, return_state=True)
        if st.graph is None:
            continue
        st.step.zero_()
        st.seq_lens.copy_(lens.to(device).int())
        lat = []
        for _ in range(new - 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); st.graph.replay(); e1.record()
            torch.cuda.synchronize()
            lat.append(e0.elapsed_time(e1))
        p50 = statistics.median(lat)
        nbytes = decode_bytes(model.engine, B, ctx, cap.shape[1] if capt is not None else 0)
        res[key] = {"p50_ms_per_token": p50, "tokens_per_s_steady": B / (p50 / 1e3),
                    "algorithmic_bytes_per_step": nbytes, "mean_ctx": ctx,
                    "hbm_gbs_achieved": nbytes / (p50 / 1e3) / 1e9,
                    "hbm_frac_of_measured": nbytes / (p50 / 1e3) / 1e9 / peaks()["hbm"],
                    "launches_per_step": getattr(st, "launches_per_step", None)}
    model.train()
    return res


def measured_traffic():
    """ncu DRAM traffic per launch / step (profiles/traffic.json, written by scripts/ncu_traffic.py from a
    committed `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` pass of this same command)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ergm_b200")
    ap.add_argument("--no-gen", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-suite", action="store_true", help="--impl reference: skip the config-1 / generation CPU suite")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-in-torch-eager-on-B200 leg of the main line")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--config", default="small", choices=["small", "medium"],
                    help="small = BASELINE configs[1] (the metric's config); medium = configs[4]: GPT-2 medium, "
                         "T=512, B=8 per GPU, 4-key-frame 768-wide feature sequences pooled + projected on device")
    args = ap.parse_args()
    global B_PER_GPU, SEQ
    if args.config == "medium":
        B_PER_GPU, SEQ = 8, 512
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.impl == "reference-gpu":
        run_reference_gpu_arm(args, rank, int(os.environ.get("LOCAL_RANK", "0")))
        return
    import torch.distributed as dist
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from ergm_b200 import ops
    from ergm_b200.optim import FusedAdamW
    from ergm_b200.trainer import GraphedTrainStep
    pk = peaks()

    medium = args.config == "medium"
    model = build(device, dropout=args.dropout, **(dict(n_layer=24, n_embd=1024, n_head=16, feat_dim=768) if medium else {}))
    opt = FusedAdamW(model, lr=2e-5)
    dp = None
    if world > 1:
        from ergm_b200.parallel import DataParallel
        dp = DataParallel(model, bucket_mb=float(os.environ.get("ERGM_BUCKET_MB", "128")),
                          grad_dtype=os.environ.get("ERGM_DP_GRAD", "bf16"))
    step = GraphedTrainStep(model, opt, dp=dp, use_graph=not args.no_graph)
    batch = host_batch(B_PER_GPU, SEQ, seed=1234 + rank, sequences=medium, kf=4 if medium else 1)
    H, L = model.config.n_embd, model.config.n_layer

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (first call captures the graph) ----
    for _ in range(args.warmup):
        loss = step(batch)
    key, st = step.copy_in(batch)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    # ---- timed: device-resident inputs ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step.run_device(key, st)
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1) / args.steps
    # ---- timed: end to end through the public API with host batches ----
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step(batch)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    clk = clocks.stop()
    t = torch.tensor([ms_dev, ms_e2e], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    tokens = B_PER_GPU * SEQ * world
    value = tokens / (ms_dev / 1e3)
    e2e_value = tokens / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel: CUDA events around every GEMM launch of two eager steps ----
    roof = None
    if rank == 0:
        eager = GraphedTrainStep(model, opt, dp=None, use_graph=False) if world == 1 else None
        if eager is not None:
            k2, s2 = eager.copy_in(batch)
            eager.run_device(k2, s2)
            torch.cuda.synchronize()
            ops.PROFILE = []
            for _ in range(2):
                eager.run_device(k2, s2)
            torch.cuda.synchronize()
            prof, ops.PROFILE = ops.PROFILE, None
            g_ms = sum(a.elapsed_time(b) for n, i, a, b in prof if n == "ergm_gemm_bf16") / 2
            # rows the label-sparse LM head scores in this batch (shifted labels != -100)
            n_scored = int((batch["labels"][:, 1:] != -100).sum())
            g_fl = sum(gemm_flops(i, n_scored) for n, i, a, b in prof if n == "ergm_gemm_bf16") / 2
            n_gemm = sum(1 for n, i, a, b in prof if n == "ergm_gemm_bf16") // 2
            by = {}
            for n, i, a, b in prof:
                by[n] = by.get(n, 0.0) + a.elapsed_time(b) / 2
            shapes = {}
            for n, i, a, b in prof:
                if n == "ergm_gemm_bf16":
                    k = "M%d N%d K%d a%d b%d sk%d" % i[:6] + (" dyn%s=%d" % ("MK"[i[6] - 1], n_scored) if i[6] else "")
                    t, c = shapes.get(k, (0.0, 0))
                    shapes[k] = (t + a.elapsed_time(b) / 2, c + 0.5)
            def shape_flops(k):
                f = k.split()
                info = [int(x[1:]) for x in f[:3]] + [0, 0, 0, ({"M": 1, "K": 2}[f[6][3]] if len(f) > 6 else 0)]
                return gemm_flops(info, n_scored)

            by_shape = {k: {"ms": round(t, 3), "launches": c, "tflops": round(shape_flops(k) * c / (t / 1e3) / 1e12, 1)}
                        for k, (t, c) in sorted(shapes.items(), key=lambda kv: -kv[1][0])}
            achieved = g_fl / (g_ms / 1e3) / 1e12
            roof = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05)", "achieved": achieved,
                    "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sust"],
                    "peak_source": pk["src"] + " (sustained bf16)",
                    "traffic": measured_traffic().get("gemm_bf16_kernel", {}).get("dram_bytes_per_launch"),
                    "traffic_source": measured_traffic().get("gemm_bf16_kernel", {}).get("source"),
                    "algorithmic_bytes_per_launch_note": "tensor-bound kernel: traffic is reported for completeness "
                                                         "(operands + output of an average launch)",
                    "gemm_launches_per_step": n_gemm, "gemm_ms_per_step": g_ms, "gemm_flops_per_step": g_fl,
                    "avg_launch_us": 1e3 * g_ms / max(n_gemm, 1), "gemm_by_shape": by_shape,
                    # the LM-head GEMMs run inside ergm_lmhead_ce_{fwd,bwd} (plan + gather + GEMM + CE in one call):
                    # timed as those entry points, not part of gemm_ms / gemm_flops above
                    "lm_head_ce": {"scored_rows": n_scored, "fwd_ms": round(by.get("ergm_lmhead_ce_fwd", 0.0), 3),
                                   "bwd_ms": round(by.get("ergm_lmhead_ce_bwd", 0.0), 3),
                                   "gemm_flops": 6.0 * n_scored * 50260 * model.config.n_embd},
                    "eager_ms_by_entry_point": {k: round(v, 3) for k, v in sorted(by.items(), key=lambda kv: -kv[1])}}
            step.eng.set_rng_step_tensor(step.rng_step)
    # ---- secondary: the mode main.py asks for (no caption_ids -> no cross-attention), SURVEY §8d ----
    nocap = None
    if not medium:
        b_nc = {k: v for k, v in batch.items() if k != "caption_ids"}
        for _ in range(3):
            step(b_nc)
        k_nc, st_nc = step.copy_in(b_nc)
        barrier()
        e0.record()
        for _ in range(args.steps):
            step.run_device(k_nc, st_nc)
        e1.record()
        barrier()
        t_nc = torch.tensor([e0.elapsed_time(e1) / args.steps], device=device)
        if world > 1:
            dist.all_reduce(t_nc, op=dist.ReduceOp.MAX)
        nocap = {"ms_per_step": float(t_nc[0]), "train_tokens_per_s": tokens / (float(t_nc[0]) / 1e3)}
    nonpad = int((batch["token_type_ids"] != 50256).sum())  # padding carries the eos id as its token type
    # ---- secondary: packed variable-length batches: only the real positions (+ position T-1 of padded
    #      samples, which the emotion head reads) are computed; same batch, same step, one CUDA graph ----
    packed = None
    if not medium:
        try:
            b_pk = dict(batch)
            b_pk["seq_lens"] = (batch["token_type_ids"] != 50256).sum(1).to(torch.int32).pin_memory()
            for _ in range(3):
                step(b_pk)
            k_pk, st_pk = step.copy_in(b_pk)
            barrier()
            e0.record()
            for _ in range(args.steps):
                step.run_device(k_pk, st_pk)
            e1.record()
            barrier()
            t_pk = torch.tensor([e0.elapsed_time(e1) / args.steps], device=device)
            if world > 1:
                dist.all_reduce(t_pk, op=dist.ReduceOp.MAX)
            rows_pk = int(model.engine.get_pack(B_PER_GPU, SEQ).n_rows.item())
            packed = {"ms_per_step": float(t_pk[0]), "train_tokens_per_s_incl_padding": tokens / (float(t_pk[0]) / 1e3),
                      "useful_tokens_per_s_rank0_rate": nonpad * world / (float(t_pk[0]) / 1e3),
                      "rows_computed_rank0": rows_pk, "rows_padded_layout": B_PER_GPU * SEQ,
                      "semantics": "reference called with the right-padded attention_mask (emotion head on position T-1)"}
        except Exception as e:
            packed = {"error": "%s: %s" % (type(e).__name__, e)}
            if world > 1:
                raise
    fl_tok, gemm_fl_tok = train_flops_per_token(H, L, VOCAB, SEQ, SEQ, caption=True)
    model_tf = value / world * fl_tok / 1e12

    gen = None
    if not args.no_gen:
        # batched inference sharded per GPU: every rank decodes its own 64 requests with its own paged-KV pool
        # (no communication); whole-job tokens/s = world x batch x new / max-over-ranks time.  Config 5 (medium): long
        # context - ragged prompts 256..512 (the model's imgs / auds here are feature SEQUENCES: generation runs
        # without fusion, like main.py's sampling loop, which passes neither)
        try:
            gen = bench_generation(model, device, prompt=512 if medium else 128)
            if world > 1:
                modes = ("nocaption", "caption", "nocaption_topk50")
                steps_k = ("decode_step", "decode_step_caption")
                tg = torch.tensor([gen[m]["ms_total"] for m in modes] +
                                  [gen.get(k, {}).get("p50_ms_per_token", 0.0) for k in steps_k], device=device)
                dist.all_reduce(tg, op=dist.ReduceOp.MAX)
                for i, mode in enumerate(modes):
                    gen[mode]["ms_total"] = float(tg[i])
                    gen[mode]["gen_tokens_per_s"] = world * gen[mode]["batch"] * gen[mode]["new_tokens"] / (float(tg[i]) / 1e3)
                    gen[mode]["batch"] *= world
                for i, k in enumerate(steps_k):
                    if k in gen:
                        gen[k]["p50_ms_per_token"] = float(tg[len(modes) + i])
                        gen[k]["tokens_per_s_steady"] = world * 64 / (float(tg[len(modes) + i]) / 1e3)
                gen["sharding"] = "%d ranks x 64 requests, no communication" % world
        except Exception as e:  # generation is a secondary line; never lose the training number
            gen = {"error": "%s: %s" % (type(e).__name__, e)}
            if world > 1:
                raise
    cpu = None
    ref_gpu = None
    if rank == 0 and world == 1 and not args.no_cpu and not medium:
        from oracle import ref_runner
        r = ref_runner.cpu_train_step(steps=2, warmup=1, B=2, T=SEQ)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
               "sample": r["sample"] + " (%.0f ms/step)" % r["ms_per_step"]}
    if rank == 0 and world == 1 and not args.no_ref_gpu and not medium:
        # the same-box bar: the reference model in torch eager on this B200, measured after
        # every number of ours has been taken (it shares nothing with our path)
        try:
            from oracle import ref_runner
            torch.cuda.empty_cache()
            ref_gpu = ref_runner.gpu_eager_suite(device, B=B_PER_GPU, T=SEQ, steps=5)
        except Exception as e:
            ref_gpu = {"error": "%s: %s" % (type(e).__name__, e)}
    roof_dec = None
    if gen and "decode_step" in gen:
        d = gen["decode_step"]
        tr = measured_traffic().get("decode_step", {})
        roof_dec = {"bound": "hbm", "kernel": "decode step = one CUDA-graph replay (%s launches): LN+QKV / paged attention "
                                              "/ out-proj / MLP slab GEMMs, LM head, arg-max" % d.get("launches_per_step"),
                    "achieved": d["hbm_gbs_achieved"], "peak": pk["hbm"], "unit": "GB/s", "frac": d["hbm_frac_of_measured"],
                    "peak_source": pk["src"], "traffic": tr.get("dram_bytes_per_step"), "traffic_source": tr.get("source"),
                    "algorithmic_bytes": d["algorithmic_bytes_per_step"], "p50_us": 1e3 * d["p50_ms_per_token"]}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "ERGM GPT-2 %s (%.1fM params, V=50260) teacher-forced training step: "
                                       "fwd + bwd + AdamW, caption mode (cross-attention) + img/aud fusion%s, "
                                       "dropout %.2f, B=%d x T=%d per GPU, synthetic %s-shaped, random init"
                                       % (args.config, sum(p.numel() for p in model.parameters()) / 1e6,
                                          " from on-device pooled + projected feature sequences" if medium else "",
                                          args.dropout, B_PER_GPU, SEQ, "IEMOCAP/MEDIC" if medium else "MELD"),
                           "per_gpu_batch": B_PER_GPU, "seq_len": SEQ, "parallelism": "dp%d" % world,
                           "dp_gradient_allreduce": (None if dp is None else
                                                     "%s buckets of >= %s MB, NCCL, overlapped with the backward, in the step's CUDA graph"
                                                     % (dp.grad_dtype, os.environ.get("ERGM_BUCKET_MB", "128"))),
                           "l2": "no flush needed: the step streams >5 GB of activations/weights/grads per "
                                 "iteration, far above the 126 MB L2",
                           "cuda_graph": bool(step.use_graph)},
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": step.h2d_bytes * world, "d2h_bytes_per_step": 24 * world,
                        "api": "ergm_b200.trainer.GraphedTrainStep(model, FusedAdamW)(pinned_host_batch) -> loss"},
                "gpu_launches": step.launches_per_step * args.steps,
                "clocks": clk, "roofline": roof, "roofline_decode": roof_dec, "cpu_baseline": cpu,
                "reference_on_b200": ref_gpu,
                "model_tflops_per_gpu": model_tf, "model_flops_per_token": fl_tok,
                "mfu_of_measured_sustained": model_tf / pk["tf_sust"], "last_loss": loss,
                "tokens_per_step": {"all_positions_incl_padding": tokens, "non_pad_rank0": nonpad},
                "nocaption_mode": nocap, "packed_mode": packed, "generation": gen}
        print(json.dumps(line), flush=True)
    if world > 1:
        # the JSON line is out; never let a teardown hang keep the launcher waiting
        sys.stdout.flush()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        step.close()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


if __name__ == "__main__":
    main()

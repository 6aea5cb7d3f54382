"""Per-CTA phase timeline of attn_bwd_kernel (%globaltimer stamps).  Profiling builds only:

    ERGM_NVCC_EXTRA=-DERGM_ATTN_TRACE python -m ergm_b200.build --force && python scripts/trace_attn_bwd.py

The shipped library has no trace code (AB_STAMP compiles to nothing)."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from ergm_b200 import _lib as L  # noqa: E402
from ergm_b200 import ops  # noqa: E402

B, nh, T, H = 32, 12, 256, 768
EV = {1: "S/dP ready", 2: "P/dS handed over", 3: "dQ (prev block) added", 4: "item MMAs done", 5: "dK/dV stored", 6: "last dQ added", 7: "dK/dV in registers", 8: "Q/dO/stats landed"}


def run(causal, p_drop):
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(B * T, 3 * H, device="cuda", generator=g).bfloat16()
    dout = torch.randn(B * T, H, device="cuda", generator=g).bfloat16()
    out = torch.zeros(B * T, H, device="cuda", dtype=torch.bfloat16)
    o32 = torch.zeros(B * T, H, device="cuda")
    lse = torch.zeros(B, nh, T, device="cuda")
    delta = torch.zeros(B, nh, T, device="cuda")
    dq = torch.zeros(B * T, H, device="cuda", dtype=torch.bfloat16)
    dkv = torch.zeros(B * T, 3 * H, device="cuda", dtype=torch.bfloat16)
    kw = dict(B=B, nh=nh, Tq=T, Tk=T, q_col0=0, k_col0=H, v_col0=2 * H, causal=causal, dropout_p=p_drop, seed=1, offset=2)
    ops.attn_fwd(qkv, qkv, qkv, out, lse, out_f32=o32, **kw)
    n_cta = 148
    trace = torch.zeros(n_cta * 64, dtype=torch.int64, device="cuda")
    lib = L.lib()
    lib.ergm_attn_bwd_set_trace.argtypes = [ctypes.c_void_p]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(4):
        if i == 3:
            lib.ergm_attn_bwd_set_trace(trace.data_ptr())
        e0.record()
        ops.attn_bwd(qkv, qkv, qkv, out, dout, lse, delta, dq, dkv, dkv, dk_col0=H, dv_col0=2 * H, out_f32=o32, **kw)
        e1.record()
        torch.cuda.synchronize()
    lib.ergm_attn_bwd_set_trace(None)
    print("== causal=%s dropout=%.1f: %.1f us (delta + bwd kernels, traced run)" % (causal, p_drop, 1e3 * e0.elapsed_time(e1)))
    t = trace.view(n_cta, 64).cpu()
    t0 = t[:, 60].min().item()
    life = (t[:, 62] - t[:, 60]).float() / 1e3
    print("CTA lifetimes: mean %.1f us, min %.1f, max %.1f; alloc+sync %.2f us; kernel span %.1f us" %
          (life.mean(), life.min(), life.max(), ((t[:, 61] - t[:, 60]).float() / 1e3).mean(), (t[:, 62].max().item() - t0) / 1e3))
    for c in (0, 100, 147):
        prev = t[c, 61].item()
        line = []
        for i in range(30):
            ev, ts = int(t[c, 2 * i]), int(t[c, 2 * i + 1])
            if ev == 0:
                break
            line.append("%s +%.2f" % (EV[ev], (ts - prev) / 1e3))
            prev = ts
        print(" CTA %d (SM %d), first events after alloc: %s" % (c, int(t[c, 63]), "; ".join(line)))
    # mean duration of each event kind (time since the previous logged event of the same CTA)
    acc = {}
    for c in range(n_cta):
        prev = t[c, 61].item()
        for i in range(30):
            ev, ts = int(t[c, 2 * i]), int(t[c, 2 * i + 1])
            if ev == 0:
                break
            if i > 0:
                acc.setdefault(ev, []).append((ts - prev) / 1e3)
            prev = ts
    for ev in sorted(acc):
        v = acc[ev]
        print("   %-24s mean +%.2f us over %d events (max %.2f)" % (EV[ev], sum(v) / len(v), len(v), max(v)))


if __name__ == "__main__":
    run(True, 0.1)
    run(False, 0.1)
    run(True, 0.0)

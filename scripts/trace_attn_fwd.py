"""Per-CTA phase timeline of the persistent attn_fwd_kernel (%globaltimer stamps of one softmax thread).  Profiling builds:

    ERGM_NVCC_EXTRA=-DERGM_ATTN_TRACE python -m ergm_b200.build --force && python scripts/trace_attn_fwd.py
"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from ergm_b200 import _lib as L  # noqa: E402
from ergm_b200 import ops  # noqa: E402

B, nh, T, H = 32, 12, 256, 768
EV = {1: "previous block done", 2: "S ready", 3: "row max (4 TMEM loads)", 4: "max exchanged", 5: "P written (exp, dropout, TMEM store)",
      6: "O_j ready", 7: "O rescaled", 8: "item stored"}


def run(causal, p_drop):
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(B * T, 3 * H, device="cuda", generator=g).bfloat16()
    out = torch.zeros(B * T, H, device="cuda", dtype=torch.bfloat16)
    o32 = torch.zeros(B * T, H, device="cuda")
    lse = torch.zeros(B, nh, T, device="cuda")
    kw = dict(B=B, nh=nh, Tq=T, Tk=T, q_col0=0, k_col0=H, v_col0=2 * H, causal=causal, dropout_p=p_drop, seed=1, offset=2)
    n_cta = 296
    trace = torch.zeros(n_cta * 64, dtype=torch.int64, device="cuda")
    lib = L.lib()
    lib.ergm_attn_fwd_set_trace.argtypes = [ctypes.c_void_p]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(4):
        if i == 3:
            lib.ergm_attn_fwd_set_trace(trace.data_ptr())
        e0.record()
        ops.attn_fwd(qkv, qkv, qkv, out, lse, out_f32=o32, **kw)
        e1.record()
        torch.cuda.synchronize()
    lib.ergm_attn_fwd_set_trace(None)
    print("== causal=%s dropout=%.1f: %.1f us (traced run)" % (causal, p_drop, 1e3 * e0.elapsed_time(e1)))
    t = trace.view(n_cta, 64).cpu()
    acc = {}
    for c in range(n_cta):
        prev = None
        for i in range(31):
            ev, ts = int(t[c, 2 * i]), int(t[c, 2 * i + 1])
            if ev == 0:
                break
            if prev is not None:
                acc.setdefault(ev, []).append((ts - prev) / 1e3)
            prev = ts
    for ev in sorted(acc):
        v = acc[ev]
        print("   %-40s mean +%.2f us over %d events (max %.2f)" % (EV[ev], sum(v) / len(v), len(v), max(v)))
    c = 0
    prev = int(t[c, 1])
    line = []
    for i in range(1, 31):
        ev, ts = int(t[c, 2 * i]), int(t[c, 2 * i + 1])
        if ev == 0:
            break
        line.append("%d:+%.2f" % (ev, (ts - prev) / 1e3))
        prev = ts
    print("   CTA 0: " + " ".join(line))


if __name__ == "__main__":
    run(True, 0.1)
    run(False, 0.1)
    run(True, 0.0)

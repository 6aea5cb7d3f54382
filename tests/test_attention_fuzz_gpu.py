"""Randomised shapes for the persistent attention kernels (forward + backward) against fp32 torch (model.py:119-148
restated in tests/test_kernels_gpu.py::_attn_ref): item counts below / above the number of resident CTAs, sequence
lengths that end inside a 128-row block, per-sample key lengths (including samples with very few keys), both the
whole-head (T <= 256) and the key-block (T > 256) work-item modes of the backward, self (causal, fused qkv buffer) and
cross (separate key / value buffer, Tk != Tq) layouts.  The persistent kernels carry their mbarrier phases across work
items, so a phase slip on an unusual item sequence shows up here as a wrong result or a hang (pytest-timeout)."""
import random

import pytest
import torch

from test_kernels_gpu import _attn_ref

pytestmark = pytest.mark.gpu


def _cases():
    rng = random.Random(1234)
    out = []
    for i in range(18):
        causal = i % 2 == 0
        B = rng.choice([1, 2, 3, 5, 9, 20])
        nh = rng.choice([1, 2, 3, 12, 16])
        if B * nh > 400:
            B = max(1, 400 // nh)
        Tq = rng.choice([1, 17, 64, 127, 128, 129, 200, 256, 257, 300, 384, 500])
        Tk = Tq if causal else rng.choice([1, 40, 128, 150, 256, 300, 411])
        lens = rng.choice([None, "ragged", "short"])
        out.append((B, nh, Tq, Tk, causal, lens, i))
    # more work items than resident CTAs (148 backward, 296 forward): several rounds of the persistent loops
    out += [(20, 16, 256, 256, True, "ragged", 18), (25, 12, 300, 300, True, None, 19), (30, 12, 129, 200, False, "ragged", 20),
            (32, 12, 256, 256, False, None, 21)]
    return out


@pytest.mark.timeout(120)
@pytest.mark.parametrize("B,nh,Tq,Tk,causal,lens_mode,seed", _cases())
def test_attention_random_shapes(cuda_device, B, nh, Tq, Tk, causal, lens_mode, seed):
    from ergm_b200 import ops
    H = nh * 64
    g = torch.Generator(device="cuda").manual_seed(100 + seed)
    lens = None
    if lens_mode == "ragged":
        lens = torch.randint(1, Tk + 1, (B,), device="cuda", generator=g).to(torch.int32)
    elif lens_mode == "short":
        lens = torch.randint(1, min(Tk, 3) + 1, (B,), device="cuda", generator=g).to(torch.int32)
    if causal:
        qkv = torch.randn(B * Tq, 3 * H, device="cuda", generator=g).bfloat16()
        qm, km, vm, qc, kc, vc = qkv, qkv, qkv, 0, H, 2 * H
        dbuf = torch.full((B * Tq, 3 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
        dqm, dkm, dvm, dqc, dkc, dvc = dbuf, dbuf, dbuf, 0, H, 2 * H
    else:
        qm = torch.randn(B * Tq, H, device="cuda", generator=g).bfloat16()
        kvm = torch.randn(B * Tk, 2 * H, device="cuda", generator=g).bfloat16()
        km, vm, qc, kc, vc = kvm, kvm, 0, 0, H
        dqm = torch.full((B * Tq, H), float("nan"), device="cuda", dtype=torch.bfloat16)
        dkv = torch.full((B * Tk, 2 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
        dkm, dvm, dqc, dkc, dvc = dkv, dkv, 0, 0, H
    out = torch.full((B * Tq, H), float("nan"), device="cuda", dtype=torch.bfloat16)
    o32 = torch.full((B * Tq, H), float("nan"), device="cuda")
    lse = torch.zeros(B, nh, Tq, device="cuda")
    kw = dict(B=B, nh=nh, Tq=Tq, Tk=Tk, q_col0=qc, k_col0=kc, v_col0=vc, causal=causal, kv_lens=lens)
    ops.attn_fwd(qm, km, vm, out, lse, out_f32=o32, **kw)
    q = qm[:, qc:qc + H].float().view(B, Tq, nh, 64).permute(0, 2, 1, 3).clone().requires_grad_(True)
    k = km[:, kc:kc + H].float().view(B, Tk, nh, 64).permute(0, 2, 1, 3).clone().requires_grad_(True)
    v = vm[:, vc:vc + H].float().view(B, Tk, nh, 64).permute(0, 2, 1, 3).clone().requires_grad_(True)
    ro, rlse = _attn_ref(q, k, v, causal, Tk - Tq, lens)
    want = ro.permute(0, 2, 1, 3).reshape(B * Tq, H)
    # causal with Tk == Tq: every query sees key 0; with key lengths every sample has >= 1 key: no empty rows
    assert torch.isfinite(out.float()).all() and torch.isfinite(o32).all()
    assert (out.float() - want).abs().max().item() < 2e-2
    assert (o32 - want).abs().max().item() < 2e-2
    assert (lse - rlse).abs().max().item() < 2e-2
    dout = torch.randn(B * Tq, H, device="cuda", generator=g).bfloat16()
    delta = torch.empty(B, nh, Tq, device="cuda")
    cs = torch.zeros(3, H, device="cuda")
    ops.attn_bwd(qm, km, vm, out, dout, lse, delta, dqm, dkm, dvm, dq_col0=dqc, dk_col0=dkc, dv_col0=dvc, out_f32=o32,
                 dq_colsum=cs[0], dk_colsum=cs[1], dv_colsum=cs[2], **kw)
    ro.backward(dout.float().view(B, Tq, nh, 64).permute(0, 2, 1, 3))
    rdq = q.grad.permute(0, 2, 1, 3).reshape(B * Tq, H)
    rdk = k.grad.permute(0, 2, 1, 3).reshape(B * Tk, H)
    rdv = v.grad.permute(0, 2, 1, 3).reshape(B * Tk, H)

    def rel(a, b):   # relative to the reference's norm, with an absolute floor (T = 1: dQ = dK = 0 exactly)
        return ((a - b).norm() / (b.norm() + 2e-2 * b.numel() ** 0.5)).item()

    dq = dqm[:, dqc:dqc + H].float()
    dk = dkm[:, dkc:dkc + H].float()
    dv = dvm[:, dvc:dvc + H].float()
    assert torch.isfinite(dq).all() and torch.isfinite(dk).all() and torch.isfinite(dv).all()   # every row written
    assert rel(dq, rdq) < 2e-2, ("dq", rel(dq, rdq))
    assert rel(dk, rdk) < 2e-2, ("dk", rel(dk, rdk))
    assert rel(dv, rdv) < 2e-2, ("dv", rel(dv, rdv))
    for got, mat in ((cs[0], dq), (cs[1], dk), (cs[2], dv)):
        w = mat.sum(0)
        assert (got - w).abs().max().item() < 4e-3 * (1 + w.abs().max().item())

"""Label-sparse LM head (model.py:698-708): the head, its cross-entropy and their backward run on
the rows whose shifted label is not -100, with a RUN-TIME row count on the device.  Kernel-level checks of the plan /
gather / scatter kernels and of the GEMM's run-time M / K extents against plain PyTorch, then the model-level
equivalence: same loss, same gradients as the dense head and as the oracle; outputs.logits still there on access.
"""
import pytest
import torch

from oracle import ergm_oracle as O
from ergm_b200 import synthetic

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def test_plan_gather_scatter(cuda_device):
    from ergm_b200 import ops
    B, T, H = 5, 37, 128
    g = torch.Generator().manual_seed(0)
    labels = torch.randint(0, 100, (B, T), generator=g)
    labels[torch.rand(B, T, generator=g) < 0.7] = -100
    lab = labels.cuda()
    M = B * T
    row_idx = torch.full((M,), -7, dtype=torch.int32, device="cuda")
    labels_c = torch.full((M,), -7, dtype=torch.int64, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.lm_rows_plan(lab, row_idx, labels_c, cnt, T=T)
    want_rows = [b * T + t for b in range(B) for t in range(T - 1) if labels[b, t + 1] != -100]
    n = int(cnt.item())
    assert n == len(want_rows)
    assert row_idx[:n].cpu().tolist() == want_rows
    assert labels_c[:n].cpu().tolist() == [int(labels.view(-1)[r + 1]) for r in want_rows]
    pad = min((n + 127) // 128 * 128, M)
    assert (row_idx[n:pad] == -1).all() and (labels_c[n:pad] == -100).all()
    src = torch.randn(M, H, generator=g).cuda().bfloat16()
    dst = torch.full((M, H), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.gather_rows_dyn(src, row_idx, cnt, dst)
    assert torch.equal(dst[:n], src[torch.tensor(want_rows, device="cuda")])
    assert (dst[n:pad] == 0).all()
    back = torch.zeros(M, H, device="cuda")
    ops.scatter_rows_dyn(dst.float().contiguous(), row_idx, cnt, back)
    want = torch.zeros(M, H, device="cuda")
    want[torch.tensor(want_rows, device="cuda")] = src.float()[torch.tensor(want_rows, device="cuda")]
    assert torch.equal(back, want)


@pytest.mark.parametrize("count", [0, 1, 130, 700, 1024])
def test_gemm_runtime_m_and_k(cuda_device, count):
    """dyn_m: D rows < count equal the static GEMM, rows >= count untouched (even with NaN operand rows behind the
    count); atomic dyn_m with the in-kernel K split; dyn_k: reduction over the first `count` rows only."""
    from ergm_b200 import _lib as L
    from ergm_b200 import ops
    cap, N, K = 1024, 512, 1088
    g = torch.Generator().manual_seed(count)
    a = torch.randn(cap, K, generator=g).cuda().bfloat16()
    w = torch.randn(N, K, generator=g).cuda().bfloat16()     # K-major B operand ([N, K])
    a[count:] = float("nan")                                  # stale rows behind the count must not leak
    cnt = torch.tensor([count], dtype=torch.int32, device="cuda")
    ref = a[:count].float() @ w.float().t()
    for bn in (0, 2256):
        d = torch.full((cap, N), -3.0, dtype=torch.bfloat16, device="cuda")
        ops.gemm(a, w, d, M=cap, N=N, K=K, a_major=L.ERGM_MAJOR_K, b_major=L.ERGM_MAJOR_K, dyn_m=cnt, block_n=bn)
        assert (d[count:] == -3.0).all()
        if count:
            assert rel(d[:count], ref) < 1e-2
        d32 = torch.zeros(cap, N, device="cuda")
        ops.gemm(a, w, d32, M=cap, N=N, K=K, a_major=L.ERGM_MAJOR_K, b_major=L.ERGM_MAJOR_K, dyn_m=cnt, block_n=bn,
                 epilogue=L.EPI_ATOMIC)
        assert (d32[count:] == 0).all()
        if count:
            assert rel(d32[:count], ref) < 1e-2
    # dyn_k: dW[K_in, N] += x[:count]^T @ dy[:count]  (both operands MN-major, rows = the reduction dimension)
    x = torch.randn(cap, 256, generator=g).cuda().bfloat16()
    dy = torch.randn(cap, 384, generator=g).cuda().bfloat16()
    pad = min((count + 127) // 128 * 128, cap)
    x[count:pad] = 0
    dy[count:pad] = 0
    x[pad:] = float("nan")
    dy[pad:] = float("nan")
    for bn, sk in ((0, 1), (2256, 1), (128, 3)):
        dw = torch.zeros(256, 384, device="cuda")
        ops.gemm(x, dy, dw, M=256, N=384, K=cap, a_major=L.ERGM_MAJOR_MN, b_major=L.ERGM_MAJOR_MN, epilogue=L.EPI_ATOMIC,
                 split_k=sk, block_n=bn, dyn_k=cnt)
        want = x[:count].float().t() @ dy[:count].float()
        if count:
            assert rel(dw, want) < 1e-2, (bn, sk)
        else:
            assert (dw == 0).all()


def _build(cfg, sd):
    from transformers import GPT2Config
    from ergm_b200.model import GPT2LMHeadModel
    hf = GPT2Config(vocab_size=cfg.vocab_size, n_positions=cfg.n_positions, n_embd=cfg.n_embd, n_layer=cfg.n_layer,
                    n_head=cfg.n_head, attn_pdrop=0.0, resid_pdrop=0.0, embd_pdrop=0.0)
    m = GPT2LMHeadModel(hf)
    m.load_state_dict(sd)
    return m.to("cuda").train()


@pytest.mark.parametrize("vocab", [1024, 50260])
def test_sparse_head_equals_dense_head_and_oracle(cuda_device, vocab):
    cfg = O.OracleConfig(vocab_size=vocab, n_positions=256, n_embd=128, n_layer=2, n_head=2)
    sd = O.init_state_dict(cfg, seed=51, perturb=True)
    b = synthetic.make_batch(6, 96, seed=52, vocab=vocab, feat_dim=128)
    kw = dict(input_ids=b["input_ids"].cuda(), token_type_ids=b["token_type_ids"].cuda(), labels=b["labels"].cuda(),
              emotion_labels=b["emotion_labels"].cuda(), caption_ids=b["caption_ids"].cuda(), imgs=b["imgs"].cuda(),
              auds=b["auds"].cuda())
    sparse, dense = _build(cfg, sd), _build(cfg, sd)
    dense.ergm_sparse_lm_head = False
    os_, od = sparse(**kw), dense(**kw)
    n_scored = int((b["labels"][:, 1:] != -100).sum())
    assert int(sparse.engine.saved["sparse"]["count"].item()) == n_scored
    assert dense.engine.saved["sparse"] is None
    assert abs(os_.loss.item() - od.loss.item()) < 2e-4 and abs(os_.lm_loss.item() - od.lm_loss.item()) < 2e-4
    # all-position logits are still there when read (main.py:160) and equal the dense head's
    assert rel(os_.logits, od.logits) < 1e-6
    os_.loss.backward()
    od.loss.backward()
    ps, pd = dict(sparse.named_parameters()), dict(dense.named_parameters())
    for n in ps:
        assert rel(ps[n].grad, pd[n].grad) < 5e-3, (n, rel(ps[n].grad, pd[n].grad))
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "lm_head.weight"}
    sdo["lm_head.weight"] = sdo["transformer.wte.weight"]
    o = O.forward(sdo, cfg, b["input_ids"], b["token_type_ids"], b["labels"], b["emotion_labels"], b["imgs"], b["auds"],
                  b["caption_ids"])
    o["loss"].backward()
    assert abs(os_.lm_loss.item() - o["lm_loss"].item()) < 1e-3
    for n in ("transformer.wte.weight", "transformer.ln_f.weight", "transformer.h.1.mlp.c_proj.weight",
              "transformer.h.0.attn.c_attn.weight"):
        assert rel(ps[n].grad, sdo[n].grad) < 3e-2, (n, rel(ps[n].grad, sdo[n].grad))


def test_sparse_head_edge_label_patterns(cuda_device):
    """All labels ignored but one; every position scored; reading .logits after the next forward raises."""
    from ergm_b200 import _lib as L
    cfg = O.OracleConfig(vocab_size=512, n_positions=64, n_embd=128, n_layer=1, n_head=2)
    sd = O.init_state_dict(cfg, seed=53, perturb=True)
    m = _build(cfg, sd)
    b = synthetic.make_batch(3, 40, seed=54, vocab=512, feat_dim=128)
    for pattern in ("one", "all"):
        lab = torch.full_like(b["labels"], -100)
        if pattern == "one":
            lab[1, 17] = 5
        else:
            lab = b["input_ids"].clone()
        out = m(input_ids=b["input_ids"].cuda(), token_type_ids=b["token_type_ids"].cuda(), labels=lab.cuda())
        o = O.forward(sd, cfg, b["input_ids"], b["token_type_ids"], lab)
        # one scored token: the loss IS one bf16-operand logit row (no averaging): 5e-3; all positions: 2e-3
        assert abs(out.loss.item() - o["lm_loss"].item()) < (5e-3 if pattern == "one" else 2e-3), pattern
        out.loss.backward()
        m.zero_grad()
    stale = m(input_ids=b["input_ids"].cuda(), token_type_ids=b["token_type_ids"].cuda(), labels=lab.cuda())
    m(input_ids=b["input_ids"].cuda(), token_type_ids=b["token_type_ids"].cuda(), labels=lab.cuda())
    with pytest.raises(L.ErgmError):
        stale.logits


def test_lmhead_ce_c_abi_vs_torch(cuda_device):
    """ergm_lmhead_ce_fwd / _bwd called directly (one workspace, no other state) against fp32 torch on the same
    bf16-rounded operands: loss sums, compacted rows, d hn (zero on unscored rows), d wte accumulated."""
    from ergm_b200 import _lib as L
    from ergm_b200 import ops
    B, T, H, V = 4, 50, 128, 1000           # 200 rows: not a multiple of the 128-row tile; V not a multiple of 64
    g = torch.Generator().manual_seed(3)
    hn = torch.randn(B * T, H, generator=g).cuda().bfloat16()
    wte = (0.2 * torch.randn(V, H, generator=g)).cuda().bfloat16()
    labels = torch.randint(0, V, (B, T), generator=g)
    labels[torch.rand(B, T, generator=g) < 0.6] = -100
    lab = labels.cuda()
    off = ops.lmhead_ce_layout(B * T, H, V, True)
    assert off[-1] > ops.lmhead_ce_layout(B * T, H, V, False)[-1]
    buf = torch.zeros(off[-1], dtype=torch.uint8, device="cuda")
    sums = torch.zeros(4, device="cuda")
    ops.lmhead_ce_fwd(hn, wte, lab, sums, buf, T=T, V=V)
    v = ops.lmhead_ce_views(buf, B * T, H, V, True)
    # reference: shifted labels, ignore_index = -100 (model.py:705-708)
    hn32 = hn.float().requires_grad_(True)
    w32 = wte.float().requires_grad_(True)
    logits = (hn32 @ w32.t()).view(B, T, V)
    loss_sum = torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, V), lab[:, 1:].reshape(-1), ignore_index=-100,
                                                 reduction="sum")
    n_valid = int((labels[:, 1:] != -100).sum())
    assert int(v["count"].item()) == n_valid and sums[1].item() == n_valid
    assert abs(sums[0].item() - loss_sum.item()) / loss_sum.item() < 2e-3
    rows = [b * T + t for b in range(B) for t in range(T - 1) if labels[b, t + 1] != -100]
    assert v["row_idx"][:n_valid].cpu().tolist() == rows
    assert rel(v["logits_c"][:n_valid, :V], logits.detach().view(B * T, V)[rows]) < 1e-2
    scale = torch.tensor([0.5], device="cuda")
    (0.5 * loss_sum).backward()
    dhn = torch.full((B * T, H), float("nan"), device="cuda")
    dwte = torch.ones(V, H, device="cuda")
    ops.lmhead_ce_bwd(wte, scale, dhn, dwte, buf, V=V)
    assert rel(dhn, hn32.grad) < 2e-2
    unscored = torch.ones(B * T, dtype=torch.bool)
    unscored[rows] = False
    assert (dhn[unscored.cuda()] == 0).all()
    assert rel(dwte - 1.0, w32.grad) < 2e-2
    # argument checks: workspace too small / misaligned -> ERGM_ERR_ARG, nothing launched
    with pytest.raises(L.ErgmError):
        ops.lmhead_ce_fwd(hn, wte, lab, sums, buf[: off[4]], T=T, V=V)
    with pytest.raises(L.ErgmError):
        ops.lmhead_ce_fwd(hn, wte, lab, sums, buf[8:], T=T, V=V)

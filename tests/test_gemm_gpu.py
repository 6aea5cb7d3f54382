"""GPU parity of ergm_gemm_bf16 (tcgen05) against a plain fp32 torch matmul of the
bf16-rounded operands (the operation Conv1D/addmm performs at model.py:222,244,263,265,698)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    B = b.float() if b_mn else b.float().t()
    return A @ B


def _run(M, N, K, a_mn, b_mn, block_n=0, out_dtype=torch.float32, split_k=1, atomic=False, seed=0):
    from ergm_b200 import ops, _lib as L
    g = torch.Generator(device="cuda").manual_seed(seed)
    if a_mn:
        a = torch.randn(K, (M + 7) // 8 * 8, device="cuda", generator=g).bfloat16()[:, :M]
    else:
        a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    if b_mn:  # rows of a TMA operand must be 16-byte aligned: pad the leading dimension
        b = torch.randn(K, (N + 7) // 8 * 8, device="cuda", generator=g).bfloat16()[:, :N]
    else:
        b = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    ldd = (N + 7) // 8 * 8
    d = torch.zeros(M, ldd, device="cuda", dtype=out_dtype)
    ops.gemm(a, b, d, M=M, N=N, K=K, a_major=int(a_mn), b_major=int(b_mn), block_n=block_n,
             split_k=split_k, epilogue=L.EPI_ATOMIC if atomic else 0)
    torch.cuda.synchronize()
    ref = _ref(a, b, a_mn, b_mn)
    got = d[:, :N].float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    return err, scale, d


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("block_n", [64, 128, 256])
def test_gemm_majors(cuda_device, a_mn, b_mn, block_n):
    err, scale, _ = _run(256, 512, 320, a_mn, b_mn, block_n)
    assert err <= 2e-3 * scale, (err, scale)


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("block_n", [2128, 2256])
@pytest.mark.parametrize("M,N,K", [(512, 512, 320), (300, 200, 136), (8192, 768, 768)])
def test_gemm_cta_pair(cuda_device, a_mn, b_mn, block_n, M, N, K):
    """tcgen05.mma.cta_group::2 kernel: 256-row tiles split over a 2-CTA cluster."""
    err, scale, d = _run(M, N, K, a_mn, b_mn, block_n)
    assert err <= 2e-3 * scale, (err, scale)
    if d.shape[1] > N:
        assert d[:, N:].abs().max().item() == 0


def test_gemm_cta_pair_splitk_and_epilogues(cuda_device):
    from ergm_b200 import ops, _lib as L
    err, scale, _ = _run(768, 768, 4096, 1, 1, block_n=2128, split_k=4, atomic=True)
    assert err <= 2e-3 * scale
    M, N, K = 1000, 3072, 768
    g = torch.Generator(device="cuda").manual_seed(1)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(K, N, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    ref = a.float() @ w.float() + bias
    out = res.clone()
    ops.gemm(a, w, out, M=M, N=N, K=K, bias=bias, residual=out, block_n=2256)
    assert (out - (ref + res)).abs().max().item() < 1e-4
    d = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, d, M=M, N=N, K=K, bias=bias, preact=pre, epilogue=L.EPI_GELU, block_n=2256)
    assert (d.float() - torch.nn.functional.gelu(ref, approximate="tanh")).abs().max().item() < 2e-2
    assert (pre.float() - ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (8192, 2304, 768), (100, 72, 200), (1, 8, 8),
                                   (333, 50260, 768), (4096, 768, 3072), (77, 1000, 1024)])
def test_gemm_shapes_fwd_layout(cuda_device, M, N, K):
    err, scale, d = _run(M, N, K, 0, 1)
    assert err <= 2e-3 * scale, (err, scale)
    # padding columns must stay untouched
    assert d[:, N:].abs().max().item() == 0 if d.shape[1] > N else True


def test_gemm_bf16_out(cuda_device):
    err, scale, _ = _run(512, 768, 768, 0, 1, out_dtype=torch.bfloat16)
    assert err <= 1e-2 * scale


@pytest.mark.parametrize("split_k", [1, 2, 5])
def test_gemm_wgrad_splitk(cuda_device, split_k):
    err, scale, _ = _run(768, 768, 4096, 1, 1, split_k=split_k, atomic=True)
    assert err <= 2e-3 * scale, (err, scale)


def test_gemm_epilogues(cuda_device):
    from ergm_b200 import ops, _lib as L
    M, N, K = 300, 3072, 768
    g = torch.Generator(device="cuda").manual_seed(1)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(K, N, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    pre_ref = a.float() @ w.float() + bias
    # bias + gelu (+ preact)
    d = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, d, M=M, N=N, K=K, bias=bias, preact=pre, epilogue=L.EPI_GELU)
    gel = torch.nn.functional.gelu(pre_ref, approximate="tanh")
    assert (pre.float() - pre_ref).abs().max().item() < 2e-2
    assert (d.float() - gel).abs().max().item() < 2e-2
    # exact gelu, fp32 out
    d32 = torch.empty(M, N, device="cuda")
    ops.gemm(a, w, d32, M=M, N=N, K=K, bias=bias, epilogue=L.EPI_GELU | L.EPI_EXACT)
    assert (d32 - gel).abs().max().item() < 1e-4
    # bias + residual, fp32 out, in place on the residual
    out = res.clone()
    ops.gemm(a, w, out, M=M, N=N, K=K, bias=bias, residual=out)
    assert (out - (pre_ref + res)).abs().max().item() < 1e-4


def test_gemm_dropout_epilogue(cuda_device):
    from ergm_b200 import ops
    M, N, K = 512, 768, 64
    a = torch.ones(M, K, device="cuda").bfloat16()
    w = torch.ones(K, N, device="cuda").bfloat16()
    d = torch.empty(M, N, device="cuda")
    ops.gemm(a, w, d, M=M, N=N, K=K, dropout_p=0.25, seed=1234, offset=7)
    keep = (d != 0).float().mean().item()
    assert abs(keep - 0.75) < 0.01
    assert torch.allclose(d[d != 0], torch.full_like(d[d != 0], 64 / 0.75))
    d2 = torch.empty(M, N, device="cuda")
    ops.gemm(a, w, d2, M=M, N=N, K=K, dropout_p=0.25, seed=1234, offset=7)
    assert torch.equal(d, d2)
    ops.gemm(a, w, d2, M=M, N=N, K=K, dropout_p=0.25, seed=1234, offset=8)
    assert not torch.equal(d, d2)


@pytest.mark.parametrize("block_n", [0, 128, 192, 256, 2256])
def test_gemm_gelu_grad_with_fused_bias_colsum(cuda_device, block_n):
    """dgrad of the MLP output projection with gelu_new' and the c_fc bias gradient fused in the epilogue
    (model.py:263-266 backward): D = (A @ B^T) * gelu_new'(u) as bf16, colsum += column sums of D as stored."""
    from ergm_b200 import ops
    M, N, K = 512, 768, 256
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = (0.1 * torch.randn(N, K, device="cuda", generator=g)).bfloat16()
    u = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    d = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    cs = torch.full((N,), 0.25, device="cuda")
    ops.gemm(a, b, d, M=M, N=N, K=K, a_major=0, b_major=0, gelu_grad_of=u, colsum=cs, block_n=block_n)
    uf = u.float().requires_grad_(True)
    torch.nn.functional.gelu(uf, approximate="tanh").sum().backward()
    ref = (a.float() @ b.float().t()) * uf.grad
    assert ((d.float() - ref).norm() / ref.norm()).item() < 1e-2
    want = 0.25 + d.float().sum(0)
    assert (cs - want).abs().max().item() < 2e-3 * (1 + want.abs().max().item())
    # without the column sum the same mode still works, and non-tile-multiple shapes refuse the fused sum
    d2 = torch.zeros_like(d)
    ops.gemm(a, b, d2, M=M, N=N, K=K, a_major=0, b_major=0, gelu_grad_of=u, block_n=block_n)
    assert torch.equal(d, d2)
    with pytest.raises(Exception):
        ops.gemm(a[:500], b, d[:500], M=500, N=N, K=K, a_major=0, b_major=0, gelu_grad_of=u[:500], colsum=cs, block_n=block_n)

"""Thin Python wrappers (raw pointers in, nothing returned) around the C ABI.

These are *not* autograd aware; `ergm_b200.functional` builds the
torch.autograd.Functions on top of them.
"""
import ctypes

import torch

from . import _lib as L


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def gemm(a, b, d, *, M, N, K, a_major=L.ERGM_MAJOR_K, b_major=L.ERGM_MAJOR_MN, lda=None, ldb=None,
         ldd=None, bias=None, residual=None, ldr=None, preact=None, epilogue=0, split_k=1, block_n=0,
         dropout_p=0.0, seed=0, offset=0):
    """D[M,N] = epilogue(A * B).  a/b bf16, d bf16 or fp32.  See include/ergm_b200.h."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    args = L.GemmArgs()
    args.a, args.b, args.d = a.data_ptr(), b.data_ptr(), d.data_ptr()
    args.bias = bias.data_ptr() if bias is not None else None
    args.residual = residual.data_ptr() if residual is not None else None
    args.preact = preact.data_ptr() if preact is not None else None
    args.lda = a.stride(0) if lda is None else lda
    args.ldb = b.stride(0) if ldb is None else ldb
    args.ldd = d.stride(0) if ldd is None else ldd
    args.ldr = (residual.stride(0) if residual is not None else 0) if ldr is None else ldr
    args.M, args.N, args.K = M, N, K
    args.a_major, args.b_major = a_major, b_major
    args.d_dtype = L.DT_F32 if d.dtype == torch.float32 else L.DT_BF16
    if bias is not None:
        epilogue |= L.EPI_BIAS
    if residual is not None:
        epilogue |= L.EPI_RESIDUAL
    if preact is not None:
        epilogue |= L.EPI_PREACT
    if dropout_p > 0.0:
        epilogue |= L.EPI_DROPOUT
    args.epilogue = epilogue
    args.split_k = split_k
    args.block_n = block_n
    args.dropout_p = dropout_p
    args.seed, args.offset = seed, offset
    L.check(L.lib().ergm_gemm_bf16(ctypes.byref(args), _stream()), "ergm_gemm_bf16")

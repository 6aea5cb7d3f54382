"""Thin Python wrappers (raw pointers in, nothing returned) around the C ABI of include/ergm_b200.h.

One function per entry point; they are not autograd aware - ergm_b200.engine composes them into the
forward and the hand-written backward, and ergm_b200.model hooks that backward into loss.backward().
Every call counts its kernel launches (bench.py's gpu_launches) and, when ops.PROFILE is a list, brackets
itself with CUDA events on the current stream.
"""
import ctypes
import functools

import torch

from . import _lib as L


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def gemm(a, b, d, *, M, N, K, a_major=L.ERGM_MAJOR_K, b_major=L.ERGM_MAJOR_MN, lda=None, ldb=None,
         ldd=None, bias=None, residual=None, ldr=None, preact=None, gelu_grad_of=None, epilogue=0, split_k=1,
         block_n=0, dropout_p=0.0, seed=0, offset=0, colsum=None, dyn_m=None, dyn_k=None, dyn_hint=0):
    """D[M,N] = epilogue(A * B).  a/b bf16, d bf16 or fp32.  See include/ergm_b200.h."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    args = L.GemmArgs()
    args.a, args.b, args.d = a.data_ptr(), b.data_ptr(), d.data_ptr()
    args.bias = bias.data_ptr() if bias is not None else None
    args.residual = residual.data_ptr() if residual is not None else None
    args.preact = preact.data_ptr() if preact is not None else None
    args.colsum = colsum.data_ptr() if colsum is not None else None
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
    if gelu_grad_of is not None:  # multiply by gelu_new'(saved pre-activation)
        args.preact = gelu_grad_of.data_ptr()
        epilogue |= L.EPI_GELU_GRAD
    if dropout_p > 0.0:
        epilogue |= L.EPI_DROPOUT
    args.epilogue = epilogue
    args.split_k = split_k
    args.block_n = block_n
    args.dropout_p = dropout_p
    args.seed, args.offset = seed, offset
    if dyn_m is not None or dyn_k is not None:  # device int32 count bounding M or K at run time (see header)
        dyn = dyn_m if dyn_m is not None else dyn_k
        assert dyn.dtype == torch.int32
        args.dyn_count, args.dyn_dim = dyn.data_ptr(), 1 if dyn_m is not None else 2
        args.dyn_hint = int(dyn_hint or DYN_HINT)
    global _launch_count
    _launch_count += 1
    if PROFILE is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.lib().ergm_gemm_bf16(ctypes.byref(args), _stream()), "ergm_gemm_bf16")
        e1.record()
        # dyn: 0 static extents, 1 / 2 = M / K bounded at run time by a device count (algorithmic FLOPs need the count)
        PROFILE.append(("ergm_gemm_bf16", (M, N, K, a_major, b_major, split_k, args.dyn_dim if args.dyn_count else 0), e0, e1))
        return
    L.check(L.lib().ergm_gemm_bf16(ctypes.byref(args), _stream()), "ergm_gemm_bf16")


# kernels launched per C-ABI call (for bench.py's gpu_launches claim)
_LAUNCHES = {"ergm_attn_bwd": 2,   # delta + the persistent kernel (+ memset and cast for Tq > 256)
             "ergm_decode_layers": 1, "ergm_decode_stack_pack": 4, "ergm_lmhead_ce_fwd": 4, "ergm_lmhead_ce_bwd": 4}
_launch_count = 0
DYN_HINT = 0    # expected run-time row count of the packed batch being processed (engine sets it; steers tile shapes)
PROFILE = None  # set to a list to collect (name, info, start_event, end_event) per call


def reset_launch_count():
    global _launch_count
    _launch_count = 0


def launch_count():
    return _launch_count


def _call(name, *args):
    global _launch_count
    fn = getattr(L.lib(), name)
    _launch_count += _LAUNCHES.get(name, 1)
    if PROFILE is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(fn(*args, torch.cuda.current_stream().cuda_stream), name)
        e1.record()
        PROFILE.append((name, None, e0, e1))
        return
    L.check(fn(*args, torch.cuda.current_stream().cuda_stream), name)


def _p(t):
    return None if t is None else t.data_ptr()


_err_flags = {}


def err_flag(device):
    """Per-device int32 flag the kernels raise on out-of-range indices (checked lazily)."""
    key = (device.type, device.index)
    if key not in _err_flags:
        _err_flags[key] = torch.zeros(1, dtype=torch.int32, device=device)
    return _err_flags[key]


def check_err_flag(device):
    """Synchronous check (also drops any asynchronous read that is still pending for this device)."""
    f = err_flag(device)
    _err_pending.pop((device.type, device.index), None)
    if int(f.item()) != 0:
        f.zero_()
        raise IndexError("ergm_b200: index out of range in input_ids / token_type_ids / labels")


_err_pending = {}


def poll_err_flag(device):
    """Default (cheap) index check: every forward enqueues a 4-byte async D2H read of the flag; the NEXT forward
    (or an explicit check) raises if an earlier one saw an out-of-range index.  No host synchronisation is added:
    a read that has not completed yet is simply looked at later.  ERGM_CHECK_INDICES=1 checks synchronously."""
    if torch.cuda.is_current_stream_capturing():
        return
    key = (device.type, device.index)
    pend = _err_pending.get(key)
    if pend is not None:
        host, ev = pend
        if not ev.query():
            return  # still in flight: look at it later, do not stack another read
        if int(host[0]) != 0:
            _err_pending.pop(key)
            err_flag(device).zero_()
            raise IndexError("ergm_b200: index out of range in input_ids / token_type_ids / position_ids / caption_ids "
                             "/ labels of an earlier forward (the reference raises IndexError there); did you call "
                             "resize_token_embeddings for the special tokens?")
    else:
        host = torch.zeros(1, dtype=torch.int32).pin_memory()
    host.copy_(err_flag(device), non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    _err_pending[key] = (host, ev)


class Pack:
    """Packed variable-length batch: device-side row layout made by ergm_pack_plan from the per-sample
    token counts.  All sizes are static (capacity B*T); the packed row count lives on the device (`n_rows`), so one
    CUDA graph serves every batch.  `desc` is the ergm_pack the C entry points take."""

    def __init__(self, B, T, device):
        i32 = torch.int32
        self.B, self.T, self.cap = B, T, B * T
        self.cu = torch.zeros(B + 1, dtype=i32, device=device)
        self.row_b = torch.zeros(self.cap, dtype=i32, device=device)
        self.row_t = torch.zeros(self.cap, dtype=i32, device=device)
        self.n_rows = torch.zeros(1, dtype=i32, device=device)
        self.kv_lens = torch.zeros(B, dtype=i32, device=device)
        self.desc = L.PackDesc(self.cu.data_ptr(), self.row_b.data_ptr(), self.row_t.data_ptr(), self.n_rows.data_ptr(),
                               self.kv_lens.data_ptr())

    def plan(self, lens):
        """lens: device int32 [B] real tokens per sample (right-padded batch)."""
        _call("ergm_pack_plan", lens.data_ptr(), self.B, self.T, self.cu.data_ptr(), self.row_b.data_ptr(),
              self.row_t.data_ptr(), self.n_rows.data_ptr(), self.kv_lens.data_ptr(), self.cap)
        return self


def _pk(pack):
    return ctypes.byref(pack.desc) if pack is not None else None


def zero_rows_dyn(buf, count):
    """Rows [count, roundup(count, 128)) of a 2-D buffer := 0 (operands of run-time-K GEMMs)."""
    _call("ergm_zero_rows_dyn", buf.data_ptr(), buf.stride(0) * buf.element_size(), count.data_ptr(), buf.shape[0])


def _pos_stride(pos_ids, T):
    return 0 if (pos_ids is None or pos_ids.numel() == T) else T


def embed_fuse_fwd(ids, tts, pos_ids, wte, wpe, imgs, auds, out, *, past_len=0, past_lens=None, dropout_p=0.0, seed=0,
                   offset=0, pack=None):
    """pos_ids: None, [T] (shared by the batch) or [B, T] (per sample)."""
    B, T = ids.shape
    H = wte.shape[1]
    _call("ergm_embed_fuse_fwd", ids.data_ptr(), _p(tts), _p(pos_ids), _pos_stride(pos_ids, T), _p(past_lens),
          wte.data_ptr(), wpe.data_ptr(),
          _p(imgs), imgs.stride(0) if imgs is not None else 0, _p(auds), auds.stride(0) if auds is not None else 0,
          out.data_ptr(), B, T, H, past_len, wte.shape[0], wpe.shape[0], dropout_p, seed, offset,
          err_flag(ids.device).data_ptr(), _pk(pack))


def gather_rows_bf16(ids, table, out):
    _call("ergm_gather_rows_bf16", ids.data_ptr(), table.data_ptr(), out.data_ptr(), ids.numel(), table.shape[1],
          table.shape[0], err_flag(ids.device).data_ptr())


def mm_pool_fwd(seq, pooled_f32, pooled_bf16, lens=None):
    """seq fp32 [B, T, D] (last dim contiguous) -> time mean [B, D] in fp32 and/or bf16."""
    B, T, D = seq.shape
    assert seq.dtype == torch.float32 and seq.stride(2) == 1
    _call("ergm_mm_pool_fwd", seq.data_ptr(), seq.stride(0), seq.stride(1), _p(lens), B, T, D, _p(pooled_f32),
          pooled_f32.stride(0) if pooled_f32 is not None else 0, _p(pooled_bf16),
          pooled_bf16.stride(0) if pooled_bf16 is not None else 0)


def embed_bwd(dh, ids, tts, pos_ids, dwte, dwpe, *, T, past_len=0, dimgs=None, dauds=None, dropout_p=0.0, seed=0, offset=0,
              pack=None):
    """dwte [vocab, H] / dwpe [n_pos, H] gradient tables: out-of-range indices are skipped and flagged."""
    rows, H = dh.shape
    _call("ergm_embed_bwd", dh.data_ptr(), _p(ids), _p(tts), _p(pos_ids), _pos_stride(pos_ids, T), _p(dwte), _p(dwpe), _p(dimgs), _p(dauds),
          rows, T, H, past_len, dwte.shape[0] if dwte is not None else 0, dwpe.shape[0] if dwpe is not None else 0,
          dropout_p, seed, offset, err_flag(dh.device).data_ptr(), _pk(pack))


def ln_fwd(x, gamma, beta, y_bf16, y_f32, mean, rstd, eps, row_idx=None, rows_dyn=None):
    rows, H = x.shape
    if row_idx is not None:
        rows = row_idx.numel()
    _call("ergm_ln_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _p(y_bf16), _p(y_f32), _p(mean), _p(rstd),
          rows, H, eps, _p(row_idx), _p(rows_dyn))


def ln_bwd(dy, x, mean, rstd, gamma, dres_in, dx_out, dx_bf16, dgamma, dbeta, dbias_next=None, *, dropout_p=0.0,
           seed=0, offset=0, rows_dyn=None):
    rows, H = x.shape
    _call("ergm_ln_bwd", dy.data_ptr(), int(dy.dtype == torch.float32), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
          gamma.data_ptr(), _p(dres_in), _p(dx_out), _p(dx_bf16), dgamma.data_ptr(), dbeta.data_ptr(), _p(dbias_next),
          rows, H, dropout_p, seed, offset, _p(rows_dyn))


def colsum_bf16(src, out, *, rows=None, N=None, ld=None):
    rows = src.shape[0] if rows is None else rows
    N = src.shape[1] if N is None else N
    _call("ergm_colsum_bf16", src.data_ptr(), src.stride(0) if ld is None else ld, rows, N, out.data_ptr())


def gelu_bwd_colsum(dg, u, colsum, exact=False, rows_dyn=None):
    rows, N = dg.shape
    _call("ergm_gelu_bwd_colsum", dg.data_ptr(), u.data_ptr(), dg.stride(0), rows, N, _p(colsum), int(exact), _p(rows_dyn))


def cast_f32_bf16_2d(src, dst, colsum=None, rows_dyn=None):
    rows, N = src.shape
    _call("ergm_cast_f32_bf16_2d", src.data_ptr(), src.stride(0), dst.data_ptr(), dst.stride(0), rows, N, _p(colsum),
          _p(rows_dyn))


def cast_f32_bf16(src, dst):
    _call("ergm_cast_f32_bf16", src.data_ptr(), dst.data_ptr(), src.numel())


def attn_fwd(q, k, v, out, lse, *, B, nh, Tq, Tk, q_col0=0, k_col0=0, v_col0=0, causal=True, causal_off=None,
             kv_lens=None, dropout_p=0.0, seed=0, offset=0, out_f32=None, pack=None, pack_kv=False):
    """q/k/v: bf16 2-D matrices [B*T, ld] (may be the same tensor with different col0).  pack: queries / outputs are
    packed rows (sample b at cu[b] ..); pack_kv: keys / values too (self attention)."""
    if causal_off is None:
        causal_off = Tk - Tq
    _call("ergm_attn_fwd", q.data_ptr(), q.stride(0), q_col0, k.data_ptr(), k.stride(0), k_col0, v.data_ptr(),
          v.stride(0), v_col0, out.data_ptr(), out.stride(0), _p(out_f32), _p(lse), _p(kv_lens), B, nh, Tq, Tk, 64,
          int(causal),
          causal_off, dropout_p, seed, offset, _pk(pack), int(pack_kv))


@functools.lru_cache(maxsize=64)
def attn_bwd_workspace_bytes(B, nh, Tq):
    import ctypes as C
    n = C.c_int64()
    L.check(L.lib().ergm_attn_bwd_workspace_bytes(B, nh, Tq, C.byref(n)), "ergm_attn_bwd_workspace_bytes")
    return n.value


def attn_bwd(q, k, v, out, dout, lse, delta, dq, dk, dv, *, B, nh, Tq, Tk, q_col0=0, k_col0=0, v_col0=0,
             dq_col0=0, dk_col0=0, dv_col0=0, causal=True, causal_off=None, kv_lens=None, dropout_p=0.0, seed=0,
             offset=0, out_f32=None, dq_colsum=None, dk_colsum=None, dv_colsum=None, pack=None, pack_kv=False,
             workspace=None):
    """dq / dk / dv: bf16 outputs.  workspace: uint8 / fp32 tensor of attn_bwd_workspace_bytes(B, nh, Tq) bytes (0 for
    Tq <= 256); allocated here when not given (tests) - the engine passes a cached one."""
    if causal_off is None:
        causal_off = Tk - Tq
    assert dq.dtype == torch.bfloat16
    need = attn_bwd_workspace_bytes(B, nh, Tq)
    if need and (workspace is None or workspace.numel() * workspace.element_size() < need):
        workspace = torch.empty(need, dtype=torch.uint8, device=q.device)
    _call("ergm_attn_bwd", q.data_ptr(), q.stride(0), q_col0, k.data_ptr(), k.stride(0), k_col0, v.data_ptr(),
          v.stride(0), v_col0, out.data_ptr(), out.stride(0), _p(out_f32), dout.data_ptr(), dout.stride(0), lse.data_ptr(),
          delta.data_ptr(), dq.data_ptr(), dq.stride(0), dq_col0, dk.data_ptr(), dk.stride(0), dk_col0,
          dv.data_ptr(), dv.stride(0), dv_col0, _p(dq_colsum), _p(dk_colsum), _p(dv_colsum), _p(kv_lens), B, nh, Tq, Tk,
          64, int(causal), causal_off, dropout_p, seed, offset, _pk(pack), int(pack_kv),
          workspace.data_ptr() if need else None, need)


def ce_fwd(logits, labels, lse, row_loss, sums, *, T, V, hn=None, w=None, rows_dyn=None):
    """T = 0: labels already aligned with the (compacted) rows; rows_dyn: device int32 run-time row count."""
    rows = logits.shape[0]
    _call("ergm_ce_fwd", logits.data_ptr(), int(logits.dtype == torch.float32), logits.stride(0), labels.data_ptr(),
          rows, T, V, lse.data_ptr(), row_loss.data_ptr(), sums.data_ptr(), err_flag(logits.device).data_ptr(),
          _p(hn), _p(w), hn.shape[1] if hn is not None else 0, _p(rows_dyn))


def ce_bwd(logits, labels, lse, scale, dlogits, *, T, V, rows_dyn=None):
    rows = logits.shape[0]
    _call("ergm_ce_bwd", logits.data_ptr(), int(logits.dtype == torch.float32), logits.stride(0), labels.data_ptr(),
          rows, T, V, lse.data_ptr(), scale.data_ptr(), dlogits.data_ptr(), dlogits.stride(0), _p(rows_dyn))


def lm_rows_plan(labels, row_idx, labels_c, count, *, T, pack=None):
    _call("ergm_lm_rows_plan", labels.data_ptr(), labels.numel(), T, row_idx.data_ptr(), labels_c.data_ptr(),
          count.data_ptr(), _pk(pack))


def gather_rows_dyn(src, row_idx, count, dst):
    _call("ergm_gather_rows_dyn", src.data_ptr(), row_idx.data_ptr(), count.data_ptr(), dst.data_ptr(), src.shape[1],
          dst.shape[0])


def scatter_rows_dyn(src, row_idx, count, dst):
    _call("ergm_scatter_rows_dyn", src.data_ptr(), row_idx.data_ptr(), count.data_ptr(), dst.data_ptr(), src.shape[1])


_LMHEAD_WS = ("count", "row_idx", "labels_c", "hn_c", "logits_c", "lse", "row_loss", "dlogits_c", "dhn_c")
_LMHEAD_DT = (torch.int32, torch.int32, torch.int64, torch.bfloat16, torch.bfloat16, torch.float32, torch.float32,
              torch.bfloat16, torch.float32)


@functools.lru_cache(maxsize=64)
def lmhead_ce_layout(rows, H, V, with_backward=True):
    """Byte offsets of the sub-buffers of the ergm_lmhead_ce_* workspace (last entry = total bytes)."""
    import ctypes as C
    off = (C.c_int64 * (len(_LMHEAD_WS) + 1))()
    L.check(L.lib().ergm_lmhead_ce_workspace_layout(rows, H, V, int(with_backward), off), "ergm_lmhead_ce_workspace_layout")
    total = C.c_int64()
    L.check(L.lib().ergm_lmhead_ce_workspace_bytes(rows, H, V, int(with_backward), C.byref(total)),
            "ergm_lmhead_ce_workspace_bytes")
    assert total.value == off[len(_LMHEAD_WS)]
    return list(off)


def lmhead_ce_views(buf, rows, H, V, with_backward=True):
    """Typed views into a uint8 workspace tensor: {count, row_idx, labels_c, hn_c, logits_c, lse, row_loss, ...}."""
    off = lmhead_ce_layout(rows, H, V, with_backward)
    ldl = (V + 63) // 64 * 64
    shapes = ((1,), (rows,), (rows,), (rows, H), (rows, ldl), (rows,), (rows,), (rows, ldl), (rows, H))
    views = {"ws": buf}
    for i, name in enumerate(_LMHEAD_WS):
        if i >= 7 and not with_backward:
            break
        n = 1
        for s in shapes[i]:
            n *= s
        nbytes = n * torch.empty(0, dtype=_LMHEAD_DT[i]).element_size()
        views[name] = buf[off[i]:off[i] + nbytes].view(_LMHEAD_DT[i]).view(shapes[i])
    return views


def lmhead_ce_fwd(hn, wte_b, labels, sums, ws_buf, *, T, V, pack=None):
    """LM head + shifted CE on the scored rows, one C-ABI call (include/ergm_b200.h; model.py:698-708)."""
    rows, H = hn.shape
    _call("ergm_lmhead_ce_fwd", hn.data_ptr(), wte_b.data_ptr(), labels.data_ptr(), rows, T, H, V, sums.data_ptr(),
          err_flag(hn.device).data_ptr(), _pk(pack), ws_buf.data_ptr(), ws_buf.numel())


def lmhead_ce_bwd(wte_b, scale, dhn, dwte, ws_buf, *, V):
    rows, H = dhn.shape
    _call("ergm_lmhead_ce_bwd", wte_b.data_ptr(), scale.data_ptr(), rows, H, V, dhn.data_ptr(), dwte.data_ptr(),
          ws_buf.data_ptr(), ws_buf.numel())


def emotion_head_fwd(x_final, mean, rstd, gamma, beta, w_emo, emo_labels, hlast, logits, dlogits, sums, *, B, T, cu_rows=None):
    H = x_final.shape[1]
    _call("ergm_emotion_head_fwd", x_final.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
          beta.data_ptr(), w_emo.data_ptr(), _p(emo_labels), B, T, H, hlast.data_ptr(), logits.data_ptr(),
          _p(dlogits), _p(sums), err_flag(x_final.device).data_ptr(), _p(cu_rows))


def emotion_head_bwd(dlogits, hlast, w_emo, scale, dw_emo, dyf, *, B, T, cu_rows=None):
    H = hlast.shape[1]
    _call("ergm_emotion_head_bwd", dlogits.data_ptr(), hlast.data_ptr(), w_emo.data_ptr(), scale.data_ptr(), B, T, H,
          _p(dw_emo), _p(dyf), _p(cu_rows))


def loss_finalize(sums, has_lm, has_emo, out):
    _call("ergm_loss_finalize", sums.data_ptr(), int(has_lm), int(has_emo), out.data_ptr())


def scalar_mul(a, b, dst):
    _call("ergm_scalar_mul", a.data_ptr(), b.data_ptr(), dst.data_ptr())


def adamw_flat(p, g, m, v, shadow, hyper, grad_scale=None):
    _call("ergm_adamw_flat", p.data_ptr(), g.data_ptr(), int(g.dtype == torch.bfloat16), m.data_ptr(), v.data_ptr(),
          _p(shadow), p.numel(), hyper.data_ptr(), _p(grad_scale))


def attn_decode_paged(qkv, pool, block_table, seq_lens, out, *, B, nh, H):
    _call("ergm_attn_decode_paged", qkv.data_ptr(), qkv.stride(0), 0, H, 2 * H, pool.data_ptr(),
          block_table.data_ptr(), seq_lens.data_ptr(), block_table.shape[1], out.data_ptr(), out.stride(0), B, nh, 64)


def attn_decode_contig(q, kv, out, *, B, nh, Tk, k_col0, v_col0, kv_lens=None, q_col0=0):
    _call("ergm_attn_decode_contig", q.data_ptr(), q.stride(0), q_col0, kv.data_ptr(), kv.stride(0), k_col0, v_col0,
          _p(kv_lens), out.data_ptr(), out.stride(0), B, nh, Tk, 64)


def dec_pack_weight(w_f32, K, N, w_is_nk=False, gamma=None, beta=None, bias=None):
    """Packs an fp32 weight (logical [K, N]) into the bf16 decode slab layout, folding a preceding
    LayerNorm's gamma / beta in (see header).  Returns (packed, folded_bias or None)."""
    assert w_f32.dtype == torch.float32
    packed = torch.empty((N + 15) // 16 * 16 * K, dtype=torch.bfloat16, device=w_f32.device)
    bias_out = torch.empty(N, dtype=torch.float32, device=w_f32.device) if (beta is not None or bias is not None) else None
    _call("ergm_dec_pack_weight", w_f32.data_ptr(), w_f32.stride(0), K, N, int(w_is_nk), _p(gamma), _p(beta), _p(bias),
          packed.data_ptr(), _p(bias_out))
    return packed, bias_out


def dec_gemm(out, w_packed, *, M, K, N, x=None, a=None, eps=1e-5, bias=None, out_mode=0, gelu=False):
    """out[M,N] = epi(norm(x) @ Wp) or epi(a @ Wp); out_mode 0 bf16 store / 1 fp32 store / 2 fp32 += (see header)."""
    src = x if x is not None else a
    _call("ergm_dec_gemm", _p(x), _p(a), src.stride(0), float(eps), w_packed.data_ptr(), K, N,
          _p(bias), out.data_ptr(), out.stride(0), out_mode, int(gelu), M)


def decode_layers(table, *, L, H, I, nh, B, x, qkv, ctx, q2, g, block_table, seq_lens, Tc, eps, sync_ctr):
    """All transformer blocks of one decode step in one persistent kernel (see header)."""
    _call("ergm_decode_layers", table.data_ptr(), L, H, I, nh, B, x.data_ptr(), qkv.data_ptr(), ctx.data_ptr(),
          _p(q2), g.data_ptr(), block_table.data_ptr(), seq_lens.data_ptr(), block_table.shape[1], Tc, float(eps),
          sync_ctr.data_ptr())


def decode_stack_pack(w_qkv, gamma1, w_o, w_fc, gamma2, w_p2, *, H, I, nh):
    """Per-(cluster, CTA) weight blobs of one block for ergm_decode_stack -> (p1, fc, pj) bf16 tensors."""
    import ctypes as C
    sizes = [C.c_int64(0) for _ in range(3)]
    L.check(L.lib().ergm_decode_stack_blob_bytes(H, I, nh, *[C.byref(s) for s in sizes]), "ergm_decode_stack_blob_bytes")
    blobs = [torch.empty(s.value // 2, dtype=torch.bfloat16, device=w_qkv.device) for s in sizes]
    _call("ergm_decode_stack_pack", w_qkv.data_ptr(), gamma1.data_ptr(), w_o.data_ptr(), w_fc.data_ptr(), gamma2.data_ptr(),
          w_p2.data_ptr(), H, I, nh, *[b.data_ptr() for b in blobs])
    return blobs


def decode_stack_supported(H, I, nh):
    import ctypes as C
    s = [C.c_int64(0) for _ in range(3)]
    return L.lib().ergm_decode_stack_blob_bytes(H, I, nh, *[C.byref(x) for x in s]) == 0


def decode_stack(table, *, L, H, I, nh, B, xring, block_table, seq_lens, eps, sync_ctr, trace=None):
    """All transformer blocks of one decode step in one persistent cluster kernel (see header); the result is in
    xring[(2 * L) % 3]."""
    _call("ergm_decode_stack", table.data_ptr(), L, H, I, nh, B, xring[0].data_ptr(), xring[1].data_ptr(),
          xring[2].data_ptr(), block_table.data_ptr(), seq_lens.data_ptr(), block_table.shape[1], float(eps),
          sync_ctr.data_ptr(), _p(trace))


def kv_to_pages(kv, pool, block_table, lens, *, B, T, nh, k_col0, v_col0):
    _call("ergm_kv_to_pages", kv.data_ptr(), kv.stride(0), k_col0, v_col0, pool.data_ptr(), block_table.data_ptr(),
          _p(lens), block_table.shape[1], B, T, nh)


def sample(logits, *, V, top_k=0, top_p=1.0, temperature=1.0, seed=0, step=None, out_ids=None, next_ids=None,
           finished=None, seq_lens=None, eos_id=-1, advance_step=False):
    B = logits.shape[0]
    _call("ergm_sample", logits.data_ptr(), logits.stride(0), B, V, top_k, float(top_p), float(temperature), seed, _p(step),
          int(advance_step), _p(out_ids), out_ids.stride(0) if out_ids is not None else 0, _p(next_ids), _p(finished), _p(seq_lens),
          eos_id)


def int_add(t, inc):
    _call("ergm_int_add", t.data_ptr(), inc)


def rng_step_advance(t, inc=1):
    _call("ergm_rng_step_advance", t.data_ptr(), inc)


def split3_expand(src, dst, *, side, kdim):
    rows, cols = src.shape
    _call("ergm_split3_expand", src.data_ptr(), src.stride(0), dst.data_ptr(), dst.stride(0), rows, cols, side, kdim)


def attn_fwd_f32(q, k, v, out, *, B, nh, Tq, Tk, q_col0=0, k_col0=0, v_col0=0, causal=True, causal_off=None,
                 kv_lens=None):
    if causal_off is None:
        causal_off = Tk - Tq
    _call("ergm_attn_fwd_f32", q.data_ptr(), q.stride(0), q_col0, k.data_ptr(), k.stride(0), k_col0, v.data_ptr(),
          v.stride(0), v_col0, out.data_ptr(), out.stride(0), _p(kv_lens), B, nh, Tq, Tk, 64, int(causal), causal_off)

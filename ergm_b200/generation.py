"""KV-cached response generation for the drop-in model (BASELINE config 4).

The reference decodes with one FULL forward per new token (main.py:253-282) and its model-side
cache grows by torch.cat (model.py:228-236).  Here a request batch is prefetched once (fused
attention over the right-padded prompts, K/V scattered into a paged pool, cross-attention K/V
of the captions projected once), then every new token is ONE CUDA-graph replay of the decode
step: embed -> L x [LN, QKV GEMM, paged one-query attention (+append), proj GEMM, (cross), MLP]
-> ln_f -> LM head on B rows -> on-device arg-max / top-k sampling.  No host synchronisation
happens until the generated ids are read back.
"""
import os

import torch

from . import _lib as L
from . import ops
from .engine import K_MAJOR

PAGE = 16


class KVPageAllocator:
    """Page pool of the self-attention K/V cache of one engine: per layer one tensor [pages][k|v][head][16 tokens][64]
    (bf16), a host-side free list, pages handed out page-index-major (page j of every sequence of a batch before
    page j+1: the order in which sequences that grow together claim them), so a sequence's pages are NOT
    contiguous and the kernels' block-table indirection is what finds them.  The pool only grows; pages of a
    finished batch go back on the free list (most recently freed first)."""

    def __init__(self, eng):
        self.eng = eng
        self.pools = None
        self.n_pages = 0
        self.free = []

    def _grow(self, need):
        """Not enough free pages: first the cached (idle) generation states give theirs back; if that is still not
        enough a larger pool replaces this one.  States that are still in use (handed out with return_state) keep
        the old tensors alive and go on working in them; their page ids stay reserved in the new pool."""
        eng = self.eng
        for _, st in eng.__dict__.get("_gen_states", []):
            st.graph = None
            st.release()
        eng.__dict__["_gen_states"] = []
        if len(self.free) >= need:
            return
        new_total = self.n_pages + need - len(self.free)
        self.pools = [torch.empty(new_total, 2, eng.nh, PAGE, 64, dtype=torch.bfloat16, device=eng.device)
                      for _ in range(eng.L)]   # never read before written: every slot < seq_len was stored first
        self.free = list(range(new_total - 1, self.n_pages - 1, -1)) + self.free
        self.n_pages = new_total

    def alloc_table(self, B, pages_per_seq):
        need = B * pages_per_seq
        if len(self.free) < need:
            self._grow(need)
        ids = [self.free.pop() for _ in range(need)]
        table = torch.tensor(ids, dtype=torch.int32).view(pages_per_seq, B).t().contiguous()   # page j of all sequences first
        return table.to(self.eng.device), ids

    def release(self, ids):
        self.free.extend(reversed(ids))


def page_allocator(eng):
    a = eng.__dict__.get("_kv_pages")
    if a is None:
        a = eng.__dict__["_kv_pages"] = KVPageAllocator(eng)
    return a


class GenState:
    """Device state of one generation batch: its pages of the engine's K/V page pool + block table, cached
    cross-attention K/V, per-sequence lengths / finished flags, output ids, and the captured decode graph."""

    def __init__(self, eng, B, max_ctx, Tc, max_new):
        dev = eng.device
        H, nh, Lyr = eng.H, eng.nh, eng.L
        self.B, self.max_ctx, self.Tc, self.max_new = B, max_ctx, Tc, max_new
        self.pages_per_seq = (max_ctx + PAGE - 1) // PAGE
        self.alloc = page_allocator(eng)
        self.block_table, self.page_ids = self.alloc.alloc_table(B, self.pages_per_seq)
        self.pool = self.alloc.pools
        self.kv2 = [torch.empty(B * Tc, 2 * H, dtype=torch.bfloat16, device=dev) for _ in range(Lyr)] if Tc else None
        self.seq_lens = torch.zeros(B, dtype=torch.int32, device=dev)
        self.finished = torch.zeros(B, dtype=torch.int32, device=dev)
        self.next_ids = torch.zeros(B, 1, dtype=torch.int64, device=dev)
        self.out_ids = torch.zeros(B, max_new, dtype=torch.int64, device=dev)
        self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.tt = None
        self.logits = torch.empty(B, (eng.V + 63) // 64 * 64, dtype=torch.float32, device=dev)
        self.graph = None

    def kv_bytes(self):
        page = self.pool[0][0].numel() * 2
        return (page * self.B * self.pages_per_seq * len(self.pool)
                + (sum(k.numel() * 2 for k in self.kv2) if self.kv2 else 0))

    def release(self):
        """Hands the pages back (the state must not be used afterwards)."""
        if self.page_ids is not None:
            self.alloc.release(self.page_ids)
            self.page_ids = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def reset(self):
        self.finished.zero_()
        self.step.zero_()
        self.out_ids.zero_()


STATE_CACHE_SIZE = 2   # generation states (pages + captured decode graph) kept per engine, least recently used evicted


def _cached_state(eng, key):
    cache = eng.__dict__.setdefault("_gen_states", [])
    for i, (k, st) in enumerate(cache):
        if k == key and st.pool is page_allocator(eng).pools:
            cache.append(cache.pop(i))
            return st
    return None


def _store_state(eng, key, st):
    cache = eng.__dict__.setdefault("_gen_states", [])
    cache.append((key, st))
    while len(cache) > STATE_CACHE_SIZE:
        _, old = cache.pop(0)
        old.graph = None
        old.release()


def _head_on_rows(eng, x, row_idx, logits):
    """ln_f + tied LM head on selected rows (fp32 logits so the arg-max is not bf16-rounded)."""
    ws = eng.ws_eval
    n = row_idx.numel() if row_idx is not None else x.shape[0]
    hn = ws.get("gen_hn", (n, eng.H), torch.bfloat16)
    ops.ln_fwd(x, eng.p("transformer.ln_f.weight"), eng.p("transformer.ln_f.bias"), hn, None, None, None,
               eng.cfg.layer_norm_epsilon, row_idx=row_idx)
    ops.gemm(hn, eng.pb("transformer.wte.weight"), logits, M=n, N=eng.V, K=eng.H, a_major=K_MAJOR, b_major=K_MAJOR)


DEC_TILE = 64  # rows per ergm_dec_gemm call


def packed_weights(eng):
    """Decode-layout copies of the weights (bf16 16-column slabs in mma fragment order with the
    preceding LayerNorm folded in, ergm_dec_pack_weight), cached per parameter version: generation
    never re-packs unless the weights changed.  Values: (packed, bias)."""
    eng.store.refresh_shadow()  # notices in-place edits of the parameters (bumps weights_epoch)
    ver = eng.store.weights_epoch
    hit = eng.__dict__.get("_dec_pack")
    if hit is not None and hit[0] == ver and hit[1] is eng.store.flat:
        return hit[2]
    H, I, V = eng.H, eng.I, eng.V
    pk = {}
    for l in range(eng.L):
        pfx = "transformer.h.%d." % l
        for name, K, N, ln in (("attn.c_attn", H, 3 * H, "ln_1"), ("attn.c_proj", H, H, None),
                               ("crossattention.q_attn", H, H, "ln_cross_attn"), ("crossattention.c_proj", H, H, None),
                               ("mlp.c_fc", H, I, "ln_2"), ("mlp.c_proj", I, H, None)):
            pk[pfx + name] = ops.dec_pack_weight(
                eng.p(pfx + name + ".weight"), K, N, gamma=eng.p(pfx + ln + ".weight") if ln else None,
                beta=eng.p(pfx + ln + ".bias") if ln else None, bias=eng.p(pfx + name + ".bias"))
    eng.__dict__["_dec_pack"] = (ver, eng.store.flat, pk)
    return pk


def mega_table(eng, st):
    """Device table of per-layer pointers for ergm_decode_layers (packed weights, folded biases, this
    batch's K/V pools): 14 pointers per layer."""
    pk = st.packed
    rows = []
    for l in range(eng.L):
        pfx = "transformer.h.%d." % l

        def pair(name):
            w, b = pk[pfx + name]
            return [w.data_ptr(), b.data_ptr() if b is not None else 0]

        cross = st.kv2 is not None
        rows.append(pair("attn.c_attn") + pair("attn.c_proj")
                    + (pair("crossattention.q_attn") if cross else [0, 0])
                    + (pair("crossattention.c_proj") if cross else [0, 0])
                    + pair("mlp.c_fc") + pair("mlp.c_proj")
                    + [st.pool[l].data_ptr(), st.kv2[l].data_ptr() if cross else 0])
    return torch.tensor(rows, dtype=torch.int64).to(eng.device)


def stack_weights(eng):
    """ergm_decode_stack's weight blobs (one contiguous fragment-packed blob per layer / phase / CTA), cached per
    parameter version next to the slab-packed weights (whose folded biases they share)."""
    pk = packed_weights(eng)
    ver = eng.store.weights_epoch
    hit = eng.__dict__.get("_dec_stack")
    if hit is not None and hit[0] == ver and hit[1] is eng.store.flat:
        return hit[2]
    out = []
    for l in range(eng.L):
        pfx = "transformer.h.%d." % l
        p1, fc, pj = ops.decode_stack_pack(eng.p(pfx + "attn.c_attn.weight"), eng.p(pfx + "ln_1.weight"),
                                           eng.p(pfx + "attn.c_proj.weight"), eng.p(pfx + "mlp.c_fc.weight"),
                                           eng.p(pfx + "ln_2.weight"), eng.p(pfx + "mlp.c_proj.weight"),
                                           H=eng.H, I=eng.I, nh=eng.nh)
        out.append((p1, fc, pj, pk[pfx + "attn.c_attn"][1], pk[pfx + "attn.c_proj"][1], pk[pfx + "mlp.c_fc"][1],
                    pk[pfx + "mlp.c_proj"][1]))
    eng.__dict__["_dec_stack"] = (ver, eng.store.flat, out)
    return out


def stack_supported(eng, B, Tc):
    """One persistent cluster kernel for all blocks of a decode step (csrc/decode_step.cu): no captions, B <= 64,
    GPT-2-small-class widths.  ERGM_DEC_STACK=0 selects the per-kernel chain."""
    return (os.environ.get("ERGM_DEC_STACK", "0") != "0" and Tc == 0 and B <= DEC_TILE
            and ops.decode_stack_supported(eng.H, eng.I, eng.nh))


def stack_table(eng, st):
    rows = []
    for l, (p1, fc, pj, b_qkv, b_o, b_fc, b_p2) in enumerate(st.stack_w):
        rows.append([p1.data_ptr(), fc.data_ptr(), pj.data_ptr(), b_qkv.data_ptr(), b_o.data_ptr(), b_fc.data_ptr(),
                     b_p2.data_ptr(), st.pool[l].data_ptr()])
    return torch.tensor(rows, dtype=torch.int64).to(eng.device)


def mega_supported(eng, B):
    # opt-in: measured 576-660 us / step against 533 us for the launch chain (profiles/r1_decode.md) - the
    # in-kernel attention phase (4 groups per CTA, two rounds) and the 60 grid barriers still cost more than
    # they save
    return (os.environ.get("ERGM_DEC_MEGA", "0") == "1" and B <= DEC_TILE and eng.H in (128, 256, 512, 768, 1024)
            and eng.I % 512 == 0)


def decode_step(eng, st, sample_kw):
    """One token for every sequence of the batch; pure device work (CUDA-graph capturable).
    Per layer: [LN1 + QKV] -> paged attention (+append) -> [out-proj += residual] -> ([LN + q] ->
    cross attention -> [out-proj +=]) -> [LN2 + FC + gelu] -> [MLP proj +=]; every bracket is one
    weight-streaming ergm_dec_gemm launch (programmatic dependent launch: its weight prefetch
    overlaps the previous kernel)."""
    H, nh, I, B, V = eng.H, eng.nh, eng.I, st.B, eng.V
    ws = eng.ws_eval
    f32, bf16 = torch.float32, torch.bfloat16
    eps = eng.cfg.layer_norm_epsilon
    pk = st.packed
    if getattr(st, "stack", None) is not None:
        # all blocks in ONE persistent cluster kernel: 2 grid-wide synchronisations per block instead of 5 dependent
        # launches (csrc/decode_step.cu); falls back to the chain below when the device refuses the geometry
        ops.embed_fuse_fwd(st.next_ids, st.tt, None, eng.p("transformer.wte.weight"), eng.p("transformer.wpe.weight"),
                           None, None, st.xring[0], past_lens=st.seq_lens)
        try:
            ops.decode_stack(st.stack, L=eng.L, H=H, I=I, nh=nh, B=B, xring=st.xring, block_table=st.block_table,
                             seq_lens=st.seq_lens, eps=eps, sync_ctr=st.sync_ctr, trace=getattr(st, "trace", None))
        except L.ErgmError as e:
            if "unsupported" not in str(e):
                raise
            st.stack = None   # this device cannot hold all clusters at once: per-kernel chain (nothing was launched)
            return decode_step(eng, st, sample_kw)
        _head_on_rows(eng, st.xring[(2 * eng.L) % 3], None, st.logits)
        ops.sample(st.logits, V=eng.V, step=st.step, out_ids=st.out_ids, next_ids=st.next_ids, finished=st.finished,
                   seq_lens=st.seq_lens, advance_step=True, **sample_kw)
        return
    x = ws.get("dec_x", (B, H), f32)
    ops.embed_fuse_fwd(st.next_ids, st.tt, None, eng.p("transformer.wte.weight"), eng.p("transformer.wpe.weight"),
                       None, None, x, past_lens=st.seq_lens)
    qkv = ws.get("dec_qkv", (B, 3 * H), bf16)
    ctx = ws.get("dec_ctx", (B, H), bf16)
    q2 = ws.get("dec_q2", (B, H), bf16)
    g = ws.get("dec_g", (B, I), bf16)
    if st.mega is not None:
        # all blocks in ONE persistent kernel (grid barriers instead of ~60 dependent launches)
        ops.decode_layers(st.mega, L=eng.L, H=H, I=I, nh=nh, B=B, x=x, qkv=qkv, ctx=ctx, q2=q2 if st.kv2 is not None else None,
                          g=g, block_table=st.block_table, seq_lens=st.seq_lens, Tc=st.Tc if st.kv2 is not None else 0,
                          eps=eps, sync_ctr=st.sync_ctr)
        _head_on_rows(eng, x, None, st.logits)
        ops.sample(st.logits, V=eng.V, step=st.step, out_ids=st.out_ids, next_ids=st.next_ids, finished=st.finished,
                   seq_lens=st.seq_lens, advance_step=True, **sample_kw)
        return
    tiles = [(r0, min(B, r0 + DEC_TILE)) for r0 in range(0, B, DEC_TILE)]

    def ln_gemm(out, name, K, N, **kw):
        w, b = pk[name]
        for r0, r1 in tiles:
            ops.dec_gemm(out[r0:r1], w, M=r1 - r0, K=K, N=N, x=x[r0:r1], eps=eps, bias=b, **kw)

    def res_gemm(a, name, K, N):
        w, b = pk[name]
        for r0, r1 in tiles:
            ops.dec_gemm(x[r0:r1], w, M=r1 - r0, K=K, N=N, a=a[r0:r1], bias=b, out_mode=2)

    for l in range(eng.L):
        pfx = "transformer.h.%d." % l
        ln_gemm(qkv, pfx + "attn.c_attn", H, 3 * H)
        ops.attn_decode_paged(qkv, st.pool[l], st.block_table, st.seq_lens, ctx, B=B, nh=nh, H=H)
        res_gemm(ctx, pfx + "attn.c_proj", H, H)
        if st.kv2 is not None:
            ln_gemm(q2, pfx + "crossattention.q_attn", H, H)
            ops.attn_decode_contig(q2, st.kv2[l], ctx, B=B, nh=nh, Tk=st.Tc, k_col0=0, v_col0=H)
            res_gemm(ctx, pfx + "crossattention.c_proj", H, H)
        ln_gemm(g, pfx + "mlp.c_fc", H, I, gelu=True)
        res_gemm(g, pfx + "mlp.c_proj", I, H)
    # LM head (77 MB of weights, N = 50260): the 128x256-tile tcgen05 kernel streams it in 21.8 us, the
    # slab kernel (3142 slabs, cross-warp reduction per slab) needs 40 us - measured, profiles/r1_decode.md
    _head_on_rows(eng, x, None, st.logits)
    ops.sample(st.logits, V=eng.V, step=st.step, out_ids=st.out_ids, next_ids=st.next_ids, finished=st.finished,
               seq_lens=st.seq_lens, advance_step=True, **sample_kw)


@torch.no_grad()
def generate(model, input_ids, token_type_ids=None, max_new_tokens=64, do_sample=False, top_k=0, top_p=1.0,
             temperature=1.0, eos_token_id=None, sp2_id=None, imgs=None, auds=None, caption_ids=None,
             prompt_lens=None, seed=0, use_cuda_graph=True, return_state=False):
    """input_ids [B, T] right-padded prompts, prompt_lens [B] their true lengths (default T).
    Every generated token gets speaker type sp2_id (main.py:277-279).  Returns int64 [B,
    max_new_tokens]; after a sequence emits eos it keeps emitting eos."""
    if getattr(model, "ergm_precision", "bf16") == "fp32":
        return _generate_fp32(model, input_ids, token_type_ids, max_new_tokens, eos_token_id, sp2_id, imgs, auds,
                              caption_ids, prompt_lens)
    if do_sample and top_p < 1.0 and top_k > 1:
        raise L.ErgmError("top_k and top_p are alternatives (main.py:253-282 uses top-p only)")
    eng = model.engine
    eng.ensure_params()
    eng.store.refresh_shadow()
    dev = eng.device
    B, T = input_ids.shape
    input_ids = input_ids.to(dev, torch.int64).contiguous()
    if token_type_ids is not None:
        token_type_ids = token_type_ids.to(dev, torch.int64).contiguous()
    if prompt_lens is None:
        lens = torch.full((B,), T, dtype=torch.int32, device=dev)
    else:
        lens = prompt_lens.to(dev, torch.int32).contiguous()
    cap = caption_ids.to(dev, torch.int64).reshape(B, -1).contiguous() if caption_ids is not None else None
    Tc = cap.shape[1] if cap is not None else 0
    max_ctx = T + max_new_tokens
    if max_ctx > eng.n_pos:
        raise ValueError("prompt + max_new_tokens = %d exceeds n_positions = %d" % (max_ctx, eng.n_pos))
    k = int(top_k) if do_sample else 0
    if do_sample and k == 0 and top_p >= 1.0:
        k = -1  # plain multinomial sampling over the whole distribution (ERGM_SAMPLE_ALL), never silently greedy
    sample_kw = dict(top_k=k, top_p=float(top_p) if do_sample else 1.0, temperature=float(temperature), seed=int(seed),
                     eos_id=int(eos_token_id) if eos_token_id is not None else -1)
    with torch.cuda.device(dev):
        packed = packed_weights(eng)
        # A request batch of the same geometry / sampling setup as an earlier one reuses that batch's state: its pages,
        # its static buffers and its CAPTURED decode graph (the first version re-allocated a zeroed pool and re-captured
        # the graph on every generate() call).  Keyed on the weight version: packed weights are part of the graph.
        key = (B, T, max_ctx, Tc, max_new_tokens, sp2_id, tuple(sorted(sample_kw.items())), bool(use_cuda_graph),
               eng.store.weights_epoch, os.environ.get("ERGM_DEC_STACK", "0"), os.environ.get("ERGM_DEC_MEGA", "0"),
               os.environ.get("ERGM_DS_TRACE", "0"))
        st = None if return_state else _cached_state(eng, key)   # a state handed to the caller is the caller's alone
        fresh = st is None
        if fresh:
            st = GenState(eng, B, max_ctx, Tc, max_new_tokens)
            if not return_state:
                _store_state(eng, key, st)
        else:
            st.reset()
        st.packed = packed
    with torch.cuda.device(dev):
        if not fresh:
            pass
        elif stack_supported(eng, B, Tc):
            st.mega = None
            st.stack_w = stack_weights(eng)
            st.xring = [torch.zeros(B, eng.H, dtype=torch.float32, device=dev) for _ in range(3)]
            st.sync_ctr = torch.zeros(1, dtype=torch.int32, device=dev)
            st.stack = stack_table(eng, st)
            # profiling aid: ERGM_DS_TRACE=1 records per-phase device timestamps of CTA 0 (scripts/decode_stack_check.py)
            st.trace = (torch.zeros(16 * eng.L, dtype=torch.int64, device=dev)
                        if os.environ.get("ERGM_DS_TRACE") == "1" else None)
        elif mega_supported(eng, B):
            st.stack = None
            st.sync_ctr = torch.zeros(1, dtype=torch.int32, device=dev)
            st.mega = mega_table(eng, st)
        else:
            st.mega = st.stack = None
        if fresh and sp2_id is not None:
            st.tt = torch.full((B, 1), int(sp2_id), dtype=torch.int64, device=dev)
        # ---- prefill: fused attention over the padded prompts, K/V -> pages, cross K/V cached ----
        out = eng.forward(input_ids, token_type_ids, None, None, imgs, auds, cap, None, kv_lens=lens,
                          training=False, save=False, heads=False, gen_state=st)
        st.seq_lens.copy_(lens)
        last = (torch.arange(B, device=dev, dtype=torch.int32) * T + lens - 1).to(torch.int32)
        _head_on_rows(eng, out["x_final"], last, st.logits)
        ops.sample(st.logits, V=eng.V, step=st.step, out_ids=st.out_ids, next_ids=st.next_ids,
                   finished=st.finished, seq_lens=None, **sample_kw)
        ops.int_add(st.step, 1)
        # ---- decode: one graph replay per token ----
        n_steps = max_new_tokens - 1
        if n_steps > 0:
            if use_cuda_graph:
                s = torch.cuda.Stream(device=dev)
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    # warm-up run outside capture is not possible without consuming a step, so the
                    # first token is decoded eagerly and the graph replays the remaining ones
                    n0 = ops.launch_count()
                    decode_step(eng, st, sample_kw)
                    st.launches_per_step = ops.launch_count() - n0
                torch.cuda.current_stream().wait_stream(s)
                if n_steps > 1:
                    if st.graph is None:
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            decode_step(eng, st, sample_kw)
                        st.graph = g
                    for _ in range(n_steps - 1):  # capture itself executes nothing
                        st.graph.replay()
            else:
                for _ in range(n_steps):
                    decode_step(eng, st, sample_kw)
    if return_state:
        return st.out_ids, st
    return st.out_ids.clone()   # the state (and its output buffer) is reused by the next batch of the same geometry


def _generate_fp32(model, input_ids, token_type_ids, max_new_tokens, eos_token_id, sp2_id, imgs, auds, caption_ids,
                   prompt_lens):
    """fp32-mode greedy decoding: full recompute per token exactly like main.py:255-279 (no KV cache,
    no bf16 anywhere), arg-max on device.  Uniform prompt lengths only."""
    if prompt_lens is not None and not bool((prompt_lens == input_ids.shape[1]).all()):
        raise L.ErgmError("fp32-mode generation needs uniform prompt lengths")
    eng = model.engine
    dev = eng.device
    ids = input_ids.to(dev, torch.int64)
    tt = token_type_ids.to(dev, torch.int64) if token_type_ids is not None else None
    B = ids.shape[0]
    out_ids = torch.zeros(B, max_new_tokens, dtype=torch.int64, device=dev)
    nxt = torch.zeros(B, dtype=torch.int64, device=dev)
    finished = torch.zeros(B, dtype=torch.int32, device=dev)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    cap = caption_ids.to(dev, torch.int64) if caption_ids is not None else None
    for t in range(max_new_tokens):
        T = ids.shape[1]
        o = eng.forward_fp32(ids.contiguous(), tt.contiguous() if tt is not None else None, None, None, imgs, auds, cap)
        last = o["logits"].view(B, T, -1)[:, T - 1].contiguous()
        ops.sample(last, V=eng.V, step=step, out_ids=out_ids, next_ids=nxt, finished=finished,
                   eos_id=int(eos_token_id) if eos_token_id is not None else -1)
        ops.int_add(step, 1)
        ids = torch.cat([ids, nxt[:, None]], 1)
        if tt is not None:
            tt = torch.cat([tt, torch.full((B, 1), int(sp2_id if sp2_id is not None else 0), dtype=torch.int64,
                                           device=dev)], 1)
    return out_ids


def legacy_past_to_rows(past_key_values, B, H, dev):
    """Legacy tuple cache L x (k, v) [B, nh, ctx, 64] (model.py:228-236) -> per layer bf16 [B, ctx, 2H] K|V rows."""
    past = []
    for k, v in past_key_values:
        ctx = k.shape[-2]
        k2 = k.to(dev).permute(0, 2, 1, 3).reshape(B, ctx, H)
        v2 = v.to(dev).permute(0, 2, 1, 3).reshape(B, ctx, H)
        past.append(torch.cat([k2, v2], dim=-1).to(torch.bfloat16).contiguous())
    return past


def legacy_cached_forward(model, input_ids, token_type_ids, pos, past_key_values, attention_mask, caption_ids,
                          use_cache, return_dict):
    """forward(..., past_key_values=tuple) — the reference's own cache surface (model.py:228-236,
    469-476), kept so that code driving the model token by token keeps working."""
    from .model import CausalLMOutputWithEmotionClassification
    eng = model.engine
    dev = eng.device
    B, T = input_ids.shape
    H, nh = eng.H, eng.nh
    past = legacy_past_to_rows(past_key_values, B, H, dev)
    ctx = past[0].shape[1]
    kv_lens = None
    if attention_mask is not None:
        kv_lens = model._kv_lens_from_mask(attention_mask.to(dev), B, ctx + T)
    out = eng.forward(input_ids, token_type_ids, None, None, None, None, caption_ids, pos, kv_lens=kv_lens,
                      training=False, save=False, want_logits=True, logits_fp32=model.fp32_logits, legacy_past=past)
    V = eng.V
    logits_buf = out["logits"]
    kv_bufs = out["kv_present"]

    def logits_fn():
        return logits_buf[:, :V].to(torch.float32).view(B, T, V)

    def past_fn():
        res = []
        for kvf in kv_bufs:
            tk = kvf.shape[1]
            k = kvf[:, :, :H].view(B, tk, nh, 64).permute(0, 2, 1, 3).float()
            v = kvf[:, :, H:].view(B, tk, nh, 64).permute(0, 2, 1, 3).float()
            res.append((k, v))
        return tuple(res)

    ret = CausalLMOutputWithEmotionClassification(loss=None, logits_fn=logits_fn,
                                                  emotion_logits=out["emotion_logits"].clone(),
                                                  past_fn=past_fn if use_cache else None)
    if not return_dict:
        return (ret.logits, ret.emotion_logits) + ((ret.past_key_values,) if use_cache else ())
    return ret

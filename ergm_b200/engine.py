"""Forward / backward composition of the ERGM hot path out of the C-ABI kernels.

This is the host-side runtime behind ergm_b200.model.GPT2LMHeadModel.forward: it mirrors the
data flow of /root/reference/src/model.py (GPT2Model.forward :420-596, GPT2Block.forward
:286-341, GPT2LMHeadModel.forward :654-737) but every arithmetic step is a call into
libergm_b200.so — PyTorch only owns the memory and the stream.  The backward pass is written
by hand (one monolithic pass in reverse layer order) so that residual-gradient adds, bias
gradients, dropout masks and bf16 operand casts are fused into the LayerNorm-backward and
GEMM epilogues instead of being separate autograd nodes.
"""
import os

import torch

from . import _lib as L
from . import ops
from .param_store import ParamStore

K_MAJOR, MN_MAJOR = L.ERGM_MAJOR_K, L.ERGM_MAJOR_MN
_RNG_KEEPALIVE = []  # device step counters ever handed to ergm_set_rng_step_ptr (see Engine.set_rng_step_tensor)


def _pad8(n):
    return (n + 7) // 8 * 8


class Workspace:
    """Named, shape-keyed device buffers, allocated once and reused every step (static
    addresses keep the whole step CUDA-graph capturable)."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}

    def get(self, name, shape, dtype, zero=False):
        key = (name, tuple(shape), dtype)
        t = self.bufs.get(key)
        if t is None:
            # always zero-initialised: with packed batches rows behind the run-time row count are never written, yet
            # the TMA tiles of the last row block read them (they must hold finite values, never stale NaN patterns)
            t = torch.zeros(shape, dtype=dtype, device=self.device)
            self.bufs[key] = t
        return t

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs.values())


class LayerParams:
    __slots__ = ("f", "b", "g")  # fp32 views, bf16 shadow views, grad views (dicts by short name)


class Engine:
    def __init__(self, model, prefix=""):
        self.model = model
        self.cfg = model.config
        self.store = ParamStore(model, prefix)
        self.device = self.store.device
        self.ws_train = Workspace(self.device)
        self.ws_eval = Workspace(self.device)
        self.saved = None
        self.forward_serial = 0
        self.seed = 0x5EED
        self.site_counter = 0
        self.rng_step = None
        self._wgrad_side = None   # set per backward: side stream for the weight-gradient GEMMs
        self._side_ev = None
        self._index()

    # ------------------------------------------------------------------
    def _index(self):
        st = self.store
        cfg = self.cfg
        self.H = cfg.n_embd
        self.nh = cfg.n_head
        self.L = cfg.n_layer
        self.I = cfg.n_inner if cfg.n_inner is not None else 4 * cfg.n_embd
        if self.H // self.nh != 64 or self.H % 128:
            raise L.ErgmError("ergm_b200 kernels need head_dim == 64 and n_embd %% 128 == 0 (got n_embd=%d, n_head=%d)"
                              % (self.H, self.nh))
        if getattr(cfg, "activation_function", "gelu_new") != "gelu_new":
            raise L.ErgmError("only activation_function='gelu_new' (model.py:259) is implemented")
        if getattr(cfg, "scale_attn_by_inverse_layer_idx", False) or getattr(cfg, "reorder_and_upcast_attn", False) \
                or not getattr(cfg, "scale_attn_weights", True):
            raise L.ErgmError("non-default attention scaling flags (model.py:122-128,150-188) are not implemented")
        names = set(st.entries)

        def trio(name):
            return st.view(name), st.shadow_view(name), st.grad_view(name)

        self.P = {}
        for name in names:
            self.P[name] = trio(name)
        self.V = st.entries["transformer.wte.weight"][2][0]
        self.n_pos = st.entries["transformer.wpe.weight"][2][0]

    def get_pack(self, B, T):
        """Device-side row layout of a packed [B, T] batch (cached per shape: static addresses for CUDA graphs)."""
        packs = self.__dict__.setdefault("_packs", {})
        pk = packs.get((B, T))
        if pk is None:
            pk = packs[(B, T)] = ops.Pack(B, T, self.device)
        return pk

    def ensure_params(self):
        if not self.store.valid():
            self.store.build()
            self._index()
            self.ws_train = Workspace(self.store.device)
            self.ws_eval = Workspace(self.store.device)
            self.device = self.store.device

    def p(self, name):
        return self.P[name][0]

    def pb(self, name):
        return self.P[name][1]

    def pg(self, name):
        return self.P[name][2]

    # ------------------------------------------------------------------
    def _new_sites(self, training):
        """Dropout site ids for one forward: (seed, base offset).  Offsets are unique per
        nn.Dropout application; the optional device step counter de-correlates graph replays."""
        self.site_counter += 1
        return self.site_counter * 4096

    def set_rng_step_tensor(self, t):
        """Registers a device uint64 step counter (see ergm_set_rng_step_ptr).  The library keeps ONE raw pointer
        for the whole process and every dropout kernel dereferences it, so (a) a registered tensor is kept alive
        for the life of the process (8 bytes) - a freed counter would be a dangling device pointer - and (b) each
        training forward / backward of an engine re-registers its own counter (or none), so that engines and raw
        kernel calls never read another trainer's counter."""
        self.rng_step = t
        if t is not None:
            _RNG_KEEPALIVE.append(t)
        self._apply_rng_step()

    def _apply_rng_step(self):
        t = self.rng_step
        L.check(L.lib().ergm_set_rng_step_ptr(t.data_ptr() if t is not None else None), "ergm_set_rng_step_ptr")

    # GEMM helpers ------------------------------------------------------
    @staticmethod
    def _fwd_gemm(x, w_b, out, M, N, K, **kw):
        # dyn_m (keyword): device int32 run-time row count of a packed batch
        ops.gemm(x, w_b, out, M=M, N=N, K=K, a_major=K_MAJOR, b_major=MN_MAJOR, **kw)

    @staticmethod
    def _dgrad_gemm(dy, w_b, out, M, N_out, K_red, **kw):
        # out[M, N_out] = dy[M, K_red] @ W[N_out, K_red]^T   (W is Conv1D [in=N_out, out=K_red])
        ops.gemm(dy, w_b, out, M=M, N=N_out, K=K_red, a_major=K_MAJOR, b_major=K_MAJOR, **kw)

    def _wgrad_gemm(self, x, dy, dw, K_in, N_out, M_red, dyn=None):
        # dw[K_in, N_out] += x[M_red, K_in]^T @ dy[M_red, N_out]
        # 128x256 tiles, K (= tokens) split so that tiles * splits fills the 148 SMs once
        # (measured, profiles/r1_gemm_epilogue.md: 768x3072 44.1 -> 32.4 us, 768x768 15.1 -> 12.9 us
        # against 128x128 tiles with round(148 / tiles) splits)
        bn = 256 if N_out >= 256 else 128
        tiles = ((K_in + 127) // 128) * ((N_out + bn - 1) // bn)
        split = max(1, min(16, 148 // tiles))

        def launch():
            ops.gemm(x, dy, dw, M=K_in, N=N_out, K=M_red, a_major=MN_MAJOR, b_major=MN_MAJOR,
                     epilogue=L.EPI_ATOMIC, split_k=split, block_n=bn, dyn_k=dyn)

        self._side_launch(launch)

    def _side_launch(self, launch):
        side = self._wgrad_side
        if side is None:
            launch()
            return
        # Weight gradients are off the critical path (nothing in this backward reads them): they go to a side
        # stream, where their CTAs fill the SMs that the 1.3-2.6-wave GEMMs / attention kernels of the main
        # chain leave idle in their last wave.  dy is ready on the main stream now; the main stream waits for
        # the side stream (_side_join) before it overwrites any dy buffer.
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            launch()
            self._side_ev = torch.cuda.Event()
            self._side_ev.record(side)

    def _side_join(self):
        """Main stream waits for every weight-gradient GEMM issued so far."""
        if self._wgrad_side is not None and self._side_ev is not None:
            torch.cuda.current_stream().wait_event(self._side_ev)
            self._side_ev = None

    # ------------------------------------------------------------------
    def forward(self, input_ids, token_type_ids=None, labels=None, emotion_labels=None, imgs=None, auds=None,
                caption_ids=None, position_ids=None, past_len=0, kv_lens=None, training=False, save=False,
                want_logits=True, logits_fp32=False, dropout=None, heads=True, gen_state=None, legacy_past=None,
                pack=None):
        """Runs the full forward.  Returns a dict of device tensors (views into the workspace):
        logits [B,T,V] (bf16 or fp32, leading dim padded), emotion_logits [B,7], losses [5]
        (loss, lm_loss, emo_loss, 1/n_valid, 1/n_samples), hidden (bf16 ln_f output).
        pack (ops.Pack, already planned from the per-sample token counts): packed variable-length batch
        - every [B*T, .] matrix then holds the samples' real rows back to back (plus one row for position T-1 of a
        padded sample, which the emotion head reads), and the run-time row count bounds every kernel."""
        self.ensure_params()
        self.store.refresh_shadow()
        cfg, H, nh, Lyr, I, V = self.cfg, self.H, self.nh, self.L, self.I, self.V
        B, T = input_ids.shape
        M = B * T
        if legacy_past is not None:
            past_len = legacy_past[0].shape[1]
        if T + past_len > self.n_pos and position_ids is None:
            raise ValueError("sequence length %d + past %d exceeds n_positions %d" % (T, past_len, self.n_pos))
        ws = self.ws_train if save else self.ws_eval
        dev = self.device
        f32, bf16 = torch.float32, torch.bfloat16
        eps = cfg.layer_norm_epsilon
        pd_embd = cfg.embd_pdrop if (training and dropout is None) else (dropout or 0.0) if training else 0.0
        pd_attn = cfg.attn_pdrop if (training and dropout is None) else (dropout or 0.0) if training else 0.0
        pd_res = cfg.resid_pdrop if (training and dropout is None) else (dropout or 0.0) if training else 0.0
        site0 = self._new_sites(training)
        seed = self.seed
        if training:
            self._apply_rng_step()
        self.forward_serial = getattr(self, "forward_serial", 0) + 1

        def lname(base, l):
            return "%s_%d" % (base, l) if save else base

        dyn = None
        if pack is not None:
            if legacy_past is not None or gen_state is not None or past_len or position_ids is not None or not heads:
                raise L.ErgmError("packed batches are a training / scoring layout (no KV cache, default positions)")
            if pack.B != B or pack.T != T:
                raise ValueError("pack was planned for a [%d, %d] batch, got [%d, %d]" % (pack.B, pack.T, B, T))
            dyn = pack.n_rows
            kv_lens = None   # the pack carries the per-sample key counts
            ops.DYN_HINT = int(getattr(pack, "rows_hint", 0))   # typical row count: steers the GEMM tile shapes
        fuse = imgs is not None and past_len == 0  # boundary decision (3): fusion on the prefill only
        if imgs is not None and auds is None:
            raise ValueError("imgs given without auds (model.py:495-498 uses both)")
        img2 = aud2 = None
        proj_saved = {}
        if fuse:
            img2 = self._fused_feature(imgs, B, H, "imgs", "visual_proj", ws, proj_saved)
            aud2 = self._fused_feature(auds, B, H, "auds", "audio_proj", ws, proj_saved)
        x = ws.get(lname("x", 0), (M, H), f32)
        ops.embed_fuse_fwd(input_ids, token_type_ids, position_ids, self.p("transformer.wte.weight"),
                           self.p("transformer.wpe.weight"), img2, aud2, x, past_len=past_len,
                           dropout_p=pd_embd, seed=seed, offset=site0, pack=pack)
        enc = None
        Tc = Mc = 0
        if caption_ids is not None:
            caption_ids = caption_ids.reshape(B, -1)
            Tc = caption_ids.shape[1]
            Mc = B * Tc
            enc = ws.get("enc", (Mc, H), bf16)
            ops.gather_rows_bf16(caption_ids, self.p("transformer.wte.weight"), enc)

        sv = dict(B=B, T=T, Tc=Tc, layers=[], site0=site0, seed=seed, pd=(pd_embd, pd_attn, pd_res),
                  past_len=past_len, kv_lens=kv_lens, ids=input_ids, tts=token_type_ids, pos=position_ids,
                  cap=caption_ids, enc=enc, labels=labels, emo=emotion_labels, fuse=fuse,
                  proj=proj_saved, pack=pack) if save else None
        kv_present = []
        for l in range(Lyr):
            pfx = "transformer.h.%d." % l
            s_attn, s_res1, s_xattn, s_res2, s_res3 = [site0 + 8 * l + 1 + i for i in range(5)]
            rec = {}
            # ---- self attention (model.py:297-309) ----
            a1 = ws.get(lname("a1", l), (M, H), bf16)
            mean1 = ws.get(lname("mean1", l), (M,), f32)
            rstd1 = ws.get(lname("rstd1", l), (M,), f32)
            ops.ln_fwd(x, self.p(pfx + "ln_1.weight"), self.p(pfx + "ln_1.bias"), a1, None, mean1, rstd1, eps, rows_dyn=dyn)
            qkv = ws.get("qkv_%d" % l if (save or past_len == 0) else "qkv", (M, 3 * H), bf16)
            self._fwd_gemm(a1, self.pb(pfx + "attn.c_attn.weight"), qkv, M, 3 * H, H,
                           bias=self.p(pfx + "attn.c_attn.bias"), dyn_m=dyn)
            ctx = ws.get(lname("ctx", l), (M, H), bf16)
            ctx32 = ws.get(lname("ctx32", l), (M, H), f32) if save else None
            lse1 = ws.get(lname("lse1", l), (B, nh, T), f32)
            if legacy_past is not None:
                # legacy tuple cache (model.py:228-231): keys/values = cat(past, new); O(ctx) copy like the
                # reference's torch.cat — the fast decode path is the paged cache in generation.py
                kvf = torch.cat([legacy_past[l], qkv.view(B, T, 3 * H)[:, :, H:]], dim=1).contiguous()
                Tk_all = past_len + T
                kv2d = kvf.view(B * Tk_all, 2 * H)
                ops.attn_fwd(qkv, kv2d, kv2d, ctx, lse1, B=B, nh=nh, Tq=T, Tk=Tk_all, q_col0=0, k_col0=0, v_col0=H,
                             causal=True, causal_off=past_len, kv_lens=kv_lens)
                kv_present.append(kvf)
            else:
                ops.attn_fwd(qkv, qkv, qkv, ctx, lse1, B=B, nh=nh, Tq=T, Tk=T, q_col0=0, k_col0=H, v_col0=2 * H,
                             causal=True, kv_lens=kv_lens, dropout_p=pd_attn, seed=seed, offset=s_attn,
                             out_f32=ctx32, pack=pack, pack_kv=pack is not None)
                kv_present.append(qkv)
            if gen_state is not None:
                ops.kv_to_pages(qkv, gen_state.pool[l], gen_state.block_table, kv_lens, B=B, T=T, nh=nh,
                                k_col0=H, v_col0=2 * H)
            x1 = ws.get(lname("x1", l), (M, H), f32) if save else x
            self._fwd_gemm(ctx, self.pb(pfx + "attn.c_proj.weight"), x1, M, H, H,
                           bias=self.p(pfx + "attn.c_proj.bias"), residual=x, dropout_p=pd_res, seed=seed,
                           offset=s_res1, dyn_m=dyn)
            rec.update(x=x, a1=a1, mean1=mean1, rstd1=rstd1, qkv=qkv, ctx=ctx, ctx32=ctx32, lse1=lse1, x1=x1)
            # ---- cross attention over caption embeddings (model.py:311-329) ----
            x2 = x1
            if enc is not None:
                a2 = ws.get(lname("a2", l), (M, H), bf16)
                mean2 = ws.get(lname("mean2", l), (M,), f32)
                rstd2 = ws.get(lname("rstd2", l), (M,), f32)
                ops.ln_fwd(x1, self.p(pfx + "ln_cross_attn.weight"), self.p(pfx + "ln_cross_attn.bias"), a2, None,
                           mean2, rstd2, eps, rows_dyn=dyn)
                q2 = ws.get(lname("q2", l), (M, H), bf16)
                self._fwd_gemm(a2, self.pb(pfx + "crossattention.q_attn.weight"), q2, M, H, H,
                               bias=self.p(pfx + "crossattention.q_attn.bias"), dyn_m=dyn)
                kv2 = gen_state.kv2[l] if gen_state is not None else ws.get(lname("kv2", l), (Mc, 2 * H), bf16)
                self._fwd_gemm(enc, self.pb(pfx + "crossattention.c_attn.weight"), kv2, Mc, 2 * H, H,
                               bias=self.p(pfx + "crossattention.c_attn.bias"))
                ctx2 = ws.get(lname("ctx2", l), (M, H), bf16)
                ctx2_32 = ws.get(lname("ctx2_32", l), (M, H), f32) if save else None
                lse2 = ws.get(lname("lse2", l), (B, nh, T), f32)
                ops.attn_fwd(q2, kv2, kv2, ctx2, lse2, B=B, nh=nh, Tq=T, Tk=Tc, q_col0=0, k_col0=0, v_col0=H,
                             causal=False, dropout_p=pd_attn, seed=seed, offset=s_xattn, out_f32=ctx2_32, pack=pack)
                x2 = ws.get(lname("x2", l), (M, H), f32) if save else x1
                self._fwd_gemm(ctx2, self.pb(pfx + "crossattention.c_proj.weight"), x2, M, H, H,
                               bias=self.p(pfx + "crossattention.c_proj.bias"), residual=x1, dropout_p=pd_res,
                               seed=seed, offset=s_res2, dyn_m=dyn)
                rec.update(a2=a2, mean2=mean2, rstd2=rstd2, q2=q2, kv2=kv2, ctx2=ctx2, ctx2_32=ctx2_32, lse2=lse2, x2=x2)
            # ---- MLP (model.py:331-334, 262-267) ----
            a3 = ws.get(lname("a3", l), (M, H), bf16)
            mean3 = ws.get(lname("mean3", l), (M,), f32)
            rstd3 = ws.get(lname("rstd3", l), (M,), f32)
            ops.ln_fwd(x2, self.p(pfx + "ln_2.weight"), self.p(pfx + "ln_2.bias"), a3, None, mean3, rstd3, eps, rows_dyn=dyn)
            g = ws.get(lname("g", l), (M, I), bf16)
            u = ws.get(lname("u", l), (M, I), bf16) if save else None
            self._fwd_gemm(a3, self.pb(pfx + "mlp.c_fc.weight"), g, M, I, H, bias=self.p(pfx + "mlp.c_fc.bias"),
                           preact=u, epilogue=L.EPI_GELU, dyn_m=dyn)
            x3 = ws.get(lname("x", l + 1), (M, H), f32) if save else x2
            self._fwd_gemm(g, self.pb(pfx + "mlp.c_proj.weight"), x3, M, H, I, bias=self.p(pfx + "mlp.c_proj.bias"),
                           residual=x2, dropout_p=pd_res, seed=seed, offset=s_res3, dyn_m=dyn)
            rec.update(a3=a3, mean3=mean3, rstd3=rstd3, g=g, u=u)
            if save:
                sv["layers"].append(rec)
            x = x3
        if not heads:
            return dict(x_final=x, kv_present=kv_present, B=B, T=T)
        if pack is None:
            ops.DYN_HINT = 0
        # ---- final LN, heads, losses (model.py:578, 698-721) ----
        hn = ws.get("hn", (M, H), bf16)
        meanf = ws.get("meanf", (M,), f32)
        rstdf = ws.get("rstdf", (M,), f32)
        ops.ln_fwd(x, self.p("transformer.ln_f.weight"), self.p("transformer.ln_f.bias"), hn, None, meanf, rstdf, eps,
                   rows_dyn=dyn)
        out = dict(hidden=hn, kv_present=kv_present, B=B, T=T, pack=pack)
        ldl = (V + 63) // 64 * 64  # padded leading dimension: 16-byte rows for TMA / vector access
        logits = None
        # Label-sparse LM head: with labels, only the rows whose shifted label is not -100 enter the loss and the
        # gradients (model.py:705-708), so the head / CE / their backward run on those rows alone (run-time row
        # count on the device: no host sync, one captured graph serves every batch).  The logits of ALL positions
        # (outputs.logits, main.py:160) are produced by full_logits() only when somebody reads them.
        sparse = (labels is not None and not want_logits and not logits_fp32
                  and getattr(self.model, "ergm_sparse_lm_head", True))
        if pack is not None and labels is not None and not sparse:
            raise L.ErgmError("packed batches score the LM loss through the label-sparse head (ergm_sparse_lm_head)")
        need_lm = (want_logits or labels is not None) and not sparse
        if need_lm:
            logits = self.full_logits(hn, ws, logits_fp32)
        out["logits"] = logits
        out["logits_src"] = (hn, ws, logits_fp32)
        sums = ws.get("loss_sums", (4,), f32)
        sums.zero_()
        losses = ws.get("losses", (5,), f32)
        hlast = ws.get("hlast", (B, H), f32)
        emo_logits = ws.get("emo_logits", (B, 7), f32)
        emo_dlog = ws.get("emo_dlog", (B, 7), f32)
        ops.emotion_head_fwd(x, meanf, rstdf, self.p("transformer.ln_f.weight"), self.p("transformer.ln_f.bias"),
                             self.p("emotion_head.weight"), emotion_labels, hlast, emo_logits, emo_dlog, sums,
                             B=B, T=T, cu_rows=pack.cu if pack is not None else None)
        out["emotion_logits"] = emo_logits
        lse = row_loss = None
        sp = None
        if sparse:
            # plan -> gather -> GEMM (run-time M) -> CE: ergm_lmhead_ce_fwd, scratch + backward state in one workspace
            nbytes = ops.lmhead_ce_layout(M, H, V, save)[-1]
            sp = ops.lmhead_ce_views(ws.get("lm_head_ws", (nbytes,), torch.uint8), M, H, V, save)
            ops.lmhead_ce_fwd(hn, self.pb("transformer.wte.weight"), labels, sums, sp["ws"], T=T, V=V, pack=pack)
            lse = sp["lse"]
        elif labels is not None:
            lse = ws.get("ce_lse", (M,), f32)
            row_loss = ws.get("ce_row_loss", (M,), f32)
            ops.ce_fwd(logits, labels, lse, row_loss, sums, T=T, V=V, hn=hn, w=self.pb("transformer.wte.weight"))
        out["loss_sums"] = sums
        out["losses"] = losses
        out["has_lm"] = labels is not None
        out["has_emo"] = emotion_labels is not None
        if save:
            sv.update(xf=x, hn=hn, meanf=meanf, rstdf=rstdf, logits=logits, ce_lse=lse, hlast=hlast,
                      emo_dlog=emo_dlog, losses=losses, ldl=ldl, sparse=sp)
            # One set of saved activations exists (named, reused workspaces): tag it so that a backward through
            # an OLDER forward - loss = model(a).loss + model(b).loss - fails loudly instead of differentiating
            # through the wrong activations.
            self.forward_id = getattr(self, "forward_id", 0) + 1
            sv["id"] = self.forward_id
            out["forward_id"] = self.forward_id
            self.saved = sv
        return out

    def full_logits(self, hn, ws, logits_fp32=False):
        """LM head on every position: logits[M, ldl] = ln_f(x) @ wte^T (model.py:698)."""
        M, H, V = hn.shape[0], self.H, self.V
        ldl = (V + 63) // 64 * 64
        f32, bf16 = torch.float32, torch.bfloat16
        logits = ws.get("logits32" if logits_fp32 else "logits", (M, ldl), f32 if logits_fp32 else bf16)
        ops.gemm(hn, self.pb("transformer.wte.weight"), logits, M=M, N=V, K=H, a_major=K_MAJOR, b_major=K_MAJOR)
        return logits

    # ------------------------------------------------------------------
    # fp32 mode: forward-only verification path (north_star: logits within 1e-4, greedy ids exact)
    # ------------------------------------------------------------------
    def _w6(self, name, linear=False):
        """6x split-expanded bf16 copy of an fp32 weight (cached per parameter version)."""
        cache = self.__dict__.setdefault("_w6_cache", {})
        w = self.p(name)
        self.store.refresh_shadow()
        ver = self.store.weights_epoch
        hit = cache.get(name)
        if hit is not None and hit[0] == ver:
            return hit[1]
        if linear:   # nn.Linear weight [N, K]: reduction dim = cols -> [N, 6K], K-major operand
            out = torch.empty(w.shape[0], 6 * w.shape[1], dtype=torch.bfloat16, device=self.device)
            ops.split3_expand(w, out, side=1, kdim=1)
        else:        # Conv1D weight [K, N]: reduction dim = rows -> [6K, N], MN-major operand
            out = torch.empty(6 * w.shape[0], w.shape[1], dtype=torch.bfloat16, device=self.device)
            ops.split3_expand(w, out, side=1, kdim=0)
        cache[name] = (ver, out)
        return out

    def _gemm32(self, x32, wname, out, bias=None, residual=None, gelu=False, linear=False, N=None):
        """out[M,N] (fp32) = x32[M,K] @ W (+bias) (+gelu) (+residual) at fp32 accuracy on the bf16 tensor cores."""
        M, K = x32.shape
        a6 = self.ws_eval.get("a6", (M, 6 * K), torch.bfloat16)
        ops.split3_expand(x32, a6, side=0, kdim=1)
        w6 = self._w6(wname, linear)
        if N is None:
            N = w6.shape[0] if linear else w6.shape[1]
        epi = (L.EPI_GELU | L.EPI_EXACT) if gelu else 0
        ops.gemm(a6, w6, out, M=M, N=N, K=6 * K, a_major=K_MAJOR, b_major=K_MAJOR if linear else MN_MAJOR,
                 bias=bias, residual=residual, epilogue=epi)

    def forward_fp32(self, input_ids, token_type_ids=None, labels=None, emotion_labels=None, imgs=None, auds=None,
                     caption_ids=None, position_ids=None, kv_lens=None):
        """Same dataflow as forward() with every product at fp32 accuracy; eval / no-grad only."""
        self.ensure_params()
        cfg, H, nh, Lyr, I, V = self.cfg, self.H, self.nh, self.L, self.I, self.V
        B, T = input_ids.shape
        M = B * T
        ws = self.ws_eval
        f32 = torch.float32
        eps = cfg.layer_norm_epsilon
        img2 = aud2 = None
        if imgs is not None:
            img2 = self._fused_feature(imgs, B, H, "imgs", "visual_proj", ws, None, fp32=True)
            aud2 = self._fused_feature(auds, B, H, "auds", "audio_proj", ws, None, fp32=True)
        x = ws.get("f32_x", (M, H), f32)
        ops.embed_fuse_fwd(input_ids, token_type_ids, position_ids, self.p("transformer.wte.weight"),
                           self.p("transformer.wpe.weight"), img2, aud2, x)
        enc = None
        if caption_ids is not None:
            caption_ids = caption_ids.reshape(B, -1)
            Tc = caption_ids.shape[1]
            enc = self.p("transformer.wte.weight").index_select(0, caption_ids.reshape(-1))  # gather only
        a = ws.get("f32_a", (M, H), f32)
        qkv = ws.get("f32_qkv", (M, 3 * H), f32)
        ctx = ws.get("f32_ctx", (M, H), f32)
        g = ws.get("f32_g", (M, I), f32)
        for l in range(Lyr):
            pfx = "transformer.h.%d." % l
            ops.ln_fwd(x, self.p(pfx + "ln_1.weight"), self.p(pfx + "ln_1.bias"), None, a, None, None, eps)
            self._gemm32(a, pfx + "attn.c_attn.weight", qkv, bias=self.p(pfx + "attn.c_attn.bias"))
            ops.attn_fwd_f32(qkv, qkv, qkv, ctx, B=B, nh=nh, Tq=T, Tk=T, q_col0=0, k_col0=H, v_col0=2 * H,
                             causal=True, kv_lens=kv_lens)
            self._gemm32(ctx, pfx + "attn.c_proj.weight", x, bias=self.p(pfx + "attn.c_proj.bias"), residual=x)
            if enc is not None:
                q2 = ws.get("f32_q2", (M, H), f32)
                kv2 = ws.get("f32_kv2", (B * Tc, 2 * H), f32)
                ops.ln_fwd(x, self.p(pfx + "ln_cross_attn.weight"), self.p(pfx + "ln_cross_attn.bias"), None, a, None,
                           None, eps)
                self._gemm32(a, pfx + "crossattention.q_attn.weight", q2, bias=self.p(pfx + "crossattention.q_attn.bias"))
                self._gemm32(enc, pfx + "crossattention.c_attn.weight", kv2,
                             bias=self.p(pfx + "crossattention.c_attn.bias"))
                ops.attn_fwd_f32(q2, kv2, kv2, ctx, B=B, nh=nh, Tq=T, Tk=Tc, q_col0=0, k_col0=0, v_col0=H, causal=False)
                self._gemm32(ctx, pfx + "crossattention.c_proj.weight", x,
                             bias=self.p(pfx + "crossattention.c_proj.bias"), residual=x)
            ops.ln_fwd(x, self.p(pfx + "ln_2.weight"), self.p(pfx + "ln_2.bias"), None, a, None, None, eps)
            self._gemm32(a, pfx + "mlp.c_fc.weight", g, bias=self.p(pfx + "mlp.c_fc.bias"), gelu=True)
            self._gemm32(g, pfx + "mlp.c_proj.weight", x, bias=self.p(pfx + "mlp.c_proj.bias"), residual=x)
        meanf = ws.get("meanf", (M,), f32)
        rstdf = ws.get("rstdf", (M,), f32)
        ops.ln_fwd(x, self.p("transformer.ln_f.weight"), self.p("transformer.ln_f.bias"), None, a, meanf, rstdf, eps)
        ldl = (V + 63) // 64 * 64
        logits = ws.get("logits32", (M, ldl), f32)
        self._gemm32(a, "transformer.wte.weight", logits, linear=True, N=V)
        sums = ws.get("loss_sums", (4,), f32)
        sums.zero_()
        losses = ws.get("losses", (5,), f32)
        hlast = ws.get("hlast", (B, H), f32)
        emo_logits = ws.get("emo_logits", (B, 7), f32)
        emo_dlog = ws.get("emo_dlog", (B, 7), f32)
        ops.emotion_head_fwd(x, meanf, rstdf, self.p("transformer.ln_f.weight"), self.p("transformer.ln_f.bias"),
                             self.p("emotion_head.weight"), emotion_labels, hlast, emo_logits, emo_dlog, sums, B=B, T=T)
        if labels is not None:
            lse = ws.get("ce_lse", (M,), f32)
            row_loss = ws.get("ce_row_loss", (M,), f32)
            ops.ce_fwd(logits, labels, lse, row_loss, sums, T=T, V=V)
        return dict(logits=logits, emotion_logits=emo_logits, loss_sums=sums, losses=losses, hidden=a,
                    has_lm=labels is not None, has_emo=emotion_labels is not None, kv_present=None, B=B, T=T)

    def finalize_loss(self, out):
        """loss = CE_lm + CE_emotion (model.py:704-721) from the (possibly all-reduced) sums."""
        ops.loss_finalize(out["loss_sums"], out["has_lm"], out["has_emo"], out["losses"])
        return out["losses"]

    def _fused_feature(self, feat, B, H, what, proj, ws, saved, fp32=False):
        """The [B, H] fp32 vector added at position 0 (imgs) / 1 (auds) (model.py:497-498).
        Reference layout (no `<proj>.weight` parameter in the model): the pooled H-wide feature is
        used as is.  A3 extension (model built with config.ergm_visual_dim / ergm_audio_dim): `feat`
        is the raw feature SEQUENCE [B, T, D] (audio [B,113,768], visual [B,197*Kf,768],
        text_feature.py:44,49): time-mean (feature_extraction.py:63,69) -> Linear(D -> H)."""
        wname = proj + ".weight"
        if wname not in self.P:
            return self._feature_rows(feat, B, H, what)
        if isinstance(feat, (list, tuple)):
            feat = torch.stack([torch.as_tensor(f) for f in feat])
        feat = feat.to(self.device)
        if feat.dim() == 2:
            feat = feat[:, None, :]
        D = self.p(wname).shape[1]
        if feat.dim() != 3 or feat.shape[0] != B or feat.shape[2] != D:
            raise ValueError("%s must be a feature sequence [B=%d, T, D=%d] for the %s projection (got %s)"
                             % (what, B, D, proj, tuple(feat.shape)))
        if feat.dtype != torch.float32 or feat.stride(2) != 1 or feat.stride(0) % 4 or feat.stride(1) % 4 \
                or feat.data_ptr() % 16:
            feat = feat.float().contiguous()
        f32, bf16 = torch.float32, torch.bfloat16
        out = ws.get(what + "_proj_out", (B, H), f32)
        if fp32:
            pooled = ws.get(what + "_pooled32", (B, D), f32)
            ops.mm_pool_fwd(feat, pooled, None)
            self._gemm32(pooled, wname, out, bias=self.p(proj + ".bias"), linear=True)
            return out
        pooled = ws.get(what + "_pooled", (B, D), bf16)
        ops.mm_pool_fwd(feat, None, pooled)
        ops.gemm(pooled, self.pb(wname), out, M=B, N=H, K=D, a_major=K_MAJOR, b_major=K_MAJOR,
                 bias=self.p(proj + ".bias"))
        if saved is not None:
            saved[what] = (proj, pooled, D)
        return out

    def _feature_rows(self, feat, B, H, what):
        """imgs: [B, >=1, H] (first row used, model.py:497 `imgs[i][0]`) or [B, H]; auds: [B, H] or
        [B, 1, H] (model.py:498 `auds[i].unsqueeze(0)` broadcasts)."""
        if isinstance(feat, (list, tuple)):
            feat = torch.stack([torch.as_tensor(f[0] if what == "imgs" and torch.as_tensor(f).dim() > 1 else f)
                                for f in feat]).to(self.device)
        if feat.dim() == 3:
            feat = feat[:, 0]
        if feat.dim() != 2 or feat.shape[0] != B or feat.shape[1] != H:
            raise ValueError("%s must be broadcastable to [B=%d, n_embd=%d] (got %s); use the pooled / projected "
                             "feature path for other widths" % (what, B, H, tuple(feat.shape)))
        if feat.dtype != torch.float32 or feat.stride(1) != 1 or feat.stride(0) % 4 or feat.data_ptr() % 16:
            feat = feat.float().contiguous()
        return feat

    # ------------------------------------------------------------------
    def backward(self, grad_loss, accumulate=False, on_layer_done=None, forward_id=None):
        """Hand-written backward of the whole path.  grad_loss: device fp32 scalar tensor (dLoss).
        Gradients are accumulated into the flat gradient buffer (zeroed first unless
        `accumulate`).  forward_id: the tag of the forward this backward belongs to (checked)."""
        sv = self.saved
        if sv is None:
            raise RuntimeError("backward() without a saved training forward (ergm_b200 keeps the activations of "
                               "ONE training forward per model: each loss must be back-propagated before the "
                               "next training forward, and only once)")
        if forward_id is not None and sv["id"] != forward_id:
            raise RuntimeError("backward() through training forward #%d, but the saved activations belong to the "
                               "later forward #%d: ergm_b200 keeps one set of saved activations per model - call "
                               "loss.backward() before the next training forward (accumulate gradients across "
                               "backward calls instead of summing losses)" % (forward_id, sv["id"]))
        self.saved = None
        self._apply_rng_step()
        cfg, H, nh, Lyr, I, V = self.cfg, self.H, self.nh, self.L, self.I, self.V
        B, T, Tc = sv["B"], sv["T"], sv["Tc"]
        M, Mc = B * T, B * Tc
        pack = sv.get("pack")
        dyn = pack.n_rows if pack is not None else None
        ops.DYN_HINT = int(getattr(pack, "rows_hint", 0)) if pack is not None else 0
        ws = self.ws_train
        f32, bf16 = torch.float32, torch.bfloat16
        seed, site0 = sv["seed"], sv["site0"]
        pd_embd, pd_attn, pd_res = sv["pd"]
        if not accumulate:
            self.store.grad.zero_()
        # opt-in: A/B over six alternating runs (30 steps each): 14.06 ms with the side stream, 13.99 ms without -
        # the GPU's block scheduler already back-fills the tails well enough, the fork/join events cost as much
        if os.environ.get("ERGM_WGRAD_STREAM", "0") == "1":
            if self.__dict__.get("_side_stream_obj") is None:
                self._side_stream_obj = torch.cuda.Stream(device=self.device)
            self._wgrad_side = self._side_stream_obj
        else:
            self._wgrad_side = None
        self._side_ev = None
        losses = sv["losses"]
        scales = ws.get("bwd_scales", (2,), f32)
        ops.scalar_mul(grad_loss, losses[3:4], scales[0:1])  # dL/d(row loss) = g / n_valid
        ops.scalar_mul(grad_loss, losses[4:5], scales[1:2])  # g / n_samples
        dhn = ws.get("dhn", (M, H), f32)
        hn = sv["hn"]
        sp = sv.get("sparse")
        if sp is not None:
            # label-sparse head backward on the compacted rows (run-time count sp["count"])
            # dlogits, d hn (overwritten), d wte (accumulated) from the forward's workspace: ergm_lmhead_ce_bwd
            ops.lmhead_ce_bwd(self.pb("transformer.wte.weight"), scales[0:1], dhn, self.pg("transformer.wte.weight"),
                              sp["ws"], V=V)
        elif sv["labels"] is not None:
            ldl = sv["ldl"]
            dlogits = ws.get("dlogits", (M, ldl), bf16)
            ops.ce_bwd(sv["logits"], sv["labels"], sv["ce_lse"], scales[0:1], dlogits, T=T, V=V)
            wte_b = self.pb("transformer.wte.weight")
            # d hn = dlogits @ wte
            ops.gemm(dlogits, wte_b, dhn, M=M, N=H, K=V, a_major=K_MAJOR, b_major=MN_MAJOR)
            # d wte += dlogits^T @ hn
            ops.gemm(dlogits, hn, self.pg("transformer.wte.weight"), M=V, N=H, K=M, a_major=MN_MAJOR,
                     b_major=MN_MAJOR, epilogue=L.EPI_ATOMIC, block_n=2256)  # pair 256x256: 0.785 -> 0.435 ms
        else:
            dhn.zero_()
        if sv["emo"] is not None:
            ops.emotion_head_bwd(sv["emo_dlog"], sv["hlast"], self.p("emotion_head.weight"), scales[1:2],
                                 self.pg("emotion_head.weight"), dhn, B=B, T=T, cu_rows=pack.cu if pack is not None else None)
        # residual-stream gradient dx (fp32) and its bf16 (dropout-masked) operand copy dxb
        dx = ws.get("dx", (M, H), f32)
        dxb = ws.get("dxb", (M, H), bf16)
        last = "transformer.h.%d." % (Lyr - 1)
        dI = ws.get("du", (M, I), bf16)
        dH = ws.get("dact", (M, H), bf16)
        dqkv = ws.get("dqkv", (M, 3 * H), bf16)
        if pack is not None:
            # gradient-side operands of the run-time-K weight-gradient GEMMs: rows [n, roundup(n, 128)) must be zero
            # (no producer of this step writes them; a previous, longer batch may have)
            for buf in (dxb, dI, dqkv):
                ops.zero_rows_dyn(buf, dyn)
        ops.ln_bwd(dhn, sv["xf"], sv["meanf"], sv["rstdf"], self.p("transformer.ln_f.weight"), None, dx, dxb,
                   self.pg("transformer.ln_f.weight"), self.pg("transformer.ln_f.bias"),
                   self.pg(last + "mlp.c_proj.bias"), dropout_p=pd_res, seed=seed,
                   offset=site0 + 8 * (Lyr - 1) + 5, rows_dyn=dyn)
        # fp32 scratch of the attention backward (0 bytes up to T = 256: dQ stays in TMEM there)
        ab_bytes = ops.attn_bwd_workspace_bytes(B, nh, T)
        ab_ws = ws.get("attn_bwd_ws", (ab_bytes,), torch.uint8) if ab_bytes else None
        delta = ws.get("delta", (B, nh, T), f32)
        denc = None
        if sv["enc"] is not None:
            denc = ws.get("denc", (Mc, H), f32)
            denc.zero_()
            dq2 = ws.get("dq2", (M, H), bf16)
            dkv2 = ws.get("dkv2", (Mc, 2 * H), bf16)
            if pack is not None:
                ops.zero_rows_dyn(dq2, dyn)
        for l in reversed(range(Lyr)):
            pfx = "transformer.h.%d." % l
            r = sv["layers"][l]
            s_attn, s_res1, s_xattn, s_res2, s_res3 = [site0 + 8 * l + 1 + i for i in range(5)]
            has_x = "a2" in r
            # ---- MLP backward ----
            self._wgrad_gemm(r["g"], dxb, self.pg(pfx + "mlp.c_proj.weight"), I, H, M, dyn)
            if M % 256 == 0 and I % 256 == 0:
                # GELU' and the c_fc bias gradient ride in the dgrad epilogue (lean FM_GELU_GRAD mode; the boundary
                # slab of a packed batch's run-time row count goes through the generic epilogue, same sums)
                self._dgrad_gemm(dxb, self.pb(pfx + "mlp.c_proj.weight"), dI, M, I, H, gelu_grad_of=r["u"],
                                 colsum=self.pg(pfx + "mlp.c_fc.bias"), block_n=2256, dyn_m=dyn)
            else:
                self._dgrad_gemm(dxb, self.pb(pfx + "mlp.c_proj.weight"), dI, M, I, H, dyn_m=dyn)
                ops.gelu_bwd_colsum(dI, r["u"], self.pg(pfx + "mlp.c_fc.bias"), rows_dyn=dyn)
            self._wgrad_gemm(r["a3"], dI, self.pg(pfx + "mlp.c_fc.weight"), H, I, M, dyn)
            self._dgrad_gemm(dI, self.pb(pfx + "mlp.c_fc.weight"), dH, M, H, I, dyn_m=dyn)
            x_in = r["x2"] if has_x else r["x1"]
            nb = pfx + ("crossattention.c_proj.bias" if has_x else "attn.c_proj.bias")
            self._side_join()  # dxb is about to be overwritten
            ops.ln_bwd(dH, x_in, r["mean3"], r["rstd3"], self.p(pfx + "ln_2.weight"), dx, dx, dxb,
                       self.pg(pfx + "ln_2.weight"), self.pg(pfx + "ln_2.bias"), self.pg(nb), dropout_p=pd_res,
                       seed=seed, offset=s_res2 if has_x else s_res1, rows_dyn=dyn)
            # ---- cross attention backward ----
            if has_x:
                self._wgrad_gemm(r["ctx2"], dxb, self.pg(pfx + "crossattention.c_proj.weight"), H, H, M, dyn)
                self._dgrad_gemm(dxb, self.pb(pfx + "crossattention.c_proj.weight"), dH, M, H, H, dyn_m=dyn)
                self._side_join()  # dkv2 is about to be overwritten
                gbx = self.pg(pfx + "crossattention.c_attn.bias")  # Q / K / V bias gradients come out of attn_bwd
                ops.attn_bwd(r["q2"], r["kv2"], r["kv2"], r["ctx2"], dH, r["lse2"], delta, dq2, dkv2, dkv2,
                             B=B, nh=nh, Tq=T, Tk=Tc, q_col0=0, k_col0=0, v_col0=H, dk_col0=0, dv_col0=H,
                             causal=False, dropout_p=pd_attn, seed=seed, offset=s_xattn, out_f32=r["ctx2_32"],
                             dq_colsum=self.pg(pfx + "crossattention.q_attn.bias"), dk_colsum=gbx[:H],
                             dv_colsum=gbx[H:], pack=pack, workspace=ab_ws)
                self._wgrad_gemm(r["a2"], dq2, self.pg(pfx + "crossattention.q_attn.weight"), H, H, M, dyn)
                self._wgrad_gemm(sv["enc"], dkv2, self.pg(pfx + "crossattention.c_attn.weight"), H, 2 * H, Mc)
                self._dgrad_gemm(dq2, self.pb(pfx + "crossattention.q_attn.weight"), dH, M, H, H, dyn_m=dyn)
                # d enc accumulates over layers (the caption embeddings feed every block, model.py:521)
                # (off the critical path too: denc is only consumed by the embedding backward at the very end)
                wkv = self.pb(pfx + "crossattention.c_attn.weight")
                self._side_launch(lambda: self._dgrad_gemm(dkv2, wkv, denc, Mc, H, 2 * H, residual=denc))
                self._side_join()
                ops.ln_bwd(dH, r["x1"], r["mean2"], r["rstd2"], self.p(pfx + "ln_cross_attn.weight"), dx, dx, dxb,
                           self.pg(pfx + "ln_cross_attn.weight"), self.pg(pfx + "ln_cross_attn.bias"),
                           self.pg(pfx + "attn.c_proj.bias"), dropout_p=pd_res, seed=seed, offset=s_res1, rows_dyn=dyn)
            # ---- self attention backward ----
            self._wgrad_gemm(r["ctx"], dxb, self.pg(pfx + "attn.c_proj.weight"), H, H, M, dyn)
            self._dgrad_gemm(dxb, self.pb(pfx + "attn.c_proj.weight"), dH, M, H, H, dyn_m=dyn)
            self._side_join()  # dqkv is about to be overwritten
            qkv = r["qkv"]
            gb = self.pg(pfx + "attn.c_attn.bias")
            ops.attn_bwd(qkv, qkv, qkv, r["ctx"], dH, r["lse1"], delta, dqkv, dqkv, dqkv, B=B, nh=nh, Tq=T, Tk=T,
                         q_col0=0, k_col0=H, v_col0=2 * H, dq_col0=0, dk_col0=H, dv_col0=2 * H, causal=True,
                         kv_lens=sv["kv_lens"], dropout_p=pd_attn, seed=seed, offset=s_attn, out_f32=r["ctx32"],
                         dq_colsum=gb[:H], dk_colsum=gb[H:2 * H], dv_colsum=gb[2 * H:], pack=pack,
                         pack_kv=pack is not None, workspace=ab_ws)
            self._wgrad_gemm(r["a1"], dqkv, self.pg(pfx + "attn.c_attn.weight"), H, 3 * H, M, dyn)
            self._dgrad_gemm(dqkv, self.pb(pfx + "attn.c_attn.weight"), dH, M, H, 3 * H, dyn_m=dyn)
            self._side_join()
            if l > 0:
                prev = "transformer.h.%d." % (l - 1)
                ops.ln_bwd(dH, r["x"], r["mean1"], r["rstd1"], self.p(pfx + "ln_1.weight"), dx, dx, dxb,
                           self.pg(pfx + "ln_1.weight"), self.pg(pfx + "ln_1.bias"),
                           self.pg(prev + "mlp.c_proj.bias"), dropout_p=pd_res, seed=seed,
                           offset=site0 + 8 * (l - 1) + 5, rows_dyn=dyn)
            else:
                ops.ln_bwd(dH, r["x"], r["mean1"], r["rstd1"], self.p(pfx + "ln_1.weight"), dx, dx, None,
                           self.pg(pfx + "ln_1.weight"), self.pg(pfx + "ln_1.bias"), None, rows_dyn=dyn)
            if on_layer_done is not None:
                # every gradient of layer l is final now: its mlp.c_proj.bias was completed earlier by
                # layer l+1's ln_1 backward, and this layer's ln_1 backward only touched layer l-1's slot
                on_layer_done(l)
        # ---- embedding backward (model.py:459, 500-506) ----
        proj = sv.get("proj") or {}
        dfeat = {}
        for what in proj:
            dfeat[what] = ws.get("d" + what, (B, H), f32)
            dfeat[what].zero_()
        ops.embed_bwd(dx, sv["ids"], sv["tts"], sv["pos"], self.pg("transformer.wte.weight"),
                      self.pg("transformer.wpe.weight"), T=T, past_len=sv["past_len"], dimgs=dfeat.get("imgs"),
                      dauds=dfeat.get("auds"), dropout_p=pd_embd, seed=seed, offset=site0, pack=pack)
        for what, (pname, pooled, D) in proj.items():
            # Linear(D -> H) backward: dW[H, D] += dfeat^T @ pooled, db += colsum(dfeat)
            dfb = ws.get("d" + what + "_bf16", (B, H), bf16)
            ops.cast_f32_bf16_2d(dfeat[what], dfb, self.pg(pname + ".bias"))
            ops.gemm(dfb, pooled, self.pg(pname + ".weight"), M=H, N=D, K=B, a_major=MN_MAJOR, b_major=MN_MAJOR,
                     epilogue=L.EPI_ATOMIC, block_n=128)
        if denc is not None:
            ops.embed_bwd(denc, sv["cap"], None, None, self.pg("transformer.wte.weight"), None, T=Tc)
        self._side_join()
        self._wgrad_side = None
        if on_layer_done is not None:
            on_layer_done(-1)
        skip = () if sv["enc"] is not None else ("crossattention.", "ln_cross_attn.")
        self.store.bind_grads(skip)

/*
 * ergm_b200 C ABI — the drop-in boundary of the B200-native ERGM hot path.
 *
 * The reference (LovesickPatience/ERGM) has no FFI layer: its model file
 * src/model.py reaches the GPU only through PyTorch library ops.  Each entry
 * point below replaces the library-op call sites of one stage of
 * src/model.py (file:line cited per function).  Conventions:
 *   - plain C: raw device pointers, explicit sizes / leading dimensions,
 *     a cudaStream_t (passed as void*), no C++ or torch types;
 *   - every function is asynchronous on `stream`, never allocates, never
 *     synchronises, and returns 0 on success, a negative ERGM_ERR_* code for
 *     an argument error or a positive cudaError_t;
 *   - bf16 tensors are `uint16_t`-sized, row-major, leading dimensions are in
 *     elements.
 */
#ifndef ERGM_B200_H_
#define ERGM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ERGM_OK 0
#define ERGM_ERR_ARG (-1)
#define ERGM_ERR_UNSUPPORTED (-2)
#define ERGM_ERR_DRIVER (-3)

/* library / device probes (no compute) */
int ergm_abi_version(void);
int ergm_device_sm_count(void);
/* TMA descriptors are encoded once per distinct (pointer, shape, stride, box) */
/* and then served from a mutex-guarded table: hit / miss counters of it.      */
int ergm_tmap_cache_stats(uint64_t* hits, uint64_t* misses);
/* Dropout masks are Philox functions of (seed + *step, offset, element); the
 * step counter lives in device memory so that CUDA-graph replays draw fresh
 * masks.  NULL (default) disables the indirection.                          */
int ergm_set_rng_step_ptr(const uint64_t* dev_ptr);
int ergm_rng_step_advance(uint64_t* dev_ptr, uint64_t inc, void* stream);

/* ------------------------------------------------------------------------ */
/* GEMM: D[M,N] = epilogue(A[M,K] * B[K,N])  — bf16 operands, fp32 accumulate
 * in TMEM (tcgen05.mma, TMA-fed).  Replaces transformers Conv1D.forward
 * (addmm) used at model.py:218,219,222,244,263,265, nn.Linear lm_head at
 * model.py:698 and every dgrad / wgrad GEMM autograd derives from them.     */
/* ------------------------------------------------------------------------ */
enum {
  ERGM_MAJOR_K = 0,  /* operand is stored [rows(M or N), K], K contiguous      */
  ERGM_MAJOR_MN = 1  /* operand is stored [K, rows(M or N)], M/N contiguous    */
};
enum {
  ERGM_EPI_BIAS = 1,       /* + bias[N] (fp32)                                  */
  ERGM_EPI_GELU = 2,       /* gelu_new(x) (model.py:259,264)                    */
  ERGM_EPI_RESIDUAL = 4,   /* + residual[M,N] (fp32) (model.py:309,328,334)     */
  ERGM_EPI_ATOMIC = 8,     /* D += result via red.add.f32 (wgrad accumulate)    */
  ERGM_EPI_DROPOUT = 16,   /* inverted dropout on (acc+bias) before residual    */
  ERGM_EPI_PREACT = 32,    /* also store (acc+bias) before GELU to `preact`     */
  ERGM_EPI_EXACT = 64,     /* exact tanhf instead of tanh.approx in GELU        */
  ERGM_EPI_GELU_GRAD = 128 /* result *= gelu_new'(preact[M,ldd]) (bf16 D only; MLP
                              backward of model.py:264; `preact` is an INPUT)   */
};
enum { ERGM_DT_BF16 = 0, ERGM_DT_F32 = 1 };

typedef struct ergm_gemm_args {
  const void* a;       /* bf16 */
  const void* b;       /* bf16 */
  void* d;             /* bf16 or fp32, row-major [M, ldd] */
  const float* bias;   /* [N] or NULL */
  const float* residual; /* fp32 [M, ldr] or NULL */
  void* preact;        /* bf16 [M, ldd] or NULL */
  float* colsum;       /* NULL, or fp32 [N] += column sums of D as stored: bias gradient fused into the
                          ERGM_EPI_GELU_GRAD dgrad (bf16 D, M and N multiples of the tile) */
  int64_t lda, ldb, ldd, ldr;
  int32_t M, N, K;
  int32_t a_major, b_major;
  int32_t d_dtype;
  int32_t epilogue;    /* ERGM_EPI_* bitmask */
  int32_t split_k;     /* >=1; >1 requires ERGM_EPI_ATOMIC and fp32 D */
  int32_t block_n;     /* 0 = auto; 64 / 128 / 256 = single-CTA tile width;
                          2128 / 2256 = CTA-pair (cta_group::2) kernel, 256 x {128,256} tiles */
  float dropout_p;
  uint64_t seed, offset; /* Philox key / subsequence of this dropout site */
  const int32_t* dyn_count; /* NULL, or a DEVICE int32: run-time extent (<= the static M / K) read by the kernel, so
                               that one captured launch serves every batch (label-sparse LM head, packed batches) */
  int32_t dyn_dim;     /* 1: dyn_count bounds M (rows beyond it are not stored; with ERGM_EPI_ATOMIC the kernel picks
                          the K split itself); 2: dyn_count bounds K (ERGM_EPI_ATOMIC only; operand rows in
                          [count, roundup(count, 128)) must be zero) */
  int32_t dyn_hint;    /* expected value of *dyn_count (0 = unknown).  Only steers the host-side tile-shape choice
                          (wave quantisation is decided by the rows that really exist, not by the capacity); any
                          value is correct */
} ergm_gemm_args;

int ergm_gemm_bf16(const ergm_gemm_args* args, void* stream);


/* ------------------------------------------------------------------------ */
/* Packed variable-length batches.  The reference's collate (custom_dataset.py:102-132) right-pads
 * every sample to the batch maximum and model.py computes every pad position.  With an ergm_pack the path works
 * on the concatenation of the samples' real rows: sample b contributes its lens[b] real positions plus, when it is
 * padded, ONE row for position T-1 (the emotion head reads the last position, model.py:700; under the right-padded
 * attention mask that position attends to the sample's real tokens only).  All members are DEVICE int32 arrays
 * written by ergm_pack_plan; the row count is a run-time value on the device (one captured graph serves every
 * batch): row kernels take it as `rows_dyn`, GEMMs as ergm_gemm_args.dyn_count.  Entry points that take
 * `const ergm_pack* pack` accept NULL (padded [B, T] layout, the reference's).                               */
typedef struct ergm_pack {
  const int32_t* cu_rows;  /* [B + 1] first packed row of every sample */
  const int32_t* row_b;    /* [B * T] sample of a packed row */
  const int32_t* row_t;    /* [B * T] position of a packed row (column of the padded layout) */
  const int32_t* n_rows;   /* [1] packed rows in this batch */
  const int32_t* kv_lens;  /* [B] attendable keys per sample = its real tokens */
} ergm_pack;
int ergm_pack_plan(const int* lens, int B, int T, int* cu_rows, int* row_b, int* row_t, int* n_rows, int* kv_lens,
                   int cap /* >= B * T: capacity of row_b / row_t */, void* stream);
/* rows [*count, roundup(*count, 128)) of buf[cap, row_bytes] := 0: operands of run-time-K (wgrad) GEMMs */
int ergm_zero_rows_dyn(void* buf, int64_t row_bytes, const int* count, int cap, void* stream);

/* ------------------------------------------------------------------------ */
/* Embedding + multimodal fusion (model.py:458-507):                          */
/*   h[b,t] = ((wte[id] (+imgs[b] if t==0) (+auds[b] if t==1)) + wpe[pos]) + wte[type]
 * then embd dropout.  pos = past_len + t unless position_ids is given, in
 * which case pos = position_ids[b * pos_stride_b + t] (stride 0 = shared).  */
/* imgs / auds: fp32 [B, ld] (the pooled vectors model.py:497-498 adds) or    */
/* NULL.  err_flag (device int) is set to 1 on an out-of-range index.         */
int ergm_embed_fuse_fwd(const int64_t* ids, const int64_t* token_type_ids,
                        const int64_t* position_ids, int64_t pos_stride_b,
                        const int* past_lens /* nullable int32 [B]: per-sequence past length */,
                        const float* wte, const float* wpe,
                        const float* imgs, int64_t ld_img, const float* auds, int64_t ld_aud,
                        float* out, int B, int T, int H, int past_len, int vocab, int n_pos,
                        float dropout_p, uint64_t seed, uint64_t offset, int* err_flag,
                        const ergm_pack* pack /* nullable: out row r <- (row_b[r], row_t[r]) */, void* stream);
/* A3 extension (north_star "projects per-utterance audio and keyframe-visual   */
/* feature sequences into the hidden space"; the reference pools offline,        */
/* feature_extraction.py:63,69, and has no projection - SURVEY Appendix A D7):   */
/* time-mean of fp32 feature sequences seq[b, t, 0:D] (strides ld_b, ld_t in     */
/* elements; lens = nullable int32 [B] valid frames per sample) -> pooled        */
/* [B, D] as fp32 and/or bf16 (the A operand of the D -> H projection            */
/* ergm_gemm_bf16 whose fp32 output feeds ergm_embed_fuse_fwd's imgs / auds).    */
int ergm_mm_pool_fwd(const float* seq, int64_t ld_b, int64_t ld_t, const int* lens, int B,
                     int T, int D, float* pooled_f32, int64_t ld_f32, void* pooled_bf16,
                     int64_t ld_bf16, void* stream);
/* caption embeddings enc = wte[caption_ids] (model.py:460-463), bf16 output  */
int ergm_gather_rows_bf16(const int64_t* ids, const float* table, void* out_bf16, int rows,
                          int H, int vocab, int* err_flag, void* stream);
/* backward of the embedding stage: scatter-add dh rows into dwte (by id and  */
/* by token type), dwpe, and optionally the fused-feature gradients.  Indices */
/* outside [0, vocab) / [0, n_pos) are skipped (never scattered) and flagged  */
/* in err_flag (nullable), like the forward does.                             */
int ergm_embed_bwd(const float* dh, const int64_t* ids, const int64_t* token_type_ids,
                   const int64_t* position_ids, int64_t pos_stride_b, float* dwte, float* dwpe, float* dimgs,
                   float* dauds, int rows, int T, int H, int past_len, int vocab, int n_pos, float dropout_p,
                   uint64_t seed, uint64_t offset, int* err_flag, const ergm_pack* pack, void* stream);

/* ------------------------------------------------------------------------ */
/* LayerNorm (model.py:298,318,332,578; eps inside the sqrt, biased variance) */
/* fwd writes bf16 and/or fp32 outputs and the row statistics.                */
int ergm_ln_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16,
                float* y_f32, float* mean, float* rstd, int rows, int H, float eps,
                const int* row_idx /* nullable: output row r normalises x[row_idx[r]] */,
                const int* rows_dyn /* nullable device int: run-time row count (packed batch) */, void* stream);
/* bwd fused with the residual-gradient add: dx_out = dres_in + LN'(dy);      */
/* dx_bf16 = bf16(dropout_mask(dx_out)) feeds the next dgrad/wgrad GEMMs;     */
/* dgamma/dbeta/dbias_next are accumulated (+=).                              */
int ergm_ln_bwd(const void* dy, int dy_is_f32, const float* x, const float* mean,
                const float* rstd, const float* gamma, const float* dres_in, float* dx_out,
                void* dx_bf16, float* dgamma, float* dbeta, float* dbias_next, int rows, int H,
                float dropout_p, uint64_t seed, uint64_t offset, const int* rows_dyn /* nullable */, void* stream);
/* out[N] += column sums of a bf16 [rows, N] matrix (Conv1D bias gradients)   */
int ergm_colsum_bf16(const void* src, int64_t ld, int rows, int N, float* out, void* stream);
/* MLP backward through gelu_new (model.py:264): dg (bf16 [rows, ld]) is overwritten with
 * dg * gelu_new'(u); colsum[N] += column sums of the result (c_fc bias gradient).   */
int ergm_gelu_bwd_colsum(void* dg_bf16, const void* u_bf16, int64_t ld, int rows, int N, float* colsum,
                         int exact, const int* rows_dyn /* nullable */, void* stream);
int ergm_cast_f32_bf16_2d(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int rows,
                          int N, float* colsum, const int* rows_dyn /* nullable */, void* stream);
int ergm_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream);

/* ------------------------------------------------------------------------ */
/* Fused attention, head_dim 64 (GPT2Attention._attn, model.py:119-148, and   */
/* the head split / merge permutes :190-198).  q/k/v are bf16 matrices        */
/* [B*T, ld] whose head h occupies columns [col0 + 64h, col0 + 64h + 64);     */
/* out is [B*Tq, ld_out] merged heads; lse is [B, nh, Tq] (natural log).      */
/* causal: query i sees key j iff j <= i + causal_off (self-attention);       */
/* kv_lens (nullable, int32 [B]) masks right-padded keys.                     */
int ergm_attn_fwd(const void* q, int64_t ld_q, int q_col0, const void* k, int64_t ld_k,
                  int k_col0, const void* v, int64_t ld_v, int v_col0, void* out, int64_t ld_out,
                  float* out_f32 /* nullable [B*Tq, nh*64]: un-rounded copy for backward */,
                  float* lse, const int* kv_lens, int B, int nh, int Tq, int Tk, int head_dim,
                  int causal, int causal_off, float dropout_p, uint64_t seed, uint64_t offset,
                  const ergm_pack* pack /* nullable: q / out rows of sample b start at cu_rows[b] (capacity B*Tq) */,
                  int pack_kv /* 1: k / v are packed the same way and kv_lens = pack->kv_lens (self attention) */,
                  void* stream);

/* Backward of ergm_attn_fwd: recomputes P from lse; dq / dk / dv are written (bf16) at [B*Tq | B*Tk, ld] column
 * offsets dq_col0 / dk_col0 / dv_col0 (+64h); delta ([B,nh,Tq] scratch) receives rowsum(dO * O).  dq_colsum /
 * dk_colsum / dv_colsum (nullable fp32 [nh*64]) are incremented by the column sums of dQ / dK / dV as stored, i.e.
 * the bias gradients of the Q / K / V projections.  Persistent kernel, one CTA per SM.  Tq <= 256: a CTA owns whole
 * (batch, head) items, dQ accumulates in TMEM and no workspace is needed; longer sequences: one item per 128-key
 * block, dQ contributions meet in an fp32 workspace (ergm_attn_bwd_workspace_bytes; zeroed, cast and column-summed by
 * this call).  Never allocates.                                               */
int ergm_attn_bwd_workspace_bytes(int B, int nh, int Tq, int64_t* bytes);
int ergm_attn_bwd(const void* q, int64_t ld_q, int q_col0, const void* k, int64_t ld_k,
                  int k_col0, const void* v, int64_t ld_v, int v_col0, const void* out,
                  int64_t ld_out, const float* out_f32 /* nullable */, const void* dout, int64_t ld_do, const float* lse, float* delta,
                  void* dq, int64_t ld_dq, int dq_col0, void* dk, int64_t ld_dk, int dk_col0, void* dv,
                  int64_t ld_dv, int dv_col0, float* dq_colsum, float* dk_colsum, float* dv_colsum, const int* kv_lens,
                  int B, int nh, int Tq, int Tk, int head_dim, int causal, int causal_off, float dropout_p,
                  uint64_t seed, uint64_t offset, const ergm_pack* pack, int pack_kv,
                  void* workspace /* 16-byte aligned; may be NULL when the query returns 0 */, int64_t workspace_bytes,
                  void* stream);

/* ------------------------------------------------------------------------ */
/* Token cross-entropy over LM-head logits with the reference's shift and     */
/* ignore_index=-100 (model.py:705-708): row (b,t) is scored against          */
/* labels[b,t+1].  sums[0] += sum of row losses, sums[1] += valid rows.       */
/* hn / w (nullable, bf16 [rows,H] / [V,H]): when given, the target logit is
 * recomputed as their fp32 dot product instead of read from rounded logits.
 * T == 0: labels[row] is the target of row `row` (compacted rows of the label-sparse head).
 * rows_dyn (nullable device int): rows >= roundup(*rows_dyn, 128) are not touched.            */
int ergm_ce_fwd(const void* logits, int logits_is_f32, int64_t ldl, const int64_t* labels,
                int rows, int T, int V, float* lse, float* row_loss, float* sums, int* err_flag,
                const void* hn_bf16, const void* w_bf16, int H, const int* rows_dyn, void* stream);
/* dlogits = (softmax - onehot) * (*scale_ptr), zero for ignored rows (bf16)   */
int ergm_ce_bwd(const void* logits, int logits_is_f32, int64_t ldl, const int64_t* labels,
                int rows, int T, int V, const float* lse, const float* scale_ptr,
                void* dlogits_bf16, int64_t ldd, const int* rows_dyn, void* stream);
/* Label-sparse LM head (model.py:698-708 scores every position, then ignores the -100 ones): the head, its
 * CE and their backward run on the compacted scored rows only.                                  */
/* plan: ordered compaction of the rows whose shifted label is not -100 -> row_idx[i] (int32 source row),
 * labels_c[i] (its target), *count; entries up to the next multiple of 128 are padded (-1 / -100).  */
int ergm_lm_rows_plan(const int64_t* labels, int rows, int T, int* row_idx, int64_t* labels_c, int* count,
                      const ergm_pack* pack /* nullable: rows are packed rows, labels stay [B, T] */, void* stream);
/* dst[i] = src[row_idx[i]] (bf16 [.., H]) for i < *count, zero rows up to the next multiple of 128 */
int ergm_gather_rows_dyn(const void* src_bf16, const int* row_idx, const int* count, void* dst_bf16, int H,
                         int cap, void* stream);
/* dst[row_idx[i]] = src[i] (fp32 [.., H]) for i < *count */
int ergm_scatter_rows_dyn(const float* src, const int* row_idx, const int* count, float* dst, int H,
                          void* stream);
/* LM head + CE as one call each way (replaces model.py:698 lm_head + :705-708 shift / CrossEntropyLoss and what
 * autograd derives from them): plan -> gather -> tcgen05 GEMM over the scored rows -> CE; backward: dlogits,
 * d hn (fp32 [rows, H], OVERWRITTEN: zero for rows that are not scored), d wte (fp32 [V, H], ACCUMULATED).
 * hn: bf16 [rows, H] ln_f output; wte: bf16 [V, H] (the tied head weight); labels: int64 [B, T] as given to the model
 * (rows = B * T, or the packed rows of `pack`); sums as ergm_ce_fwd.  Never allocates: scratch and the forward state
 * the backward reads live in `workspace` (256-byte aligned, >= ergm_lmhead_ce_workspace_bytes(rows, H, V, 1) when
 * the backward follows; the backward must get the SAME workspace, untouched in between).  The layout query returns
 * the byte offsets of the sub-buffers (ERGM_LMHEAD_WS_*; [END] = total) so that a caller can read the device row
 * count, the compacted logits or the per-row losses.                                              */
enum {
  ERGM_LMHEAD_WS_COUNT = 0,    /* int32 [1]: scored rows of this batch (device) */
  ERGM_LMHEAD_WS_ROW_IDX = 1,  /* int32 [rows]: source row of compacted row i */
  ERGM_LMHEAD_WS_LABELS = 2,   /* int64 [rows]: its target */
  ERGM_LMHEAD_WS_HN = 3,       /* bf16 [rows, H]: gathered ln_f rows */
  ERGM_LMHEAD_WS_LOGITS = 4,   /* bf16 [rows, roundup(V, 64)]: logits of the scored rows */
  ERGM_LMHEAD_WS_LSE = 5,      /* fp32 [rows] */
  ERGM_LMHEAD_WS_ROW_LOSS = 6, /* fp32 [rows] */
  ERGM_LMHEAD_WS_DLOGITS = 7,  /* bf16 [rows, roundup(V, 64)] (backward) */
  ERGM_LMHEAD_WS_DHN = 8,      /* fp32 [rows, H] (backward) */
  ERGM_LMHEAD_WS_END = 9
};
int ergm_lmhead_ce_workspace_bytes(int rows, int H, int V, int with_backward, int64_t* bytes);
int ergm_lmhead_ce_workspace_layout(int rows, int H, int V, int with_backward, int64_t* offsets /* [END + 1] */);
int ergm_lmhead_ce_fwd(const void* hn_bf16, const void* wte_bf16, const int64_t* labels, int rows, int T, int H,
                       int V, float* sums, int* err_flag, const ergm_pack* pack /* nullable */, void* workspace,
                       int64_t workspace_bytes, void* stream);
int ergm_lmhead_ce_bwd(const void* wte_bf16, const float* scale_ptr /* device: d loss / d row loss */, int rows, int H,
                       int V, float* dhn, float* dwte, void* workspace, int64_t workspace_bytes, void* stream);
/* Emotion head on the last position + 7-way CE (model.py:700-701,710-711).   */
/* sums[2] += sum of sample losses, sums[3] += samples.                       */
int ergm_emotion_head_fwd(const float* x_final, const float* mean, const float* rstd,
                          const float* gamma, const float* beta, const float* w_emo,
                          const int64_t* emotion_labels, int B, int T, int H, float* hlast,
                          float* logits, float* dlogits, float* sums, int* err_flag,
                          const int* cu_rows /* nullable: packed batch, sample b's last row is cu_rows[b+1]-1 */,
                          void* stream);
int ergm_emotion_head_bwd(const float* dlogits, const float* hlast, const float* w_emo,
                          const float* scale_ptr, int B, int T, int H, float* dw_emo, float* dyf,
                          const int* cu_rows /* nullable */, void* stream);
/* out = [loss, lm_loss, emo_loss, 1/lm_valid, 1/emo_count] (model.py:713)    */
int ergm_loss_finalize(const float* sums, int has_lm, int has_emotion, float* out, void* stream);
int ergm_scalar_mul(const float* a, const float* b, float* dst, void* stream);

/* ------------------------------------------------------------------------ */
/* fp32 mode (logits within 1e-4, bit-exact greedy ids): fp32 operands are split
 * exactly into three bf16 pieces and expanded 6x along the reduction dimension
 * so that ONE ergm_gemm_bf16 call computes an fp32-accurate product
 * (see csrc/fp32_mode.cu).  side 0 = A operand, 1 = B operand; kdim = which
 * dimension of src is the reduction dimension (1: cols -> dst [rows, 6*cols];
 * 0: rows -> dst [6*rows, cols]).                                            */
int ergm_split3_expand(const float* src, int64_t ld_src, void* dst_bf16, int64_t ld_dst, int rows,
                       int cols, int side, int kdim, void* stream);
/* fp32 CUDA-core attention (verification path of model.py:119-148)            */
int ergm_attn_fwd_f32(const float* q, int64_t ld_q, int q_col0, const float* k, int64_t ld_k,
                      int k_col0, const float* v, int64_t ld_v, int v_col0, float* out, int64_t ld_out,
                      const int* kv_lens, int B, int nh, int Tq, int Tk, int head_dim, int causal,
                      int causal_off, void* stream);

/* The whole block stack of one decode step as ONE persistent cooperative kernel (148 CTAs, in-kernel grid */
/* barriers instead of ~60 dependent launches; model.py:286-341 per block).  layer_table: device array of  */
/* L records of 14 pointers each: {w_qkv, b_qkv, w_o, b_o, w_q2, b_q2, w_o2, b_o2, w_fc, b_fc, w_p2, b_p2,  */
/* kv_pool, kv2} - packed weights / folded biases from ergm_dec_pack_weight, this layer's paged K/V pool,   */
/* its cached cross-attention K/V [B*Tc, 2H] (NULL pointers when Tc == 0).  x: fp32 [B, H] residual stream  */
/* (in: embeddings; out: last block's output).  qkv / ctx / q2 / g: bf16 scratch [B, 3H] / [B, H] / [B, H] / */
/* [B, I].  sync_ctr: device uint32 used by the grid barrier.  B <= 64.                                      */
int ergm_decode_layers(const void* layer_table, int L, int H, int I, int nh, int B, float* x, void* qkv,
                       void* ctx, void* q2, void* g, const int* block_table, const int* seq_lens,
                       int max_pages, int Tc, float eps, unsigned int* sync_ctr, void* stream);
/* Decode-step GEMMs (M = batch <= 64 rows, one new token per sequence): every      */
/* Conv1D / Linear of the block (model.py:218-222,244,263,265) and the tied LM head  */
/* (model.py:698) is weight-streaming bound at this M.  ergm_dec_pack_weight re-packs */
/* an fp32 weight (logical [K, N]: stored [K, ld] row-major as Conv1D does, or, with  */
/* w_is_nk = 1, [N, ld] as nn.Linear / wte do) into bf16 16-column slabs in mma       */
/* B-fragment order: packed holds ceil(N/16)*16*K bf16, K % 16 == 0.  When the GEMM   */
/* consumes a LayerNorm output (model.py:298,318,332,578) the affine parameters are   */
/* folded in here: packed = bf16(diag(gamma) W), bias_out = bias_in + beta @ W        */
/* (gamma / beta / bias_in nullable; bias_out [N] required when beta or bias_in).     */
/* ------------------------------------------------------------------------ */
/* ergm_decode_stack: ALL transformer blocks of one decode step (L x model.py:286-341 on one new token per
 * sequence, B <= 64, no captions) in one persistent kernel of head-clusters (8 CTAs per head): two grid-wide
 * synchronisations per block, everything else through distributed shared memory (csrc/decode_step.cu).
 * Weights: per layer three blob arrays made by ergm_decode_stack_pack (sizes from ergm_decode_stack_blob_bytes);
 * ln_1 / ln_2 gamma are folded into the q|k|v / fc blobs, their beta into the biases the caller passes (the
 * folded biases ergm_dec_pack_weight returns).
 * layer_table: device array of L records of 8 pointers {p1 blobs, fc blobs, proj blobs, b_qkv[3H], b_o[H],
 * b_fc[I], b_p2[H], K/V page pool}.  x0 holds the embeddings; x1 / x2 are scratch of the same size; the block
 * stack's output is in x{(2L) % 3}.  Appends the new token's K / V at slot seq_lens[b] like
 * ergm_attn_decode_paged.  sync_ctr: one device uint32 of scratch.  trace: NULL, or (profiling aid) a device
 * int64[16 * L] that receives %globaltimer stamps of CTA 0 at the phase boundaries of every block.
 * ERGM_ERR_UNSUPPORTED (-2): geometry outside H in {128,256,512,768}, I = 4H, head_dim 64, B <= 64, or the device
 * cannot hold all clusters at once: callers fall back to the per-kernel chain.                          */
int ergm_decode_stack_blob_bytes(int H, int I, int nh, int64_t* p1_bytes, int64_t* fc_bytes, int64_t* pj_bytes);
int ergm_decode_stack_pack(const float* w_qkv, const float* gamma1, const float* w_o, const float* w_fc,
                           const float* gamma2, const float* w_p2, int H, int I, int nh, void* p1_blobs,
                           void* fc_blobs, void* pj_blobs, void* stream);
int ergm_decode_stack(const void* layer_table, int L, int H, int I, int nh, int B, float* x0, float* x1, float* x2,
                      const int* block_table, const int* seq_lens, int max_pages, float eps, uint32_t* sync_ctr,
                      int64_t* trace, void* stream);
int ergm_dec_pack_weight(const float* w_f32, int64_t ld, int K, int N, int w_is_nk, const float* gamma,
                         const float* beta, const float* bias_in, void* packed, float* bias_out,
                         void* stream);
/* out[M, N] = epi(A @ Wp):  A = (x - mean) * rstd per row of x_f32[M, lda] computed  */
/* in the prologue (x_f32 != NULL; gamma / beta live in Wp / bias), or the bf16       */
/* matrix a_bf16[M, lda].  out_mode 0: bf16 store (+bias, +gelu_new if gelu);         */
/* 1: fp32 store (+bias); 2: fp32 "+=" into out (residual stream, model.py:309,329,   */
/* 334; bias added once; K split across CTAs, fp32 reductions; a_bf16 only).          */
int ergm_dec_gemm(const float* x_f32, const void* a_bf16, int64_t lda, float eps, const void* w_packed,
                  int K, int N, const float* bias, void* out, int64_t ldo, int out_mode, int gelu,
                  int M, void* stream);
/* ------------------------------------------------------------------------ */
/* Decode: paged KV cache + one-query attention + on-device sampling.  Replaces
 * the torch.cat cache growth of model.py:228-236 and the per-token sampling /
 * host sync of main.py:253-282.  Pool layout (bf16):
 * [page][k|v][head][16 tokens][64]; block_table is int32 [B, max_pages].     */
/* appends the new token's K/V (columns k_col0 / v_col0 of qkv [B, ld_q]) at
 * slot seq_lens[b] and attends over seq_lens[b] + 1 tokens.                  */
int ergm_attn_decode_paged(const void* qkv, int64_t ld_q, int q_col0, int k_col0, int v_col0,
                           void* pool, const int* block_table, const int* seq_lens, int max_pages,
                           void* out, int64_t ld_out, int B, int nh, int head_dim, void* stream);
/* one-query attention over a contiguous [B*Tk, ld_k] K/V matrix (cached
 * cross-attention K/V of the caption embeddings, model.py:219)               */
int ergm_attn_decode_contig(const void* q, int64_t ld_q, int q_col0, const void* kv, int64_t ld_k,
                            int k_col0, int v_col0, const int* kv_lens, void* out, int64_t ld_out,
                            int B, int nh, int Tk, int head_dim, void* stream);
/* prompt K/V rows -> pages (rows t >= lens[b] are skipped when lens != NULL)  */
int ergm_kv_to_pages(const void* kv, int64_t ld, int k_col0, int v_col0, void* pool,
                     const int* block_table, const int* lens, int max_pages, int B, int T, int nh,
                     void* stream);
/* next token per row of fp32 logits: top_k 0 or 1 = greedy arg-max (lowest index
 * on ties); top_k >= 2 = top-k / temperature sampling (Philox(seed, *step_ptr));
 * top_k == -1 (ERGM_SAMPLE_ALL) = multinomial over the whole distribution
 * (the nucleus kernel at top_p = 1), with temperature.
 * Writes out_ids[b, *step_ptr], next_ids[b]; finished rows emit eos_id;
 * seq_lens[b] += 1.                                                          */
#define ERGM_SAMPLE_ALL (-1)
int ergm_sample(const float* logits, int64_t ld, int B, int V, int top_k,
                float top_p /* < 1: nucleus sampling with main.py:263-265's shifted mask; 1 = off */,
                float temperature, uint64_t seed, int* step_ptr,
                int advance_step /* 1: *step_ptr += 1 once every row has been sampled */,
                int64_t* out_ids, int64_t out_ld, int64_t* next_ids, int* finished, int* seq_lens,
                int64_t eos_id, void* stream);
int ergm_int_add(int* dev_ptr, int inc, void* stream);

/* ------------------------------------------------------------------------ */
/* Flat AdamW, torch.optim.AdamW arithmetic (main.py:68,155).  hyper (device) */
/* = [lr, beta1, beta2, eps, weight_decay, 1-beta1^t, 1-beta2^t]; also writes */
/* the bf16 weight shadow consumed by the GEMMs.  g: fp32, or bf16 when        */
/* g_is_bf16 (data-parallel gradient buckets all-reduced in bf16). */
int ergm_adamw_flat(float* p, const void* g, int g_is_bf16, float* m, float* v, void* shadow_bf16, int64_t n,
                    const float* hyper, const float* grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ERGM_B200_H_ */

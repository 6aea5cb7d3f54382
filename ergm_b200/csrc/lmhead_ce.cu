// LM head + token cross-entropy as ONE C-ABI call each way (model.py:698-708).
//
// The reference computes lm_head(h) for every position, copies the shifted logits and lets CrossEntropyLoss ignore
// the -100 rows.  Here the head, its CE and their backward run on the scored rows only: plan (ordered compaction of
// the rows whose shifted label is not -100, run-time count on the device) -> gather of the ln_f rows -> tcgen05 GEMM
// bounded by the device count -> CE.  Backward: dlogits on the compacted rows, d hn_c = dlogits_c @ wte (the kernel
// splits K = V itself), scatter to the full rows, d wte += dlogits_c^T @ hn_c (run-time K).  Nothing here allocates
// or synchronises: all scratch (and the forward state the backward needs) lives in the caller's workspace whose size
// and layout ergm_lmhead_ce_workspace_bytes / _layout report.
#include "../../include/ergm_b200.h"
#include "common.cuh"

namespace {

constexpr int64_t WS_ALIGN = 256;
inline int64_t up(int64_t x) { return (x + WS_ALIGN - 1) / WS_ALIGN * WS_ALIGN; }
inline int64_t ldl_of(int V) { return ((int64_t)V + 63) / 64 * 64; }  // 16-byte rows for TMA / vector access

// offsets[ERGM_LMHEAD_WS_*]; offsets[ERGM_LMHEAD_WS_END] = total bytes
void layout(int rows, int H, int V, int with_backward, int64_t* off) {
  const int64_t ldl = ldl_of(V);
  int64_t o = 0;
  off[ERGM_LMHEAD_WS_COUNT] = o;     o += up(sizeof(int32_t));
  off[ERGM_LMHEAD_WS_ROW_IDX] = o;   o += up((int64_t)rows * sizeof(int32_t));
  off[ERGM_LMHEAD_WS_LABELS] = o;    o += up((int64_t)rows * sizeof(int64_t));
  off[ERGM_LMHEAD_WS_HN] = o;        o += up((int64_t)rows * H * 2);
  off[ERGM_LMHEAD_WS_LOGITS] = o;    o += up((int64_t)rows * ldl * 2);
  off[ERGM_LMHEAD_WS_LSE] = o;       o += up((int64_t)rows * sizeof(float));
  off[ERGM_LMHEAD_WS_ROW_LOSS] = o;  o += up((int64_t)rows * sizeof(float));
  off[ERGM_LMHEAD_WS_DLOGITS] = o;   o += with_backward ? up((int64_t)rows * ldl * 2) : 0;
  off[ERGM_LMHEAD_WS_DHN] = o;       o += with_backward ? up((int64_t)rows * H * sizeof(float)) : 0;
  off[ERGM_LMHEAD_WS_END] = o;
}

#define LM_TRY(call)            \
  do {                          \
    const int rc_ = (call);     \
    if (rc_ != ERGM_OK) return rc_; \
  } while (0)

}  // namespace

extern "C" int ergm_lmhead_ce_workspace_bytes(int rows, int H, int V, int with_backward, int64_t* bytes) {
  if (rows <= 0 || H <= 0 || V <= 0 || !bytes) return ERGM_ERR_ARG;
  int64_t off[ERGM_LMHEAD_WS_END + 1];
  layout(rows, H, V, with_backward, off);
  *bytes = off[ERGM_LMHEAD_WS_END];
  return ERGM_OK;
}

extern "C" int ergm_lmhead_ce_workspace_layout(int rows, int H, int V, int with_backward, int64_t* offsets) {
  if (rows <= 0 || H <= 0 || V <= 0 || !offsets) return ERGM_ERR_ARG;
  layout(rows, H, V, with_backward, offsets);
  return ERGM_OK;
}

extern "C" int ergm_lmhead_ce_fwd(const void* hn_bf16, const void* wte_bf16, const int64_t* labels, int rows, int T,
                                  int H, int V, float* sums, int* err_flag, const ergm_pack* pack, void* workspace,
                                  int64_t workspace_bytes, void* stream) {
  if (!hn_bf16 || !wte_bf16 || !labels || !sums || !workspace || rows <= 0 || T <= 0 || H <= 0 || V <= 0)
    return ERGM_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(workspace) % WS_ALIGN) return ERGM_ERR_ARG;
  int64_t off[ERGM_LMHEAD_WS_END + 1];
  layout(rows, H, V, 0, off);
  if (workspace_bytes < off[ERGM_LMHEAD_WS_END]) return ERGM_ERR_ARG;
  char* ws = static_cast<char*>(workspace);
  int* count = reinterpret_cast<int*>(ws + off[ERGM_LMHEAD_WS_COUNT]);
  int* row_idx = reinterpret_cast<int*>(ws + off[ERGM_LMHEAD_WS_ROW_IDX]);
  int64_t* labels_c = reinterpret_cast<int64_t*>(ws + off[ERGM_LMHEAD_WS_LABELS]);
  void* hn_c = ws + off[ERGM_LMHEAD_WS_HN];
  void* logits_c = ws + off[ERGM_LMHEAD_WS_LOGITS];
  float* lse = reinterpret_cast<float*>(ws + off[ERGM_LMHEAD_WS_LSE]);
  float* row_loss = reinterpret_cast<float*>(ws + off[ERGM_LMHEAD_WS_ROW_LOSS]);
  const int64_t ldl = ldl_of(V);

  LM_TRY(ergm_lm_rows_plan(labels, rows, T, row_idx, labels_c, count, pack, stream));
  LM_TRY(ergm_gather_rows_dyn(hn_bf16, row_idx, count, hn_c, H, rows, stream));
  ergm_gemm_args g = {};
  g.a = hn_c; g.b = wte_bf16; g.d = logits_c;
  g.lda = H; g.ldb = H; g.ldd = ldl;
  g.M = rows; g.N = V; g.K = H;
  g.a_major = ERGM_MAJOR_K; g.b_major = ERGM_MAJOR_K;
  g.d_dtype = ERGM_DT_BF16;
  g.split_k = 1;
  g.dyn_count = count; g.dyn_dim = 1;
  LM_TRY(ergm_gemm_bf16(&g, stream));
  // the target logit is recomputed in fp32 from the GEMM operands (hn_c, wte): the loss does not see bf16 rounding of it
  return ergm_ce_fwd(logits_c, 0, ldl, labels_c, rows, /*T=*/0, V, lse, row_loss, sums, err_flag, hn_c, wte_bf16, H,
                     count, stream);
}

extern "C" int ergm_lmhead_ce_bwd(const void* wte_bf16, const float* scale_ptr, int rows, int H, int V, float* dhn,
                                  float* dwte, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!wte_bf16 || !scale_ptr || !dhn || !dwte || !workspace || rows <= 0 || H <= 0 || V <= 0) return ERGM_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(workspace) % WS_ALIGN) return ERGM_ERR_ARG;
  int64_t off[ERGM_LMHEAD_WS_END + 1];
  layout(rows, H, V, 1, off);
  if (workspace_bytes < off[ERGM_LMHEAD_WS_END]) return ERGM_ERR_ARG;
  char* ws = static_cast<char*>(workspace);
  const int* count = reinterpret_cast<const int*>(ws + off[ERGM_LMHEAD_WS_COUNT]);
  const int* row_idx = reinterpret_cast<const int*>(ws + off[ERGM_LMHEAD_WS_ROW_IDX]);
  const int64_t* labels_c = reinterpret_cast<const int64_t*>(ws + off[ERGM_LMHEAD_WS_LABELS]);
  const void* hn_c = ws + off[ERGM_LMHEAD_WS_HN];
  const void* logits_c = ws + off[ERGM_LMHEAD_WS_LOGITS];
  const float* lse = reinterpret_cast<const float*>(ws + off[ERGM_LMHEAD_WS_LSE]);
  void* dlogits_c = ws + off[ERGM_LMHEAD_WS_DLOGITS];
  float* dhn_c = reinterpret_cast<float*>(ws + off[ERGM_LMHEAD_WS_DHN]);
  const int64_t ldl = ldl_of(V);
  cudaStream_t st = (cudaStream_t)stream;

  LM_TRY(ergm_ce_bwd(logits_c, 0, ldl, labels_c, rows, /*T=*/0, V, lse, scale_ptr, dlogits_c, ldl, count, stream));
  cudaError_t ce = cudaMemsetAsync(dhn_c, 0, (size_t)rows * H * sizeof(float), st);
  if (ce != cudaSuccess) return (int)ce;
  ce = cudaMemsetAsync(dhn, 0, (size_t)rows * H * sizeof(float), st);
  if (ce != cudaSuccess) return (int)ce;
  // d hn_c = dlogits_c @ wte: few row tiles, very long K (= V): the pair kernel splits K itself to fill the GPU
  ergm_gemm_args g = {};
  g.a = dlogits_c; g.b = wte_bf16; g.d = dhn_c;
  g.lda = ldl; g.ldb = H; g.ldd = H;
  g.M = rows; g.N = H; g.K = V;
  g.a_major = ERGM_MAJOR_K; g.b_major = ERGM_MAJOR_MN;
  g.d_dtype = ERGM_DT_F32;
  g.epilogue = ERGM_EPI_ATOMIC;
  g.split_k = 1; g.block_n = 2256;
  g.dyn_count = count; g.dyn_dim = 1; g.dyn_hint = -1;
  LM_TRY(ergm_gemm_bf16(&g, stream));
  LM_TRY(ergm_scatter_rows_dyn(dhn_c, row_idx, count, dhn, H, stream));
  // d wte += dlogits_c^T @ hn_c: reduction over the run-time row count
  ergm_gemm_args w = {};
  w.a = dlogits_c; w.b = hn_c; w.d = dwte;
  w.lda = ldl; w.ldb = H; w.ldd = H;
  w.M = V; w.N = H; w.K = rows;
  w.a_major = ERGM_MAJOR_MN; w.b_major = ERGM_MAJOR_MN;
  w.d_dtype = ERGM_DT_F32;
  w.epilogue = ERGM_EPI_ATOMIC;
  w.split_k = 1; w.block_n = 2256;
  w.dyn_count = count; w.dyn_dim = 2;
  return ergm_gemm_bf16(&w, stream);
}

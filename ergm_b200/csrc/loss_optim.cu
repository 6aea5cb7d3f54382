// Loss-side and optimiser kernels of the ERGM path (all HBM-bound, vectorised, coalesced):
//   * token cross-entropy over the LM-head logits with the reference's shift-by-one and
//     ignore_index = -100 semantics (model.py:705-708 / :715-718), forward and backward;
//   * the 7-way emotion head on the last position + its cross-entropy (model.py:700-701,
//     710-711), forward and backward;
//   * loss finalisation  loss = CE_lm + CE_emotion (model.py:713);
//   * flat multi-tensor AdamW with torch.optim.AdamW arithmetic (main.py:68,155) that also
//     refreshes the bf16 weight shadow used by the tensor-core GEMMs.
#include "../../include/ergm_b200.h"
#include "common.cuh"

namespace ergm {

constexpr int CE_THREADS = 256;

ERGM_DEVINL void online_merge(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  if (mn == -INFINITY) { s = 0.f; return; }  // both partials empty: exp(-inf - -inf) would be NaN
  s = s * __expf(m - mn) + s2 * __expf(m2 - mn);
  m = mn;
}

template <bool F32>
ERGM_DEVINL void load8(const void* row, int chunk, float (&v)[8]) {
  if (F32) {
    const float4 a = reinterpret_cast<const float4*>(row)[2 * chunk];
    const float4 b = reinterpret_cast<const float4*>(row)[2 * chunk + 1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 u = reinterpret_cast<const uint4*>(row)[chunk];
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
}
template <bool F32>
ERGM_DEVINL float load1(const void* row, int col) {
  return F32 ? reinterpret_cast<const float*>(row)[col]
             : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(row)[col]);
}

// target of logits row (b, t) is labels[b, t+1]; the last position has none.  T == 0: `labels` is already
// aligned with the rows (compacted label-sparse LM head: row i scores labels[i]).
ERGM_DEVINL int64_t shifted_target(const int64_t* labels, int row, int T) {
  if (T == 0) return labels[row];
  const int t = row % T;
  return (t + 1 < T) ? labels[row + 1] : (int64_t)-100;
}
// rows at or beyond the 128-aligned run-time row count take no part at all (their buffers are never read)
ERGM_DEVINL bool row_beyond_dyn(const int* rows_dyn, int row) {
  return rows_dyn && row >= ((*rows_dyn + 127) & ~127);
}

// One CTA per row.  Rows whose target is ignore_index are skipped (lse = 0, loss = 0).
template <bool F32>
__global__ void __launch_bounds__(CE_THREADS)
ce_fwd_kernel(const void* __restrict__ logits, int64_t ldl, const int64_t* __restrict__ labels,
              int T, int V, float* __restrict__ lse_out, float* __restrict__ row_loss,
              float* __restrict__ sums /* [0]=loss sum, [1]=valid count */, int* err_flag,
              const __nv_bfloat16* __restrict__ hn, const __nv_bfloat16* __restrict__ w, int H,
              const int* __restrict__ rows_dyn) {
  __shared__ float sm[CE_THREADS / 32], ss[CE_THREADS / 32], st[CE_THREADS / 32];
  const int row = blockIdx.x;
  if (row_beyond_dyn(rows_dyn, row)) return;
  const int64_t tgt = shifted_target(labels, row, T);
  if (tgt == -100) {
    if (threadIdx.x == 0) { lse_out[row] = 0.f; row_loss[row] = 0.f; }
    return;
  }
  if (tgt < 0 || tgt >= V) {
    if (threadIdx.x == 0) { *err_flag = 1; lse_out[row] = 0.f; row_loss[row] = 0.f; }
    return;
  }
  const char* rp = reinterpret_cast<const char*>(logits) + (int64_t)row * ldl * (F32 ? 4 : 2);
  float m = -INFINITY, s = 0.f;
  const int nfull = V / 8;
  for (int c = threadIdx.x; c < nfull; c += CE_THREADS) {
    float v[8];
    load8<F32>(rp, c, v);
    float mx = v[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, v[i]);
    const float mn = fmaxf(m, mx);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += __expf(v[i] - mn);
    s = s * __expf(m - mn) + acc;
    m = mn;
  }
  for (int c = nfull * 8 + threadIdx.x; c < V; c += CE_THREADS) online_merge(m, s, load1<F32>(rp, c), 1.f);
  // Target logit from the fp32 dot product of the GEMM operands (the stored logits may be bf16-
  // rounded; the loss must come from fp32 accumulation).
  float tl = 0.f;
  if (hn) {
    const uint4* hr = reinterpret_cast<const uint4*>(hn + (int64_t)row * H);
    const uint4* wr = reinterpret_cast<const uint4*>(w + tgt * H);
    for (int c = threadIdx.x; c < H / 8; c += CE_THREADS) {
      const uint4 a = hr[c], b = wr[c];
      const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
      const float2 b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y), b2 = unpack_bf16x2(b.z), b3 = unpack_bf16x2(b.w);
      tl += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
    }
    tl = warp_sum(tl);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    online_merge(m, s, m2, s2);
  }
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m; ss[threadIdx.x >> 5] = s; st[threadIdx.x >> 5] = tl; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = sm[0], S = ss[0], TL = st[0];
    for (int w = 1; w < CE_THREADS / 32; ++w) { online_merge(M, S, sm[w], ss[w]); TL += st[w]; }
    const float lse = M + logf(S);
    const float loss = lse - (hn ? TL : load1<F32>(rp, (int)tgt));
    lse_out[row] = lse;
    row_loss[row] = loss;
    atomicAdd(sums, loss);
    atomicAdd(sums + 1, 1.f);
  }
}

// dlogits[row, v] = (softmax(logits[row])[v] - [v == target]) * scale ; zero for ignored rows.
// scale is read from device memory (= upstream grad / global valid count).
template <bool F32>
__global__ void __launch_bounds__(CE_THREADS)
ce_bwd_kernel(const void* __restrict__ logits, int64_t ldl, const int64_t* __restrict__ labels,
              int T, int V, const float* __restrict__ lse, const float* __restrict__ scale_ptr,
              __nv_bfloat16* __restrict__ dlogits, int64_t ldd, const int* __restrict__ rows_dyn) {
  const int row = blockIdx.x;
  if (row_beyond_dyn(rows_dyn, row)) return;
  const int64_t tgt = shifted_target(labels, row, T);
  __nv_bfloat16* dp = dlogits + (int64_t)row * ldd;
  const int nvec = (int)(ldd / 8);  // whole padded row is written so the pad stays finite
  if (tgt < 0 || tgt >= V) {
    for (int c = threadIdx.x; c < nvec; c += CE_THREADS)
      reinterpret_cast<uint4*>(dp)[c] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const char* rp = reinterpret_cast<const char*>(logits) + (int64_t)row * ldl * (F32 ? 4 : 2);
  const float l = lse[row], scale = *scale_ptr;
  const int nfull = V / 8;
  for (int c = threadIdx.x; c < nvec; c += CE_THREADS) {
    float v[8];
    if (c < nfull) {
      load8<F32>(rp, c, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __expf(v[i] - l);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int col = c * 8 + i;
        v[i] = col < V ? __expf(load1<F32>(rp, col) - l) : 0.f;
      }
    }
    const int rel = (int)tgt - c * 8;
    if (rel >= 0 && rel < 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i == rel) v[i] -= 1.f;
    }
    reinterpret_cast<uint4*>(dp)[c] =
        make_uint4(pack_bf16x2(v[0] * scale, v[1] * scale), pack_bf16x2(v[2] * scale, v[3] * scale),
                   pack_bf16x2(v[4] * scale, v[5] * scale), pack_bf16x2(v[6] * scale, v[7] * scale));
  }
}

// ------------------------------------------------------------------------------------------
// emotion head: one warp per sample
// ------------------------------------------------------------------------------------------
constexpr int NUM_EMO = 7;  // model.py:607

__global__ void __launch_bounds__(128)
emotion_fwd_kernel(const float* __restrict__ x_final, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ w_emo,
                   const int64_t* __restrict__ emo_labels, int B, int T, int H,
                   float* __restrict__ hlast, float* __restrict__ logits_out,
                   float* __restrict__ dlogits_out, float* __restrict__ sums, int* err_flag,
                   const int* __restrict__ cu_rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + warp;
  if (b >= B) return;
  // last (possibly padded) position, model.py:700; packed batches: the sample's last packed row (position T-1)
  const int row = cu_rows ? cu_rows[b + 1] - 1 : b * T + T - 1;
  const float mu = mean[row], rs = rstd[row];
  float acc[NUM_EMO];
#pragma unroll
  for (int j = 0; j < NUM_EMO; ++j) acc[j] = 0.f;
  for (int c = lane; c < H; c += 32) {
    const float h = (x_final[(int64_t)row * H + c] - mu) * rs * gamma[c] + beta[c];
    hlast[(int64_t)b * H + c] = h;
#pragma unroll
    for (int j = 0; j < NUM_EMO; ++j) acc[j] += h * w_emo[j * H + c];
  }
#pragma unroll
  for (int j = 0; j < NUM_EMO; ++j) acc[j] = warp_sum(acc[j]);
  if (lane == 0) {
    float mx = acc[0];
#pragma unroll
    for (int j = 1; j < NUM_EMO; ++j) mx = fmaxf(mx, acc[j]);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NUM_EMO; ++j) s += expf(acc[j] - mx);
    const float lse = mx + logf(s);
#pragma unroll
    for (int j = 0; j < NUM_EMO; ++j) logits_out[b * NUM_EMO + j] = acc[j];
    if (emo_labels) {
      const int64_t y = emo_labels[b];
      if (y < 0 || y >= NUM_EMO) { *err_flag = 1; return; }
#pragma unroll
      for (int j = 0; j < NUM_EMO; ++j)
        dlogits_out[b * NUM_EMO + j] = expf(acc[j] - lse) - (j == (int)y ? 1.f : 0.f);
      atomicAdd(sums + 2, lse - acc[y]);
      atomicAdd(sums + 3, 1.f);
    }
  }
}

// dW_emo[j,c] += s * sum_b dlog[b,j] * hlast[b,c];  dyf[(b,T-1), c] += s * sum_j dlog[b,j] W[j,c]
// grid (column blocks, B): one thread per (sample, column) - the first version looped over the batch inside one thread
// per column (3 CTAs, 32 dependent global round trips: 38 us on the critical path of the backward).
__global__ void __launch_bounds__(128)
emotion_bwd_kernel(const float* __restrict__ dlog, const float* __restrict__ hlast,
                   const float* __restrict__ w_emo, const float* __restrict__ scale_ptr, int B,
                   int T, int H, float* __restrict__ dw_emo, float* __restrict__ dyf,
                   const int* __restrict__ cu_rows) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (c >= H) return;
  const float s = *scale_ptr;
  const float h = hlast[(int64_t)b * H + c];
  float dh = 0.f;
#pragma unroll
  for (int j = 0; j < NUM_EMO; ++j) {
    const float d = dlog[b * NUM_EMO + j] * s;
    dh += d * w_emo[j * H + c];
    if (dw_emo) atomicAdd(dw_emo + j * H + c, d * h);
  }
  if (dyf) dyf[(int64_t)(cu_rows ? cu_rows[b + 1] - 1 : b * T + T - 1) * H + c] += dh;
}

// sums = [lm_loss_sum, lm_valid, emo_loss_sum, emo_count] (already globally reduced under DP)
// out  = [loss, lm_loss, emo_loss, lm_scale (= 1/lm_valid), emo_scale (= 1/emo_count)]
__global__ void loss_finalize_kernel(const float* __restrict__ sums, int has_lm, int has_emo,
                                     float* __restrict__ out) {
  const float lm = has_lm ? sums[0] / sums[1] : 0.f;   // mean over valid (NaN if none, as torch)
  const float em = has_emo ? sums[2] / sums[3] : 0.f;
  out[0] = lm + em;
  out[1] = lm;
  out[2] = em;
  out[3] = has_lm ? 1.f / sums[1] : 0.f;
  out[4] = has_emo ? 1.f / sums[3] : 0.f;
}

// dst[i] = a[i] * b[0]   (tiny helper: upstream grad * 1/count, stays on device)
__global__ void scalar_mul_kernel(const float* a, const float* b, float* dst) { dst[0] = a[0] * b[0]; }

// ------------------------------------------------------------------------------------------
// AdamW over a flat fp32 buffer; hyper = [lr, beta1, beta2, eps, weight_decay, bc1, bc2] on device
// ------------------------------------------------------------------------------------------
template <bool GBF16>   // gradient buffer dtype: fp32, or bf16 (data-parallel buckets reduced in bf16)
__global__ void __launch_bounds__(256)
adamw_flat_kernel(float* __restrict__ p, const void* __restrict__ g, float* __restrict__ m,
                  float* __restrict__ v, __nv_bfloat16* __restrict__ shadow, int64_t n4,
                  const float* __restrict__ hyper, const float* __restrict__ grad_scale) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4],
              bc1 = hyper[5], bc2 = hyper[6];
  const float gs = grad_scale ? *grad_scale : 1.f;
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.f - lr * wd;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg;
    if (GBF16) {
      const uint2 u = reinterpret_cast<const uint2*>(g)[i];
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
      gg = make_float4(a.x, a.y, b.x, b.y);
    } else {
      gg = reinterpret_cast<const float4*>(g)[i];
    }
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = ga[k] * gs;
      pa[k] *= decay;
      ma[k] = ma[k] + (gk - ma[k]) * (1.f - b1);
      va[k] = va[k] * b2 + (1.f - b2) * gk * gk;
      const float denom = sqrtf(va[k]) * inv_sqrt_bc2 + eps;
      pa[k] -= step_size * (ma[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (shadow)
      reinterpret_cast<uint2*>(shadow)[i] = make_uint2(pack_bf16x2(pp.x, pp.y), pack_bf16x2(pp.z, pp.w));
  }
}

// ------------------------------------------------------------------------------------------
// Label-sparse LM head.  model.py:698-708 scores EVERY position with the [H, V] head and then ignores the
// positions whose shifted label is -100 (the dialogue history: ~90 % of a MELD-shaped batch).  Loss and all
// gradients only depend on the scored rows, so the head runs on the compacted rows: plan (ordered stream
// compaction of the scored row indices + their targets, on the device: no host synchronisation, CUDA-graph
// safe), gather of the ln_f output rows, GEMM / CE / dgrad / wgrad with a run-time row count, scatter of
// d(ln_f output) back.  `.logits` of all positions are produced only if somebody reads them.
// ------------------------------------------------------------------------------------------
constexpr int PLAN_THREADS = 1024;
__global__ void __launch_bounds__(PLAN_THREADS)
lm_rows_plan_kernel(const int64_t* __restrict__ labels, int rows, int T, int* __restrict__ row_idx,
                    int64_t* __restrict__ labels_c, int* __restrict__ count_out, const PackView pk) {
  __shared__ int s_warp[PLAN_THREADS / 32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  const int cap = rows;
  if (pk.on()) rows = min(rows, *pk.n_rows);   // packed batch: row r is (sample row_b[r], position row_t[r])
  for (int r0 = 0; r0 < rows; r0 += PLAN_THREADS) {
    const int row = r0 + tid;
    int64_t tgt = -100;
    if (row < rows) {
      if (pk.on()) {
        const int t = pk.row_t[row];
        tgt = (t + 1 < T) ? labels[pk.row_b[row] * T + t + 1] : (int64_t)-100;
      } else {
        tgt = shifted_target(labels, row, T);
      }
    }
    const bool keep = tgt != -100;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < PLAN_THREADS / 32; ++w) {
      const int c = s_warp[w];
      before += w < warp ? c : 0;
      total += c;
    }
    if (keep) {
      const int dst = s_base + before + __popc(m & ((1u << lane) - 1u));
      row_idx[dst] = row;
      labels_c[dst] = tgt;
    }
    __syncthreads();
    if (tid == 0) s_base += total;
    __syncthreads();
  }
  const int n = s_base;
  // pad up to the next multiple of 128 rows: ignored targets, "no source row"
  for (int i = n + tid; i < ((n + 127) & ~127) && i < cap; i += PLAN_THREADS) { row_idx[i] = -1; labels_c[i] = -100; }
  if (tid == 0) *count_out = n;
}

// dst[i, :] = src[row_idx[i], :] for i < count; zero rows up to the next multiple of 128 (bf16, 16-byte vectors)
__global__ void __launch_bounds__(256)
gather_rows_dyn_kernel(const __nv_bfloat16* __restrict__ src, const int* __restrict__ row_idx,
                       const int* __restrict__ count, __nv_bfloat16* __restrict__ dst, int H, int cap) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = *count, n_pad = min((n + 127) & ~127, cap);
  for (int i = blockIdx.x * 8 + warp; i < n_pad; i += gridDim.x * 8) {
    uint4* d = reinterpret_cast<uint4*>(dst + (int64_t)i * H);
    if (i < n) {
      const uint4* s = reinterpret_cast<const uint4*>(src + (int64_t)row_idx[i] * H);
      for (int c = lane; c < H / 8; c += 32) d[c] = s[c];
    } else {
      for (int c = lane; c < H / 8; c += 32) d[c] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

// dst[row_idx[i], :] = src[i, :] for i < count (fp32; the scored rows are distinct, plain stores)
__global__ void __launch_bounds__(256)
scatter_rows_dyn_kernel(const float* __restrict__ src, const int* __restrict__ row_idx,
                        const int* __restrict__ count, float* __restrict__ dst, int H) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = *count;
  for (int i = blockIdx.x * 8 + warp; i < n; i += gridDim.x * 8) {
    const float4* s = reinterpret_cast<const float4*>(src + (int64_t)i * H);
    float4* d = reinterpret_cast<float4*>(dst + (int64_t)row_idx[i] * H);
    for (int c = lane; c < H / 4; c += 32) d[c] = s[c];
  }
}

// ------------------------------------------------------------------------------------------
// Packed variable-length batches: plan of the packed row layout from the per-sample token counts.
// Sample b contributes its lens[b] real positions and, when it is padded (lens[b] < T), ONE more row for position
// T-1: the emotion head reads the last position (model.py:700) whatever it holds, and under the right-padded
// attention mask that position attends to the sample's real tokens only - so its hidden state needs no other pad row.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
pack_plan_kernel(const int* __restrict__ lens, int B, int T, int* __restrict__ cu, int* __restrict__ row_b,
                 int* __restrict__ row_t, int* __restrict__ n_rows, int* __restrict__ kv_lens, int cap) {
  __shared__ int s_cu[1025];
  const int tid = threadIdx.x;
  if (tid == 0) {
    int acc = 0;
    for (int b = 0; b < B; ++b) {
      const int len = max(1, min(lens[b], T));
      s_cu[b] = acc;
      acc += len + (len < T ? 1 : 0);
    }
    s_cu[B] = acc;
    *n_rows = acc;
  }
  __syncthreads();
  for (int b = tid; b <= B; b += 1024) cu[b] = s_cu[b];
  for (int b = 0; b < B; ++b) {
    const int len = max(1, min(lens[b], T)), base = s_cu[b], n = s_cu[b + 1] - base;
    if (tid == 0) kv_lens[b] = len;
    for (int i = tid; i < n; i += 1024) { row_b[base + i] = b; row_t[base + i] = i < len ? i : T - 1; }
  }
  const int total = s_cu[B];
  for (int i = total + tid; i < ((total + 127) & ~127) && i < cap; i += 1024) { row_b[i] = 0; row_t[i] = 0; }
}

// rows [*count, roundup(*count, 128)) of a [cap, row_bytes] buffer := 0 (operands of run-time-K GEMMs)
__global__ void __launch_bounds__(256)
zero_rows_dyn_kernel(unsigned char* __restrict__ buf, int64_t row_bytes, const int* __restrict__ count, int cap) {
  const int n = *count;
  const int n_pad = min((n + 127) & ~127, cap);
  const int64_t total16 = (int64_t)(n_pad - n) * row_bytes / 16;
  uint4* base = reinterpret_cast<uint4*>(buf + (int64_t)n * row_bytes);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total16; i += (int64_t)gridDim.x * blockDim.x)
    base[i] = make_uint4(0u, 0u, 0u, 0u);
}

}  // namespace ergm

using namespace ergm;

extern "C" int ergm_pack_plan(const int* lens, int B, int T, int* cu_rows, int* row_b, int* row_t, int* n_rows,
                              int* kv_lens, int cap, void* stream) {
  if (!lens || !cu_rows || !row_b || !row_t || !n_rows || !kv_lens || B <= 0 || B > 1024 || T <= 0 || cap < B * T)
    return ERGM_ERR_ARG;
  pack_plan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(lens, B, T, cu_rows, row_b, row_t, n_rows, kv_lens, cap);
  return (int)cudaGetLastError();
}

extern "C" int ergm_zero_rows_dyn(void* buf, int64_t row_bytes, const int* count, int cap, void* stream) {
  if (!buf || !count || row_bytes <= 0 || row_bytes % 16 || cap <= 0 || (reinterpret_cast<uintptr_t>(buf) & 15))
    return ERGM_ERR_ARG;
  zero_rows_dyn_kernel<<<num_sms(), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned char*>(buf), row_bytes, count, cap);
  return (int)cudaGetLastError();
}

extern "C" int ergm_lm_rows_plan(const int64_t* labels, int rows, int T, int* row_idx, int64_t* labels_c,
                                 int* count, const ergm_pack* pack, void* stream) {
  if (!labels || !row_idx || !labels_c || !count || rows <= 0 || T <= 0) return ERGM_ERR_ARG;
  if (pack && (!pack->row_b || !pack->row_t || !pack->n_rows)) return ERGM_ERR_ARG;
  lm_rows_plan_kernel<<<1, PLAN_THREADS, 0, (cudaStream_t)stream>>>(labels, rows, T, row_idx, labels_c, count,
                                                                     ERGM_PACK_VIEW(pack));
  return (int)cudaGetLastError();
}

extern "C" int ergm_gather_rows_dyn(const void* src_bf16, const int* row_idx, const int* count, void* dst_bf16,
                                    int H, int cap, void* stream) {
  if (!src_bf16 || !row_idx || !count || !dst_bf16 || H <= 0 || H % 8 || cap <= 0) return ERGM_ERR_ARG;
  gather_rows_dyn_kernel<<<num_sms() * 2, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(src_bf16), row_idx, count, reinterpret_cast<__nv_bfloat16*>(dst_bf16), H, cap);
  return (int)cudaGetLastError();
}

extern "C" int ergm_scatter_rows_dyn(const float* src, const int* row_idx, const int* count, float* dst, int H,
                                     void* stream) {
  if (!src || !row_idx || !count || !dst || H <= 0 || H % 4) return ERGM_ERR_ARG;
  scatter_rows_dyn_kernel<<<num_sms() * 2, 256, 0, (cudaStream_t)stream>>>(src, row_idx, count, dst, H);
  return (int)cudaGetLastError();
}

extern "C" int ergm_ce_fwd(const void* logits, int logits_is_f32, int64_t ldl, const int64_t* labels,
                           int rows, int T, int V, float* lse, float* row_loss, float* sums,
                           int* err_flag, const void* hn_bf16, const void* w_bf16, int H,
                           const int* rows_dyn, void* stream) {
  if (!logits || !labels || !lse || !row_loss || !sums || !err_flag || rows <= 0 || T < 0 || V <= 0)
    return ERGM_ERR_ARG;
  if (ldl % 8 || (reinterpret_cast<uintptr_t>(logits) & 15)) return ERGM_ERR_ARG;
  if ((hn_bf16 != nullptr) != (w_bf16 != nullptr) || (hn_bf16 && H % 8)) return ERGM_ERR_ARG;
  const __nv_bfloat16* hn = reinterpret_cast<const __nv_bfloat16*>(hn_bf16);
  const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(w_bf16);
  if (logits_is_f32)
    ce_fwd_kernel<true><<<rows, CE_THREADS, 0, (cudaStream_t)stream>>>(logits, ldl, labels, T, V, lse, row_loss, sums, err_flag, hn, w, H, rows_dyn);
  else
    ce_fwd_kernel<false><<<rows, CE_THREADS, 0, (cudaStream_t)stream>>>(logits, ldl, labels, T, V, lse, row_loss, sums, err_flag, hn, w, H, rows_dyn);
  return (int)cudaGetLastError();
}

extern "C" int ergm_ce_bwd(const void* logits, int logits_is_f32, int64_t ldl, const int64_t* labels,
                           int rows, int T, int V, const float* lse, const float* scale_ptr,
                           void* dlogits_bf16, int64_t ldd, const int* rows_dyn, void* stream) {
  if (!logits || !labels || !lse || !scale_ptr || !dlogits_bf16 || rows <= 0 || T < 0) return ERGM_ERR_ARG;
  if (ldl % 8 || ldd % 8 || ldd < V) return ERGM_ERR_ARG;
  if (logits_is_f32)
    ce_bwd_kernel<true><<<rows, CE_THREADS, 0, (cudaStream_t)stream>>>(
        logits, ldl, labels, T, V, lse, scale_ptr, reinterpret_cast<__nv_bfloat16*>(dlogits_bf16), ldd, rows_dyn);
  else
    ce_bwd_kernel<false><<<rows, CE_THREADS, 0, (cudaStream_t)stream>>>(
        logits, ldl, labels, T, V, lse, scale_ptr, reinterpret_cast<__nv_bfloat16*>(dlogits_bf16), ldd, rows_dyn);
  return (int)cudaGetLastError();
}

extern "C" int ergm_emotion_head_fwd(const float* x_final, const float* mean, const float* rstd,
                                     const float* gamma, const float* beta, const float* w_emo,
                                     const int64_t* emotion_labels, int B, int T, int H,
                                     float* hlast, float* logits, float* dlogits, float* sums,
                                     int* err_flag, const int* cu_rows, void* stream) {
  if (!x_final || !mean || !rstd || !gamma || !beta || !w_emo || !hlast || !logits || B <= 0)
    return ERGM_ERR_ARG;
  if (emotion_labels && (!dlogits || !sums || !err_flag)) return ERGM_ERR_ARG;
  emotion_fwd_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
      x_final, mean, rstd, gamma, beta, w_emo, emotion_labels, B, T, H, hlast, logits, dlogits, sums, err_flag, cu_rows);
  return (int)cudaGetLastError();
}

extern "C" int ergm_emotion_head_bwd(const float* dlogits, const float* hlast, const float* w_emo,
                                     const float* scale_ptr, int B, int T, int H, float* dw_emo,
                                     float* dyf, const int* cu_rows, void* stream) {
  if (!dlogits || !hlast || !w_emo || !scale_ptr || B <= 0) return ERGM_ERR_ARG;
  emotion_bwd_kernel<<<dim3((H + 127) / 128, B), 128, 0, (cudaStream_t)stream>>>(dlogits, hlast, w_emo, scale_ptr, B, T, H, dw_emo, dyf, cu_rows);
  return (int)cudaGetLastError();
}

extern "C" int ergm_loss_finalize(const float* sums, int has_lm, int has_emotion, float* out,
                                  void* stream) {
  if (!sums || !out) return ERGM_ERR_ARG;
  loss_finalize_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sums, has_lm, has_emotion, out);
  return (int)cudaGetLastError();
}

extern "C" int ergm_scalar_mul(const float* a, const float* b, float* dst, void* stream) {
  if (!a || !b || !dst) return ERGM_ERR_ARG;
  scalar_mul_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(a, b, dst);
  return (int)cudaGetLastError();
}

extern "C" int ergm_adamw_flat(float* p, const void* g, int g_is_bf16, float* m, float* v, void* shadow_bf16,
                               int64_t n, const float* hyper, const float* grad_scale, void* stream) {
  if (!p || !g || !m || !v || !hyper || n <= 0 || n % 4) return ERGM_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(g) & (g_is_bf16 ? 7 : 15)) return ERGM_ERR_ARG;
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  if (g_is_bf16)
    adamw_flat_kernel<true><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
        p, g, m, v, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), n4, hyper, grad_scale);
  else
    adamw_flat_kernel<false><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
        p, g, m, v, reinterpret_cast<__nv_bfloat16*>(shadow_bf16), n4, hyper, grad_scale);
  return (int)cudaGetLastError();
}

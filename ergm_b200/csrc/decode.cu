// Decode-side kernels (HBM-bound): paged KV cache, one-query attention, on-device sampling.
//
// The reference has no real decode path in use: main.py:253-282 re-runs the FULL forward per
// generated token, and the model's own cache (model.py:228-236) grows by torch.cat — an
// O(context) copy per layer per token.  Here:
//   * K/V live in a paged pool  [page][k|v][head][16 tokens][64]  (bf16) addressed through a
//     per-sequence block table, so appending a token writes 2 x 128 B per head and nothing moves;
//   * ergm_attn_decode_paged appends the new token's K/V and attends over the whole context in
//     one kernel: one CTA per (head, sequence), 16-byte vector loads (a 128-byte K or V row is
//     read by 8 lanes), fp32 scores / softmax / accumulation;
//   * ergm_attn_decode_contig is the same kernel over a contiguous [B, Tk, ld] matrix: the
//     cross-attention K/V of the caption embeddings are projected ONCE per request and reused
//     by every step (the reference recomputes them each call, model.py:319-326);
//   * ergm_sample does greedy arg-max or top-k / temperature sampling per row without a host
//     round trip (main.py:271's .item() sync per token disappears), tracks finished sequences
//     and advances the per-sequence lengths.
#include "../../include/ergm_b200.h"
#include "common.cuh"

namespace ergm {

constexpr int PAGE = 16;           // tokens per KV page
constexpr int DEC_THREADS = 128;   // 16 groups of 8 lanes; a group owns one token at a time
constexpr int DEC_MAX_CTX = 2048;  // scores staged in smem

struct DecodeAttnParams {
  const __nv_bfloat16* q;    // [B, ld_q]: head h at q_col0 + 64h
  const __nv_bfloat16* kv_new;  // paged mode: new token's k at k_col0+64h, v at v_col0+64h of [B, ld_q]
  __nv_bfloat16* pool;       // paged: [pages][2][nh][PAGE][64]
  const int* block_table;    // [B, max_pages]
  const int* seq_lens;       // paged: tokens already cached (the new token goes to this slot)
  const __nv_bfloat16* kc;   // contiguous mode: [B*Tk, ld_k]
  const int* kv_lens;        // contiguous mode (nullable): valid keys per sequence
  __nv_bfloat16* out;        // [B, ld_out]
  int64_t ld_q, ld_k, ld_out;
  int q_col0, k_col0, v_col0;
  int nh, max_pages, Tk;
  float scale;
};

ERGM_DEVINL float dot8(const uint4 a, const uint4 b) {
  const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
  const float2 b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y), b2 = unpack_bf16x2(b.z), b3 = unpack_bf16x2(b.w);
  return a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
}

template <bool PAGED>
__global__ void __launch_bounds__(DEC_THREADS) attn_decode_kernel(const DecodeAttnParams p) {
  __shared__ float s_score[DEC_MAX_CTX];
  __shared__ float s_red[DEC_THREADS / 32];
  __shared__ float s_acc[DEC_THREADS / 8][64];
  const int h = blockIdx.x, b = blockIdx.y;
  const int grp = threadIdx.x >> 3, gl = threadIdx.x & 7;  // token group, lane inside the 128 B row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint4 qv = *reinterpret_cast<const uint4*>(p.q + (int64_t)b * p.ld_q + p.q_col0 + h * 64 + gl * 8);
  int ctx;
  if (PAGED) {
    const int pos = p.seq_lens[b];
    ctx = pos + 1;
    // append this head's new K / V row to the page pool (8 lanes x 16 B each)
    if (grp < 2) {
      const int page = p.block_table[b * p.max_pages + pos / PAGE];
      const int col = (grp == 0 ? p.k_col0 : p.v_col0) + h * 64 + gl * 8;
      const uint4 nv = *reinterpret_cast<const uint4*>(p.kv_new + (int64_t)b * p.ld_q + col);
      __nv_bfloat16* dst = p.pool + ((((int64_t)page * 2 + grp) * p.nh + h) * PAGE + pos % PAGE) * 64 + gl * 8;
      *reinterpret_cast<uint4*>(dst) = nv;
    }
    __syncthreads();  // the appended row is read back below by other groups
  } else {
    ctx = p.kv_lens ? min(p.Tk, p.kv_lens[b]) : p.Tk;
  }
  if (ctx > DEC_MAX_CTX) ctx = DEC_MAX_CTX;
  auto k_row = [&](int t) -> const uint4* {
    if (PAGED) {
      const int page = p.block_table[b * p.max_pages + t / PAGE];
      return reinterpret_cast<const uint4*>(p.pool + ((((int64_t)page * 2 + 0) * p.nh + h) * PAGE + t % PAGE) * 64) + gl;
    }
    return reinterpret_cast<const uint4*>(p.kc + ((int64_t)b * p.Tk + t) * p.ld_k + p.k_col0 + h * 64) + gl;
  };
  auto v_row = [&](int t) -> const uint4* {
    if (PAGED) {
      const int page = p.block_table[b * p.max_pages + t / PAGE];
      return reinterpret_cast<const uint4*>(p.pool + ((((int64_t)page * 2 + 1) * p.nh + h) * PAGE + t % PAGE) * 64) + gl;
    }
    return reinterpret_cast<const uint4*>(p.kc + ((int64_t)b * p.Tk + t) * p.ld_k + p.v_col0 + h * 64) + gl;
  };
  // phase 1: scores
  float mx = -INFINITY;
  // uniform trip count for the whole warp: the shuffles below are full-mask collectives
  for (int tb = 0; tb < ctx; tb += DEC_THREADS / 8) {
    const int t = tb + grp;
    const bool ok = t < ctx;
    float s = ok ? dot8(qv, *k_row(t)) : 0.f;
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s *= p.scale;
    if (ok) {
      if (gl == 0) s_score[t] = s;
      mx = fmaxf(mx, s);
    }
  }
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
  __syncthreads();
  // phase 2: exp + sum
  float sum = 0.f;
  for (int t = threadIdx.x; t < ctx; t += DEC_THREADS) {
    const float e = __expf(s_score[t] - mx);
    s_score[t] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  const float inv = 1.f / (s_red[0] + s_red[1] + s_red[2] + s_red[3]);
  // phase 3: weighted sum of V rows (each lane owns 8 of the 64 dims)
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int t = grp; t < ctx; t += DEC_THREADS / 8) {
    const float w = s_score[t];
    const uint4 v = *v_row(t);
    const float2 v0 = unpack_bf16x2(v.x), v1 = unpack_bf16x2(v.y), v2 = unpack_bf16x2(v.z), v3 = unpack_bf16x2(v.w);
    acc[0] += w * v0.x; acc[1] += w * v0.y; acc[2] += w * v1.x; acc[3] += w * v1.y;
    acc[4] += w * v2.x; acc[5] += w * v2.y; acc[6] += w * v3.x; acc[7] += w * v3.y;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s_acc[grp][gl * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 64) {
    float o = 0.f;
#pragma unroll
    for (int g = 0; g < DEC_THREADS / 8; ++g) o += s_acc[g][threadIdx.x];
    p.out[(int64_t)b * p.ld_out + h * 64 + threadIdx.x] = __float2bfloat16_rn(o * inv);
  }
}

// prompt K/V ([B*T, ld] projection output) -> pages.  grid (T, B), 128 threads = 16 x 8 lanes
__global__ void __launch_bounds__(128)
kv_to_pages_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ld, int k_col0, int v_col0,
                   __nv_bfloat16* __restrict__ pool, const int* __restrict__ block_table,
                   const int* __restrict__ lens, int T, int nh, int max_pages) {
  const int t = blockIdx.x, b = blockIdx.y;
  if (lens && t >= lens[b]) return;
  const int page = block_table[b * max_pages + t / PAGE];
  const int gl = threadIdx.x & 7;
  for (int item = threadIdx.x >> 3; item < 2 * nh; item += 16) {
    const int kv = item / nh, h = item % nh;
    const uint4 v = *reinterpret_cast<const uint4*>(qkv + ((int64_t)b * T + t) * ld + (kv ? v_col0 : k_col0) + h * 64 + gl * 8);
    __nv_bfloat16* dst = pool + ((((int64_t)page * 2 + kv) * nh + h) * PAGE + t % PAGE) * 64 + gl * 8;
    *reinterpret_cast<uint4*>(dst) = v;
  }
}

// ------------------------------------------------------------------------------------------
// sampling: one CTA per row of fp32 logits
// ------------------------------------------------------------------------------------------
constexpr int SMP_THREADS = 256;
constexpr int SMP_MAX_K = 64;

struct SampleParams {
  const float* logits;   // [B, ld]
  int64_t ld;
  int V;
  int top_k;             // 0 = greedy arg-max
  float inv_temperature;
  uint64_t seed;
  const int* step_ptr;   // device step counter: output column and RNG subsequence
  int64_t* out_ids;      // [B, out_ld]: out_ids[b, *step] = token
  int64_t out_ld;
  int64_t* next_ids;     // [B]: input ids of the next decode step
  int* finished;         // [B]
  int* seq_lens;         // [B]: advanced by one (nullable)
  int64_t eos_id;        // < 0: never finishes
};

// order-preserving float -> uint key (larger float = larger key)
ERGM_DEVINL uint32_t fkey(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(SMP_THREADS) sample_kernel(const SampleParams p) {
  __shared__ unsigned long long s_best[SMP_THREADS / 32];
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_prefix, s_need;
  __shared__ float s_cv[SMP_MAX_K];
  __shared__ int s_ci[SMP_MAX_K];
  __shared__ int s_cn;
  const int b = blockIdx.x;
  const float* row = p.logits + (int64_t)b * p.ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int token;
  if (p.top_k <= 1) {
    // arg-max, lowest index on ties (torch.argmax semantics): pack (key, ~index) and take the max
    unsigned long long best = 0ull;
    for (int i = threadIdx.x; i < p.V; i += SMP_THREADS) {
      const unsigned long long cand = ((unsigned long long)fkey(row[i]) << 32) | (uint32_t)(~(uint32_t)i);
      best = cand > best ? cand : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0) s_best[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < SMP_THREADS / 32; ++w) best = s_best[w] > best ? s_best[w] : best;
      s_ci[0] = (int)(~(uint32_t)(best & 0xffffffffu));
    }
    __syncthreads();
    token = s_ci[0];
  } else {
    // radix select (4 x 8 bits) of the k-th largest key, then gather the candidates
    const int k = min(p.top_k, SMP_MAX_K);
    if (threadIdx.x == 0) { s_prefix = 0u; s_need = (uint32_t)k; s_cn = 0; }
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      s_hist[threadIdx.x] = 0u;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const uint32_t mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
      for (int i = threadIdx.x; i < p.V; i += SMP_THREADS) {
        const uint32_t key = fkey(row[i]);
        if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        uint32_t need = s_need, cum = 0u;
        int bin = 255;
        for (; bin > 0; --bin) {
          if (cum + s_hist[bin] >= need) break;
          cum += s_hist[bin];
        }
        s_need = need - cum;
        s_prefix = prefix | ((uint32_t)bin << shift);
      }
      __syncthreads();
    }
    const uint32_t kth = s_prefix;  // key of the k-th largest logit
    for (int i = threadIdx.x; i < p.V; i += SMP_THREADS) {
      const float v = row[i];
      if (fkey(v) > kth) {
        const int slot = atomicAdd(&s_cn, 1);
        if (slot < SMP_MAX_K) { s_cv[slot] = v; s_ci[slot] = i; }
      }
    }
    __syncthreads();
    const int n_gt = min(s_cn, SMP_MAX_K);
    __syncthreads();
    // ties at the threshold: take the lowest indices until k candidates are collected
    if (threadIdx.x == 0) {
      int n = n_gt;
      for (int i = 0; i < p.V && n < k; ++i)
        if (fkey(row[i]) == kth) { s_cv[n] = row[i]; s_ci[n] = i; ++n; }
      s_cn = n;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const int n = s_cn;
      float mx = -INFINITY;
      for (int i = 0; i < n; ++i) mx = fmaxf(mx, s_cv[i]);
      float tot = 0.f;
      for (int i = 0; i < n; ++i) { s_cv[i] = __expf((s_cv[i] - mx) * p.inv_temperature); tot += s_cv[i]; }
      const int step = p.step_ptr ? *p.step_ptr : 0;
      Philox ph(p.seed, (uint64_t)step);
      const float u = u01(ph((uint64_t)b).x) * tot;
      // candidates are in arbitrary (atomic) order: walk them in index order for determinism
      float cum = 0.f;
      int pick = -1, last = -1;
      for (int it = 0; it < n; ++it) {
        int bi = -1;
        for (int i = 0; i < n; ++i)
          if (s_ci[i] > last && (bi < 0 || s_ci[i] < s_ci[bi])) bi = i;
        last = s_ci[bi];
        cum += s_cv[bi];
        pick = s_ci[bi];
        if (u < cum) break;
      }
      s_ci[0] = pick;
    }
    __syncthreads();
    token = s_ci[0];
  }
  if (threadIdx.x == 0) {
    int64_t tok = token;
    if (p.finished) {
      if (p.finished[b]) tok = p.eos_id;
      else if (p.eos_id >= 0 && tok == p.eos_id) p.finished[b] = 1;
    }
    const int step = p.step_ptr ? *p.step_ptr : 0;
    if (p.out_ids) p.out_ids[(int64_t)b * p.out_ld + step] = tok;
    if (p.next_ids) p.next_ids[b] = tok;
    if (p.seq_lens) p.seq_lens[b] += 1;
  }
}

__global__ void int_add_kernel(int* p, int inc) { *p += inc; }

}  // namespace ergm

using namespace ergm;

extern "C" int ergm_attn_decode_paged(const void* qkv, int64_t ld_q, int q_col0, int k_col0,
                                      int v_col0, void* pool, const int* block_table,
                                      const int* seq_lens, int max_pages, void* out, int64_t ld_out,
                                      int B, int nh, int head_dim, void* stream) {
  if (!qkv || !pool || !block_table || !seq_lens || !out || B <= 0 || nh <= 0) return ERGM_ERR_ARG;
  if (head_dim != 64) return ERGM_ERR_UNSUPPORTED;
  if (ld_q % 8 || q_col0 % 8 || k_col0 % 8 || v_col0 % 8) return ERGM_ERR_ARG;
  DecodeAttnParams p{};
  p.q = p.kv_new = reinterpret_cast<const __nv_bfloat16*>(qkv);
  p.pool = reinterpret_cast<__nv_bfloat16*>(pool);
  p.block_table = block_table; p.seq_lens = seq_lens;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ld_q = ld_q; p.ld_out = ld_out;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.nh = nh; p.max_pages = max_pages;
  p.scale = 1.0f / sqrtf((float)head_dim);
  attn_decode_kernel<true><<<dim3(nh, B), DEC_THREADS, 0, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}

extern "C" int ergm_attn_decode_contig(const void* q, int64_t ld_q, int q_col0, const void* kv,
                                       int64_t ld_k, int k_col0, int v_col0, const int* kv_lens,
                                       void* out, int64_t ld_out, int B, int nh, int Tk,
                                       int head_dim, void* stream) {
  if (!q || !kv || !out || B <= 0 || nh <= 0 || Tk <= 0) return ERGM_ERR_ARG;
  if (head_dim != 64) return ERGM_ERR_UNSUPPORTED;
  if (Tk > DEC_MAX_CTX) return ERGM_ERR_UNSUPPORTED;
  if (ld_q % 8 || ld_k % 8 || q_col0 % 8 || k_col0 % 8 || v_col0 % 8) return ERGM_ERR_ARG;
  DecodeAttnParams p{};
  p.q = reinterpret_cast<const __nv_bfloat16*>(q);
  p.kc = reinterpret_cast<const __nv_bfloat16*>(kv);
  p.kv_lens = kv_lens;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ld_q = ld_q; p.ld_k = ld_k; p.ld_out = ld_out;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.nh = nh; p.Tk = Tk;
  p.scale = 1.0f / sqrtf((float)head_dim);
  attn_decode_kernel<false><<<dim3(nh, B), DEC_THREADS, 0, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}

extern "C" int ergm_kv_to_pages(const void* kv, int64_t ld, int k_col0, int v_col0, void* pool,
                                const int* block_table, const int* lens, int max_pages, int B,
                                int T, int nh, void* stream) {
  if (!kv || !pool || !block_table || B <= 0 || T <= 0 || nh <= 0) return ERGM_ERR_ARG;
  if (ld % 8 || k_col0 % 8 || v_col0 % 8) return ERGM_ERR_ARG;
  kv_to_pages_kernel<<<dim3(T, B), 128, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(kv), ld, k_col0, v_col0, reinterpret_cast<__nv_bfloat16*>(pool),
      block_table, lens, T, nh, max_pages);
  return (int)cudaGetLastError();
}

extern "C" int ergm_sample(const float* logits, int64_t ld, int B, int V, int top_k,
                           float temperature, uint64_t seed, const int* step_ptr, int64_t* out_ids,
                           int64_t out_ld, int64_t* next_ids, int* finished, int* seq_lens,
                           int64_t eos_id, void* stream) {
  if (!logits || B <= 0 || V <= 0 || top_k < 0 || top_k > SMP_MAX_K) return ERGM_ERR_ARG;
  if (top_k > 1 && !(temperature > 0.f)) return ERGM_ERR_ARG;
  SampleParams p{logits, ld, V, top_k, top_k > 1 ? 1.f / temperature : 1.f, seed, step_ptr, out_ids, out_ld,
                 next_ids, finished, seq_lens, eos_id};
  sample_kernel<<<B, SMP_THREADS, 0, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}

extern "C" int ergm_int_add(int* dev_ptr, int inc, void* stream) {
  if (!dev_ptr) return ERGM_ERR_ARG;
  int_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dev_ptr, inc);
  return (int)cudaGetLastError();
}

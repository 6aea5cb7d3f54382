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
#include <cstdlib>

namespace ergm {

constexpr int PAGE = 16;           // tokens per KV page
constexpr int DEC_THREADS = 128;   // 16 groups of 8 lanes; a group owns one token at a time

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

// One CTA per (head, sequence).  16 groups of 8 lanes; a group owns tokens grp, grp+16, ... and keeps
// its OWN online-softmax state (running max, sum, 8 output dims per lane), so the token loop has no
// block-wide synchronisation and every thread keeps 16 independent 16-byte K/V loads in flight
// (8 tokens x {K, V}); the 16 partial states are merged once at the end (flash-decoding inside a CTA).
// The cached K/V of the context do not depend on the kernel that produced q: their first 128 tokens
// are fetched BEFORE the programmatic-dependency wait, i.e. while the QKV projection is still running.
constexpr int DEC_GROUPS = DEC_THREADS / 8;
constexpr int DEC_BT_SMEM = 256;   // block-table entries staged in smem (4096 tokens)

template <bool PAGED, int DEC_UNROLL>
__global__ void __launch_bounds__(DEC_THREADS) attn_decode_kernel(const DecodeAttnParams p) {
  __shared__ float s_m[DEC_GROUPS], s_l[DEC_GROUPS];
  __shared__ float s_acc[DEC_GROUPS][64];
  __shared__ int s_bt[DEC_BT_SMEM];
  const int h = blockIdx.x, b = blockIdx.y;
  const int grp = threadIdx.x >> 3, gl = threadIdx.x & 7;  // token group, lane inside the 128 B row
  pdl_launch_dependents();
  int n_old, pos = 0;  // cached tokens (the new one of paged mode is handled from registers)
  if (PAGED) {
    pos = p.seq_lens[b];
    n_old = pos;
    const int npages = min(pos / PAGE + 1, DEC_BT_SMEM);
    for (int i = threadIdx.x; i < npages; i += DEC_THREADS) s_bt[i] = p.block_table[b * p.max_pages + i];
    __syncthreads();
  } else {
    n_old = p.kv_lens ? min(p.Tk, p.kv_lens[b]) : p.Tk;
  }
  auto kv_row = [&](int t, int which) -> const uint4* {
    if (PAGED) {
      const int pi = t / PAGE;
      const int page = pi < DEC_BT_SMEM ? s_bt[pi] : p.block_table[b * p.max_pages + pi];
      return reinterpret_cast<const uint4*>(p.pool + ((((int64_t)page * 2 + which) * p.nh + h) * PAGE + t % PAGE) * 64) + gl;
    }
    return reinterpret_cast<const uint4*>(p.kc + ((int64_t)b * p.Tk + t) * p.ld_k + (which ? p.v_col0 : p.k_col0) + h * 64) + gl;
  };
  uint4 kk[DEC_UNROLL], vv[DEC_UNROLL];
  auto load_chunk = [&](int tb) {
#pragma unroll
    for (int u = 0; u < DEC_UNROLL; ++u) {
      const int t = tb + u * DEC_GROUPS + grp;
      const bool ok = t < n_old;
      kk[u] = ok ? __ldg(kv_row(t, 0)) : make_uint4(0u, 0u, 0u, 0u);
      vv[u] = ok ? __ldg(kv_row(t, 1)) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  load_chunk(0);
  pdl_wait();  // q (and the new token's K / V) come from the projection kernel just before this one
  const uint4 qv = *reinterpret_cast<const uint4*>(p.q + (int64_t)b * p.ld_q + p.q_col0 + h * 64 + gl * 8);
  float m_run = -INFINITY, l_run = 0.f;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  // uniform trip count for the whole warp: the shuffles below are full-mask collectives
  for (int tb = 0; tb < n_old; tb += DEC_GROUPS * DEC_UNROLL) {
    if (tb > 0) load_chunk(tb);
    float sc[DEC_UNROLL];
    float m_new = m_run;
#pragma unroll
    for (int u = 0; u < DEC_UNROLL; ++u) {
      float s = dot8(qv, kk[u]);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      sc[u] = (tb + u * DEC_GROUPS + grp < n_old) ? s * p.scale : -INFINITY;
      m_new = fmaxf(m_new, sc[u]);
    }
    if (m_new > -INFINITY) {
      const float corr = __expf(m_run - m_new);  // m_run = -inf -> 0
      l_run *= corr;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] *= corr;
#pragma unroll
      for (int u = 0; u < DEC_UNROLL; ++u) {
        const float w = __expf(sc[u] - m_new);  // masked tokens: exp(-inf) = 0
        l_run += w;
        const float2 v0 = unpack_bf16x2(vv[u].x), v1 = unpack_bf16x2(vv[u].y), v2 = unpack_bf16x2(vv[u].z),
                     v3 = unpack_bf16x2(vv[u].w);
        acc[0] += w * v0.x; acc[1] += w * v0.y; acc[2] += w * v1.x; acc[3] += w * v1.y;
        acc[4] += w * v2.x; acc[5] += w * v2.y; acc[6] += w * v3.x; acc[7] += w * v3.y;
      }
      m_run = m_new;
    }
  }
  if (PAGED && threadIdx.x < 32) {
    // the new token: straight from the projection output; warp 0 computes its score (all four groups,
    // redundantly, so the shuffles stay warp-uniform), group 0 folds it into its state, groups 0 / 1
    // append K / V to the page pool (8 lanes x 16 B each)
    const uint4 kn = *reinterpret_cast<const uint4*>(p.kv_new + (int64_t)b * p.ld_q + p.k_col0 + h * 64 + gl * 8);
    const uint4 vn = *reinterpret_cast<const uint4*>(p.kv_new + (int64_t)b * p.ld_q + p.v_col0 + h * 64 + gl * 8);
    float s = dot8(qv, kn);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s *= p.scale;
    if (grp == 0) {
      const float m_new = fmaxf(m_run, s);
      const float corr = __expf(m_run - m_new), w = __expf(s - m_new);
      l_run = l_run * corr + w;
      const float2 v0 = unpack_bf16x2(vn.x), v1 = unpack_bf16x2(vn.y), v2 = unpack_bf16x2(vn.z), v3 = unpack_bf16x2(vn.w);
      acc[0] = acc[0] * corr + w * v0.x; acc[1] = acc[1] * corr + w * v0.y;
      acc[2] = acc[2] * corr + w * v1.x; acc[3] = acc[3] * corr + w * v1.y;
      acc[4] = acc[4] * corr + w * v2.x; acc[5] = acc[5] * corr + w * v2.y;
      acc[6] = acc[6] * corr + w * v3.x; acc[7] = acc[7] * corr + w * v3.y;
      m_run = m_new;
    }
    if (grp < 2) {
      const int pi = pos / PAGE;
      const int page = pi < DEC_BT_SMEM ? s_bt[pi] : p.block_table[b * p.max_pages + pi];
      __nv_bfloat16* dst = p.pool + ((((int64_t)page * 2 + grp) * p.nh + h) * PAGE + pos % PAGE) * 64 + gl * 8;
      *reinterpret_cast<uint4*>(dst) = grp == 0 ? kn : vn;
    }
  }
  if (gl == 0) { s_m[grp] = m_run; s_l[grp] = l_run; }
#pragma unroll
  for (int i = 0; i < 8; ++i) s_acc[grp][gl * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 64) {
    float M = -INFINITY;
#pragma unroll
    for (int g = 0; g < DEC_GROUPS; ++g) M = fmaxf(M, s_m[g]);
    float Lsum = 0.f, o = 0.f;
#pragma unroll
    for (int g = 0; g < DEC_GROUPS; ++g) {
      const float w = s_m[g] > -INFINITY ? __expf(s_m[g] - M) : 0.f;
      Lsum += s_l[g] * w;
      o += s_acc[g][threadIdx.x] * w;
    }
    p.out[(int64_t)b * p.ld_out + h * 64 + threadIdx.x] = __float2bfloat16_rn(Lsum > 0.f ? o / Lsum : 0.f);
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
  float top_p;           // < 1: nucleus sampling (main.py:253-282)
  float inv_temperature;
  uint64_t seed;
  const int* step_ptr;   // device step counter: output column and RNG subsequence
  int64_t* out_ids;      // [B, out_ld]: out_ids[b, *step] = token
  int64_t out_ld;
  int64_t* next_ids;     // [B]: input ids of the next decode step
  int* finished;         // [B]
  int* seq_lens;         // [B]: advanced by one (nullable)
  int64_t eos_id;        // < 0: never finishes
  int* step_inc;         // nullable: incremented by one after EVERY row has been sampled (last block)
  unsigned int* done_ctr; // device counter used to find the last block (self-resetting)
};

// order-preserving float -> uint key (larger float = larger key)
ERGM_DEVINL uint32_t fkey(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ unsigned int g_sample_done_ctr = 0u;
ERGM_DEVINL void sample_commit(const SampleParams& p, int b, int token, int step_now);  // last-block detection of sample_kernel (one stream per device)

__global__ void __launch_bounds__(SMP_THREADS) sample_kernel(const SampleParams p) {
  __shared__ unsigned long long s_best[SMP_THREADS / 32];
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_prefix, s_need;
  __shared__ float s_cv[SMP_MAX_K];
  __shared__ int s_ci[SMP_MAX_K];
  __shared__ int s_cn;
  __shared__ int s_ti[SMP_MAX_K];   // indices of logits EQUAL to the k-th largest (ties at the threshold)
  __shared__ int s_tn;
  const int b = blockIdx.x;
  const float* row = p.logits + (int64_t)b * p.ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int token;
  pdl_launch_dependents();
  pdl_wait();
  const int step_now = p.step_ptr ? *p.step_ptr : 0;  // read before any block can bump it
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
    if (threadIdx.x == 0) { s_prefix = 0u; s_need = (uint32_t)k; s_cn = 0; s_tn = 0; }
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
    // one parallel pass collects the logits above the threshold AND the indices of those equal to it (the first
    // version found the ties with a serial scan of the row by one thread: 3.5 ms per step at V = 50260)
    for (int i = threadIdx.x; i < p.V; i += SMP_THREADS) {
      const float v = row[i];
      const uint32_t key = fkey(v);
      if (key > kth) {
        const int slot = atomicAdd(&s_cn, 1);
        if (slot < SMP_MAX_K) { s_cv[slot] = v; s_ci[slot] = i; }
      } else if (key == kth) {
        const int slot = atomicAdd(&s_tn, 1);
        if (slot < SMP_MAX_K) s_ti[slot] = i;
      }
    }
    __syncthreads();
    const int n_gt = min(s_cn, SMP_MAX_K);
    const int n_tie = s_tn;
    __syncthreads();
    // ties at the threshold: take the lowest indices until k candidates are collected
    if (threadIdx.x == 0) {
      int n = n_gt;
      if (n_tie <= SMP_MAX_K) {
        int last = -1;
        while (n < k) {                       // at most k - n_gt (normally 1) selections out of <= 64 entries
          int best = -1;
          for (int j = 0; j < n_tie; ++j)
            if (s_ti[j] > last && (best < 0 || s_ti[j] < best)) best = s_ti[j];
          if (best < 0) break;
          s_cv[n] = row[best]; s_ci[n] = best; ++n;
          last = best;
        }
      } else {                                // > 64 identical logits at the threshold (degenerate rows): serial scan
        for (int i = 0; i < p.V && n < k; ++i)
          if (fkey(row[i]) == kth) { s_cv[n] = row[i]; s_ci[n] = i; ++n; }
      }
      s_cn = n;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const int n = s_cn;
      float mx = -INFINITY;
      for (int i = 0; i < n; ++i) mx = fmaxf(mx, s_cv[i]);
      float tot = 0.f;
      for (int i = 0; i < n; ++i) { s_cv[i] = __expf((s_cv[i] - mx) * p.inv_temperature); tot += s_cv[i]; }
      Philox ph(p.seed, (uint64_t)step_now);
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
  if (threadIdx.x == 0) sample_commit(p, b, token, step_now);
}

// thread 0 of the row's block: finished / eos bookkeeping, output column, next input id, lengths,
// and (last block to get here) the step counter
ERGM_DEVINL void sample_commit(const SampleParams& p, int b, int token, int step_now) {
  int64_t tok = token;
  if (p.finished) {
    if (p.finished[b]) tok = p.eos_id;
    else if (p.eos_id >= 0 && tok == p.eos_id) p.finished[b] = 1;
  }
  if (p.out_ids) p.out_ids[(int64_t)b * p.out_ld + step_now] = tok;
  if (p.next_ids) p.next_ids[b] = tok;
  if (p.seq_lens) p.seq_lens[b] += 1;
  if (p.step_inc) {
    // every block read the step at its start; the last one to finish advances it
    __threadfence();
    if (atomicAdd(p.done_ctr, 1u) == gridDim.x - 1) {
      *p.done_ctr = 0u;
      *p.step_inc += 1;
    }
  }
}

// greedy arg-max (torch.argmax semantics: lowest index on ties): 1024 threads per row, 16-byte loads
constexpr int AMX_THREADS = 1024;
__global__ void __launch_bounds__(AMX_THREADS) argmax_kernel(const SampleParams p) {
  __shared__ unsigned long long s_best[AMX_THREADS / 32];
  const int b = blockIdx.x;
  const float* row = p.logits + (int64_t)b * p.ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  const int step_now = p.step_ptr ? *p.step_ptr : 0;
  unsigned long long best = 0ull;
  auto consider = [&](float v, int i) {
    const unsigned long long cand = ((unsigned long long)fkey(v) << 32) | (uint32_t)(~(uint32_t)i);
    best = cand > best ? cand : best;
  };
  int done = 0;
  if ((reinterpret_cast<uintptr_t>(row) & 15) == 0) {
    const int n4 = p.V >> 2;
    const float4* row4 = reinterpret_cast<const float4*>(row);
    int i = threadIdx.x;
    for (; i + 3 * AMX_THREADS < n4; i += 4 * AMX_THREADS) {
      const float4 v0 = row4[i], v1 = row4[i + AMX_THREADS], v2 = row4[i + 2 * AMX_THREADS], v3 = row4[i + 3 * AMX_THREADS];
      consider(v0.x, 4 * i); consider(v0.y, 4 * i + 1); consider(v0.z, 4 * i + 2); consider(v0.w, 4 * i + 3);
      const int i1 = i + AMX_THREADS, i2 = i + 2 * AMX_THREADS, i3 = i + 3 * AMX_THREADS;
      consider(v1.x, 4 * i1); consider(v1.y, 4 * i1 + 1); consider(v1.z, 4 * i1 + 2); consider(v1.w, 4 * i1 + 3);
      consider(v2.x, 4 * i2); consider(v2.y, 4 * i2 + 1); consider(v2.z, 4 * i2 + 2); consider(v2.w, 4 * i2 + 3);
      consider(v3.x, 4 * i3); consider(v3.y, 4 * i3 + 1); consider(v3.z, 4 * i3 + 2); consider(v3.w, 4 * i3 + 3);
    }
    for (; i < n4; i += AMX_THREADS) {
      const float4 v = row4[i];
      consider(v.x, 4 * i); consider(v.y, 4 * i + 1); consider(v.z, 4 * i + 2); consider(v.w, 4 * i + 3);
    }
    done = n4 << 2;
  }
  for (int i = done + threadIdx.x; i < p.V; i += AMX_THREADS) consider(row[i], i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if (lane == 0) s_best[warp] = best;
  __syncthreads();
  if (warp == 0) {
    best = s_best[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0) sample_commit(p, b, (int)(~(uint32_t)(best & 0xffffffffu)), step_now);
  }
}


// ------------------------------------------------------------------------------------------
// Nucleus (top-p) sampling, main.py:258-270, one 1024-thread CTA per row, no host round trip:
//   probs = softmax(logits); sort descending; remove[i] = cumsum[i-1] > top_p  (the mask is shifted
//   right by one, :263-265, so the token that crosses top_p is KEPT); renormalise; multinomial.
// Without a sort: e_i = exp(l_i - max) is cached in shared memory (V * 4 bytes), G(k) = sum of the
// e_i whose bit pattern exceeds k is non-increasing in k, and the kept set is {i : G(bits(e_i)) <=
// top_p * Z} = {bits(e_i) >= k_min}: ~30 bisection steps over the float bit patterns, each one pass
// over shared memory + a block reduction.  Tokens tied at k_min are kept lowest-index first, as many as
// the shifted rule admits.  The draw walks the kept tokens in INDEX order (deterministic for a given
// (seed, step, row), independent of any sort order).
// ------------------------------------------------------------------------------------------
constexpr int NUC_THREADS = 1024;

ERGM_DEVINL float block_sum_1024(float v, float* s_red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = s_red[threadIdx.x & 31];
  t = warp_sum(t);
  return t;
}

__global__ void __launch_bounds__(NUC_THREADS) nucleus_kernel(const SampleParams p) {
  extern __shared__ float s_e[];          // [V]
  __shared__ float s_red[32];
  __shared__ float s_scan_f[NUC_THREADS / 32];
  __shared__ int s_scan_i[NUC_THREADS / 32];
  __shared__ int s_pick;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* row = p.logits + (int64_t)b * p.ld;
  pdl_launch_dependents();
  pdl_wait();
  const int step_now = p.step_ptr ? *p.step_ptr : 0;
  // softmax numerator in smem
  float mx = -INFINITY;
  for (int i = tid; i < p.V; i += NUC_THREADS) mx = fmaxf(mx, row[i]);
  mx = warp_max(mx);
  __syncthreads();
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = warp_max(s_red[lane]);
  float z = 0.f;
  for (int i = tid; i < p.V; i += NUC_THREADS) {
    const float e = __expf((row[i] - mx) * p.inv_temperature);
    s_e[i] = e;
    z += e;
  }
  const float Z = block_sum_1024(z, s_red);
  const float target = p.top_p * Z;
  // bisection: smallest k in [0, bits(1.0f)] with G(k) <= target
  uint32_t lo = 0u, hi = 0x3f800000u;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    float g = 0.f;
    for (int i = tid; i < p.V; i += NUC_THREADS) {
      const float e = s_e[i];
      g += (__float_as_uint(e) > mid) ? e : 0.f;
    }
    g = block_sum_1024(g, s_red);
    if (g <= target) hi = mid; else lo = mid + 1u;
  }
  const uint32_t kmin = lo;
  float mass_gt = 0.f;
  int cnt_tie = 0;
  for (int i = tid; i < p.V; i += NUC_THREADS) {
    const float e = s_e[i];
    const uint32_t u = __float_as_uint(e);
    mass_gt += u > kmin ? e : 0.f;
    cnt_tie += u == kmin ? 1 : 0;
  }
  mass_gt = block_sum_1024(mass_gt, s_red);
  const int n_tie = (int)(block_sum_1024((float)cnt_tie, s_red) + 0.5f);
  const float pstar = __uint_as_float(kmin);
  int tie_keep = n_tie;
  if (pstar > 0.f) tie_keep = min(n_tie, (int)floorf(fmaxf(target - mass_gt, 0.f) / pstar) + 1);
  const float total = mass_gt + (float)tie_keep * pstar;
  Philox ph(p.seed, (uint64_t)step_now);
  const float u = u01(ph((uint64_t)b).x) * total;
  // index-order walk: thread t owns the contiguous indices [t * npt, (t + 1) * npt)
  const int npt = (p.V + NUC_THREADS - 1) / NUC_THREADS;
  const int i0 = min(tid * npt, p.V), i1 = min(i0 + npt, p.V);
  // (a) exclusive prefix of the tie counts
  int my_ties = 0;
  for (int i = i0; i < i1; ++i) my_ties += __float_as_uint(s_e[i]) == kmin ? 1 : 0;
  int incl = my_ties;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  __syncthreads();
  if (lane == 31) s_scan_i[warp] = incl;
  __syncthreads();
  int wbase = 0;
  for (int w = 0; w < warp; ++w) wbase += s_scan_i[w];
  int tie_rank = wbase + incl - my_ties;
  // (b) kept mass of my chunk, exclusive prefix over threads
  float my_mass = 0.f;
  {
    int r = tie_rank;
    for (int i = i0; i < i1; ++i) {
      const float e = s_e[i];
      const uint32_t k = __float_as_uint(e);
      if (k > kmin) my_mass += e;
      else if (k == kmin) { if (r < tie_keep) my_mass += e; ++r; }
    }
  }
  float fincl = my_mass;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, fincl, o);
    if (lane >= o) fincl += n;
  }
  if (tid == 0) s_pick = -1;
  __syncthreads();
  if (lane == 31) s_scan_f[warp] = fincl;
  __syncthreads();
  float fbase = 0.f;
  for (int w = 0; w < warp; ++w) fbase += s_scan_f[w];
  const float before = fbase + fincl - my_mass;
  // (c) the thread whose kept-mass interval contains u walks its chunk; the last kept token overall is
  //     the fallback if rounding leaves u just past the end
  if (my_mass > 0.f && u >= before && u < before + my_mass) {
    float c = before;
    int r = tie_rank, pick = -1;
    for (int i = i0; i < i1; ++i) {
      const float e = s_e[i];
      const uint32_t k = __float_as_uint(e);
      bool kept = k > kmin;
      if (k == kmin) { kept = r < tie_keep; ++r; }
      if (kept) { pick = i; c += e; if (u < c) break; }
    }
    s_pick = pick;
  }
  __syncthreads();
  if (s_pick < 0) {  // u == total after rounding: take the highest-index kept token
    int last = -1;
    int r = tie_rank;
    for (int i = i0; i < i1; ++i) {
      const uint32_t k = __float_as_uint(s_e[i]);
      bool kept = k > kmin;
      if (k == kmin) { kept = r < tie_keep; ++r; }
      if (kept) last = i;
    }
    if (last >= 0) atomicMax(&s_pick, last);
  }
  __syncthreads();
  if (tid == 0) sample_commit(p, b, max(s_pick, 0), step_now);
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
  static int unroll = 0;
  if (!unroll) { const char* e = getenv("ERGM_DEC_UNROLL"); unroll = e ? atoi(e) : 8; }
  if (unroll == 4) return (int)launch_pdl(attn_decode_kernel<true, 4>, dim3(nh, B), dim3(DEC_THREADS), 0, (cudaStream_t)stream, 1, p);
  return (int)launch_pdl(attn_decode_kernel<true, 8>, dim3(nh, B), dim3(DEC_THREADS), 0, (cudaStream_t)stream, 1, p);
}

extern "C" int ergm_attn_decode_contig(const void* q, int64_t ld_q, int q_col0, const void* kv,
                                       int64_t ld_k, int k_col0, int v_col0, const int* kv_lens,
                                       void* out, int64_t ld_out, int B, int nh, int Tk,
                                       int head_dim, void* stream) {
  if (!q || !kv || !out || B <= 0 || nh <= 0 || Tk <= 0) return ERGM_ERR_ARG;
  if (head_dim != 64) return ERGM_ERR_UNSUPPORTED;
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
  return (int)launch_pdl(attn_decode_kernel<false, 8>, dim3(nh, B), dim3(DEC_THREADS), 0, (cudaStream_t)stream, 1, p);
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

extern "C" int ergm_sample(const float* logits, int64_t ld, int B, int V, int top_k, float top_p,
                           float temperature, uint64_t seed, int* step_ptr, int advance_step, int64_t* out_ids,
                           int64_t out_ld, int64_t* next_ids, int* finished, int* seq_lens,
                           int64_t eos_id, void* stream) {
  if (!logits || B <= 0 || V <= 0 || top_k < -1 || top_k > SMP_MAX_K) return ERGM_ERR_ARG;
  // top_k == -1 (ERGM_SAMPLE_ALL): multinomial over the whole distribution = the nucleus kernel at top_p = 1
  const bool all = top_k == -1;
  if (all) { top_k = 0; if (top_p > 1.0f) top_p = 1.0f; }
  const bool nucleus = top_p < 1.0f || all;
  if (nucleus && (!(top_p > 0.f) || top_k > 1)) return ERGM_ERR_ARG;  // top-k and top-p are alternatives
  if ((top_k > 1 || nucleus) && !(temperature > 0.f)) return ERGM_ERR_ARG;
  if (advance_step && !step_ptr) return ERGM_ERR_ARG;
  unsigned int* ctr = nullptr;
  if (advance_step) ERGM_CUDA_TRY(cudaGetSymbolAddress(reinterpret_cast<void**>(&ctr), g_sample_done_ctr));
  SampleParams p{logits, ld, V, top_k, top_p, (top_k > 1 || nucleus) ? 1.f / temperature : 1.f, seed, step_ptr, out_ids,
                 out_ld, next_ids, finished, seq_lens, eos_id, advance_step ? step_ptr : nullptr, ctr};
  cudaStream_t st = (cudaStream_t)stream;
  if (nucleus) {
    const size_t smem = (size_t)V * 4;
    if (smem > 220 * 1024) return ERGM_ERR_UNSUPPORTED;
    ERGM_SET_SMEM_ATTR(nucleus_kernel, 220 * 1024);
    return (int)launch_pdl(nucleus_kernel, dim3((unsigned)B), dim3(NUC_THREADS), smem, st, 1, p);
  }
  if (top_k <= 1) return (int)launch_pdl(argmax_kernel, dim3((unsigned)B), dim3(AMX_THREADS), 0, st, 1, p);
  return (int)launch_pdl(sample_kernel, dim3((unsigned)B), dim3(SMP_THREADS), 0, st, 1, p);
}

extern "C" int ergm_int_add(int* dev_ptr, int inc, void* stream) {
  if (!dev_ptr) return ERGM_ERR_ARG;
  int_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dev_ptr, inc);
  return (int)cudaGetLastError();
}

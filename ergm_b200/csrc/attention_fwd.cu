// ergm_attn_fwd — fused flash-style attention forward on tcgen05 / TMEM for head_dim 64.
//
// Replaces GPT2Attention._attn (/root/reference/src/model.py:119-148): QK^T, /sqrt(hd),
// causal where-mask (self) or none (cross), softmax in fp32, attn_dropout, PV, and the
// _split_heads / _merge_heads permutes (:190-198) — Q/K/V are read straight out of the
// [rows, ld] projection outputs with 3-D TMA maps (col, row-in-sequence, batch) and the
// context is written merged as [B*Tq, nh*64].
//
// One CTA = one (batch, head, 128-query block).  Warp roles: 0 = TMA, 1 = MMA issuer,
// 2 = TMEM allocator, 4..7 = softmax (thread t owns query row t: tcgen05.ld 32x32b gives a
// thread its whole row, so max / sum need no shuffles).  Per 128-key block:
//   S = Q K^T            (SS MMA, M128 N128 K64)  -> TMEM cols [0,128)
//   P = exp2(S*c - m*c)  (softmax warps)          -> smem, K-major SW128 (hand swizzled)
//   O_j = P V            (SS MMA, M128 N64 K128)  -> TMEM cols [128,192)
//   O = O*alpha + O_j    (registers)
// 256 TMEM columns and ~113 KB smem per CTA -> two CTAs per SM overlap each other's phases.
#include "../../include/ergm_b200.h"
#include "common.cuh"
#include "dropout.cuh"

namespace ergm {

constexpr int AT_BQ = 128, AT_BKV = 128, AT_D = 64;
constexpr int AT_THREADS = 256;
constexpr int AT_TILE = 128 * 64 * 2;  // 16 KB: one [128 rows x 64 d] bf16 tile
constexpr int AT_SMEM = AT_TILE /*Q*/ + 4 * AT_TILE /*K,V x2*/ + 2 * AT_TILE /*P*/ + 1024 + 256;

struct AttnFwdParams {
  __nv_bfloat16* out;  // [B*Tq, ld_out], head h at columns [h*64, h*64+64)
  float* lse;          // [B, nh, Tq]
  float* out_f32;      // nullable [B*Tq, nh*64]: un-rounded context, used by backward's delta
  const int* kv_lens;  // nullable [B]: keys >= kv_lens[b] are masked
  int64_t ld_out;
  int Tq, Tk, nh;
  int q_col0, k_col0, v_col0;
  int causal_off;      // query i may see key j iff j <= i + causal_off
  float scale;         // 1/sqrt(hd)
  DropoutSite drop;
  int do_drop;
};

template <bool CAUSAL>
__global__ void __launch_bounds__(AT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const AttnFwdParams p_in) {
  extern __shared__ uint8_t smem_raw[];
  AttnFwdParams p = p_in;
  p.drop = p_in.drop.resolved();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = base + AT_TILE, sV = base + 3 * AT_TILE, sP = base + 5 * AT_TILE;
  const uint32_t bars = base + 7 * AT_TILE;
  const uint32_t bar_q = bars, bar_s = bars + 8, bar_p = bars + 16, bar_o = bars + 24;
  auto kv_full = [&](int s) { return bars + 32 + 8u * s; };
  auto kv_empty = [&](int s) { return bars + 48 + 8u * s; };
  const uint32_t tmem_slot = bars + 64;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qb * AT_BQ;
  int kv_len = p.Tk;
  if (p.kv_lens) kv_len = min(kv_len, p.kv_lens[b]);
  int n_kv = (kv_len + AT_BKV - 1) / AT_BKV;
  if (CAUSAL) {
    const int last_key = min(q0 + AT_BQ - 1, p.Tq - 1) + p.causal_off;  // largest visible key
    n_kv = min(n_kv, last_key / AT_BKV + 1);
  }
  if (n_kv < 1) n_kv = 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 128); mbar_init(bar_o, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tS = tmem, tO = tmem + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar_q, AT_TILE);
      tma_load_3d(sQ, &tm_q, bar_q, p.q_col0 + h * AT_D, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(kv_empty(st), ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(kv_full(st), 2 * AT_TILE);
        tma_load_3d(sK + st * AT_TILE, &tm_k, kv_full(st), p.k_col0 + h * AT_D, j * AT_BKV, b);
        tma_load_3d(sV + st * AT_TILE, &tm_v, kv_full(st), p.v_col0 + h * AT_D, j * AT_BKV, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
      mbar_wait(bar_q, 0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(kv_full(st), (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < AT_D / 16; ++ks)
          umma_ss(tS, make_smem_desc_sw128(sQ + ks * 32, 16, 1024),
                  make_smem_desc_sw128(sK + st * AT_TILE + ks * 32, 16, 1024), idesc_s, ks > 0);
        umma_commit(bar_s);
        mbar_wait(bar_p, j & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < AT_BKV / 16; ++ks)
          umma_ss(tO, make_smem_desc_sw128(sP + (ks >> 2) * AT_TILE + (ks & 3) * 32, 16, 1024),
                  make_smem_desc_sw128(sV + st * AT_TILE + ks * 2048, 8192, 1024), idesc_o, ks > 0);
        umma_commit(bar_o);
        umma_commit(kv_empty(st));
      }
    }
  } else if (warp >= 4) {
    const int r = (warp & 3) * 32 + lane;  // query row inside the block == TMEM lane
    const int qi = q0 + r;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const float c = p.scale * 1.4426950408889634f;  // fold log2(e): exp(x*scale) = exp2(x*c)
    const float keep_scale = p.do_drop ? 1.f / (1.f - p.drop.p) : 1.f;
    const uint32_t drop_row = (uint32_t)((b * p.nh + h) * p.Tq + qi);
    const int vis = CAUSAL ? min(kv_len - 1, qi + p.causal_off) : kv_len - 1;  // last visible key
    float m = -INFINITY, l = 0.f;
    float o[AT_D];
#pragma unroll
    for (int i = 0; i < AT_D; ++i) o[i] = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      const int k0 = j * AT_BKV;
      // pass 1: row max over the visible keys of this block
      float mx = -INFINITY;
#pragma unroll 1
      for (int cc = 0; cc < AT_BKV; cc += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tS + lane_addr + cc, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (k0 + cc + i <= vis) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m, mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = exp2f((m - m_use) * c);  // m = -inf -> 0
      float sum = 0.f;
      // pass 2: probabilities -> bf16 -> swizzled smem (A operand of the PV MMA)
#pragma unroll 1
      for (int cc = 0; cc < AT_BKV; cc += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tS + lane_addr + cc, v);
        tmem_ld_wait();
        float pr[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e = exp2f((__uint_as_float(v[i]) - m_use) * c);
          pr[i] = (k0 + cc + i <= vis) ? e : 0.f;
          sum += pr[i];
        }
        if (p.do_drop) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const uint32_t k = p.drop.keep4(drop_row, (uint32_t)(k0 + cc + i) >> 2);
            pr[i] = (k & 1u) ? pr[i] * keep_scale : 0.f;
            pr[i + 1] = (k & 2u) ? pr[i + 1] * keep_scale : 0.f;
            pr[i + 2] = (k & 4u) ? pr[i + 2] * keep_scale : 0.f;
            pr[i + 3] = (k & 8u) ? pr[i + 3] * keep_scale : 0.f;
          }
        }
        const uint32_t rowbase = sP + (cc >> 6) * AT_TILE + r * 128;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          const uint32_t piece = (uint32_t)(((cc & 63) + i) >> 3);
          const uint32_t addr = rowbase + ((piece ^ (uint32_t)(r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                       "r"(pack_bf16x2(pr[i], pr[i + 1])), "r"(pack_bf16x2(pr[i + 2], pr[i + 3])),
                       "r"(pack_bf16x2(pr[i + 4], pr[i + 5])), "r"(pack_bf16x2(pr[i + 6], pr[i + 7]))
                       : "memory");
        }
      }
      l = l * alpha + sum;
      m = m_new;
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core
      tc_fence_before();
      mbar_arrive(bar_p);
      mbar_wait(bar_o, j & 1);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < AT_D; cc += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tO + lane_addr + cc, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[cc + i] = o[cc + i] * alpha + __uint_as_float(v[i]);
      }
      tc_fence_before();
    }
    if (qi < p.Tq) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      __nv_bfloat16* op = p.out + ((int64_t)b * p.Tq + qi) * p.ld_out + h * AT_D;
#pragma unroll
      for (int i = 0; i < AT_D; i += 8) {
        const uint4 u = make_uint4(pack_bf16x2(o[i] * inv, o[i + 1] * inv), pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv),
                                   pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv), pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv));
        *reinterpret_cast<uint4*>(op + i) = u;
      }
      if (p.out_f32) {
        float* of = p.out_f32 + ((int64_t)b * p.Tq + qi) * (p.nh * AT_D) + h * AT_D;
#pragma unroll
        for (int i = 0; i < AT_D; i += 4)
          *reinterpret_cast<float4*>(of + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
      }
      if (p.lse) p.lse[((int64_t)b * p.nh + h) * p.Tq + qi] = (m == -INFINITY ? 0.f : m) * p.scale + logf(l);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 256);
}

}  // namespace ergm

using namespace ergm;

extern "C" int ergm_attn_fwd(const void* q, int64_t ld_q, int q_col0, const void* k, int64_t ld_k,
                             int k_col0, const void* v, int64_t ld_v, int v_col0, void* out,
                             int64_t ld_out, float* out_f32, float* lse, const int* kv_lens, int B, int nh, int Tq,
                             int Tk, int head_dim, int causal, int causal_off, float dropout_p,
                             uint64_t seed, uint64_t offset, void* stream) {
  if (!q || !k || !v || !out || B <= 0 || nh <= 0 || Tq <= 0 || Tk <= 0) return ERGM_ERR_ARG;
  if (head_dim != AT_D) return ERGM_ERR_UNSUPPORTED;
  if (ld_q % 8 || ld_k % 8 || ld_v % 8 || ld_out % 8 || q_col0 % 8 || k_col0 % 8 || v_col0 % 8)
    return ERGM_ERR_ARG;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = encode_tmap_3d(&tq, q, 2, (uint64_t)(q_col0 + nh * AT_D), (uint64_t)Tq, (uint64_t)B,
                           (uint64_t)ld_q * 2, (uint64_t)Tq * ld_q * 2, AT_D, AT_BQ, 1)))
    return rc;
  if ((rc = encode_tmap_3d(&tk, k, 2, (uint64_t)(k_col0 + nh * AT_D), (uint64_t)Tk, (uint64_t)B,
                           (uint64_t)ld_k * 2, (uint64_t)Tk * ld_k * 2, AT_D, AT_BKV, 1)))
    return rc;
  if ((rc = encode_tmap_3d(&tv, v, 2, (uint64_t)(v_col0 + nh * AT_D), (uint64_t)Tk, (uint64_t)B,
                           (uint64_t)ld_v * 2, (uint64_t)Tk * ld_v * 2, AT_D, AT_BKV, 1)))
    return rc;
  AttnFwdParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse; p.out_f32 = out_f32; p.kv_lens = kv_lens; p.ld_out = ld_out;
  p.Tq = Tq; p.Tk = Tk; p.nh = nh;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.causal_off = causal_off;
  p.scale = 1.0f / sqrtf((float)head_dim);
  p.drop = make_site(seed, offset, dropout_p, (uint32_t)Tk);
  p.do_drop = dropout_p > 0.f;
  static bool attr = false;
  if (!attr) {
    ERGM_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    ERGM_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    attr = true;
  }
  dim3 grid((Tq + AT_BQ - 1) / AT_BQ, nh, B);
  if (causal)
    attn_fwd_kernel<true><<<grid, AT_THREADS, AT_SMEM, (cudaStream_t)stream>>>(tq, tk, tv, p);
  else
    attn_fwd_kernel<false><<<grid, AT_THREADS, AT_SMEM, (cudaStream_t)stream>>>(tq, tk, tv, p);
  return (int)cudaGetLastError();
}

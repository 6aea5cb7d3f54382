// ergm_attn_fwd — fused flash-style attention forward on tcgen05 / TMEM for head_dim 64.
//
// Replaces GPT2Attention._attn (/root/reference/src/model.py:119-148): QK^T, /sqrt(hd),
// causal where-mask (self) or none (cross), softmax in fp32, attn_dropout, PV, and the
// _split_heads / _merge_heads permutes (:190-198) — Q/K/V are read straight out of the
// [rows, ld] projection outputs with 3-D TMA maps (col, row-in-sequence, batch) and the
// context is written merged as [B*Tq, nh*64].
//
// One CTA = one (batch, head, 128-query block); TWO CTAs are resident per SM (82 KB smem, 256 TMEM
// columns, <= 80 registers x 384 threads each) so that one CTA's exp phase (MUFU-bound) overlaps
// the other's load / MMA / barrier latencies.  Warp roles: 0 = TMA, 1 = MMA issuer, 2 = TMEM
// allocator, 4..11 = softmax.  Two softmax threads share a query row (TMEM lane): each owns 64 of
// the 128 key columns of the current block and 32 of the 64 output columns; the row maximum is
// exchanged through shared memory once per key block, the row sum once at the end.
//   S   = Q K_j^T           SS MMA M128 N128 K64    -> TMEM cols [0,128)
//   P   = exp2(S c - m c)   -> bf16 pairs, tcgen05.st -> TMEM cols [128,192)   (never touches smem)
//   O_j = P V_j             TS MMA (A from TMEM) M128 N64 K128 -> TMEM cols [192,256)
//   O   = O * alpha + O_j   (registers)
// Masks are only evaluated on key blocks that need them (diagonal block / key-length tail).
#include "../../include/ergm_b200.h"
#include "common.cuh"
#include "dropout.cuh"

namespace ergm {


constexpr int AT_D = 64;
constexpr int AT_THREADS = 384;
constexpr int AT_TILE = 128 * 64 * 2;  // 16 KB: one [128 rows x 64 d] bf16 tile
// Q | K x2 | V x2 | xch 1 KB | barriers ; + 1 KB alignment slack
constexpr int AT_SMEM = 5 * AT_TILE + 1024 + 128 + 1024;

struct AttnFwdParams {
  __nv_bfloat16* out;  // [B*Tq, ld_out], head h at columns [h*64, h*64+64)
  float* lse;          // [B, nh, Tq]
  float* out_f32;      // nullable [B*Tq, nh*64]: un-rounded context, used by backward's delta
  const int* kv_lens;  // nullable [B]: keys >= kv_lens[b] are masked
  const int* cu_q;     // nullable [B+1]: packed batch, queries (and outputs) of sample b are rows cu_q[b] .. cu_q[b+1]
  const int* cu_k;     // nullable [B+1]: keys / values packed the same way (self attention); NULL: rows b*Tk ..
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
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = base + AT_TILE, sV = base + 3 * AT_TILE;
  const uint32_t sX = base + 5 * AT_TILE;  // float xch[2 halves][128]
  const uint32_t bars = sX + 1024;
  const uint32_t bar_q = bars, bar_s = bars + 8, bar_p = bars + 16, bar_o = bars + 24;
  auto kv_full = [&](int s) { return bars + 32 + 8u * s; };
  auto kv_empty = [&](int s) { return bars + 48 + 8u * s; };
  const uint32_t tmem_slot = bars + 64;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();               // programmatic dependent launch: the CTA may be resident before the previous kernel ended
  pdl_launch_dependents();
  p.drop = p_in.drop.resolved();
  // heaviest CTAs first: causal query block qb visits qb + 1 key blocks, so the LAST query blocks are scheduled
  // first (blockIdx.z is the slowest-varying index of the block scheduler) and the light ones fill the tail
  const int qb = (int)gridDim.z - 1 - (int)blockIdx.z, h = blockIdx.y, b = blockIdx.x;
  const int q0 = qb * 128;
  // packed batches: per-sample extents and row bases (TMA batch coordinate 0: the tensor is one long row sequence)
  int q_row0 = 0, q_bat = b, k_row0 = 0, k_bat = b;
  int64_t out_row0 = (int64_t)b * p.Tq;
  if (p.cu_q) {
    q_row0 = p.cu_q[b]; q_bat = 0; out_row0 = q_row0;
    p.Tq = p.cu_q[b + 1] - q_row0;          // p is this thread's copy: from here on Tq is the sample's own
    if (q0 >= p.Tq) return;                 // CTA-uniform, before any barrier / TMEM allocation
  }
  if (p.cu_k) { k_row0 = p.cu_k[b]; k_bat = 0; p.Tk = p.cu_k[b + 1] - k_row0; }
  const int64_t stat_row0 = ((int64_t)b * p.nh + h) * p_in.Tq;   // lse / dropout rows keep the padded [B, nh, T] indexing
  int kv_len = p.Tk;
  if (p.kv_lens) kv_len = min(kv_len, p.kv_lens[b]);
  int n_kv = max(1, (kv_len + 127) / 128);
  if (CAUSAL) {
    const int last_key = min(q0 + 127, p.Tq - 1) + p.causal_off;  // largest visible key
    n_kv = max(1, min(n_kv, max(0, last_key) / 128 + 1));
  }

  if (warp == 0 && lane == 0) {
    // the producer owns the load barriers and fires Q / K_0 / V_0 BEFORE the block-wide sync below, so the
    // first TMA round trip (~1 us) overlaps the TMEM allocation instead of following it
    mbar_init(bar_q, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    fence_mbar_init();
    mbar_expect_tx(bar_q, AT_TILE);
    tma_load_3d(sQ, &tm_q, bar_q, p.q_col0 + h * AT_D, q_row0 + q0, q_bat);
    mbar_expect_tx(kv_full(0), 2 * AT_TILE);
    tma_load_3d(sK, &tm_k, kv_full(0), p.k_col0 + h * AT_D, k_row0, k_bat);
    tma_load_3d(sV, &tm_v, kv_full(0), p.v_col0 + h * AT_D, k_row0, k_bat);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar_s, 1); mbar_init(bar_p, 256); mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tS = tmem, tP = tmem + 128, tO = tmem + 192;

  // Register reallocation between the warpgroups (setmaxnreg): the kernel is compiled for 80 registers per thread
  // (two CTAs per SM); the helper warpgroup (TMA / MMA issue / TMEM allocation) gives most of its share to the
  // two softmax warpgroups, whose per-row state (o[32] + a 32-column S chunk + its exponentials) spilled at 80.
#ifdef ERGM_ATTN_SETMAXNREG
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
  }
#endif
  if (warp == 0) {
    if (lane == 0) {
      for (int j = 1; j < n_kv; ++j) {
        const int st = j & 1;
        mbar_wait(kv_empty(st), ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(kv_full(st), 2 * AT_TILE);
        tma_load_3d(sK + st * AT_TILE, &tm_k, kv_full(st), p.k_col0 + h * AT_D, k_row0 + j * 128, k_bat);
        tma_load_3d(sV + st * AT_TILE, &tm_v, kv_full(st), p.v_col0 + h * AT_D, k_row0 + j * 128, k_bat);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
      auto issue_s = [&](int j) {
        const uint32_t k_t = sK + (j & 1) * AT_TILE;
#pragma unroll
        for (int ks = 0; ks < AT_D / 16; ++ks)
          umma_ss(tS, make_smem_desc_sw128(sQ + ks * 32, 16, 1024),
                  make_smem_desc_sw128(k_t + ks * 32, 16, 1024), idesc_s, ks > 0);
        umma_commit(bar_s);
      };
      mbar_wait(bar_q, 0);
      mbar_wait(kv_full(0), 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const uint32_t v_t = sV + st * AT_TILE;
        mbar_wait(bar_p, j & 1);  // P_j in TMEM, S consumed
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)   // A = P (TMEM, 8 columns = 16 bf16 per K step)
          umma_ts(tO, tP + ks * 8, make_smem_desc_sw128(v_t + ks * 2048, 8192, 1024), idesc_o, ks > 0);
        umma_commit(bar_o);
        umma_commit(kv_empty(st));
        if (j + 1 < n_kv) {
          mbar_wait(kv_full((j + 1) & 1), ((j + 1) >> 1) & 1);
          tc_fence_after();
          issue_s(j + 1);
        }
      }
    }
  } else if (warp >= 4) {
    const int hf = (warp - 4) >> 2;          // which half of the key columns / output columns
    const int r = (warp & 3) * 32 + lane;    // query row inside the block == TMEM lane
    const int qi = q0 + r;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS_mine = tS + lane_addr + 64 * hf;
    const uint32_t tP_mine = tP + lane_addr + 32 * hf;
    const uint32_t tO_mine = tO + lane_addr + 32 * hf;
    const float c = p.scale * 1.4426950408889634f;  // exp(x*scale) = exp2(x*c)
    const uint32_t thr16 = p.drop.thr16();
    const float keep_scale = p.do_drop ? p.drop.keep_scale() : 1.f;
    const uint32_t drop_row = (uint32_t)(stat_row0 + qi);
    const int vis = CAUSAL ? min(kv_len - 1, qi + p.causal_off) : kv_len - 1;  // last visible key
    const uint32_t xch_mine = sX + (hf * 128 + r) * 4;
    const uint32_t xch_other = sX + ((hf ^ 1) * 128 + r) * 4;
    float m = -INFINITY, l = 0.f;
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int k0 = j * 128 + 64 * hf;  // first key column this thread owns in block j
      // warp-uniform: does any row of this warp need masking in this block?
      const bool need_mask = (j * 128 + 127 > kv_len - 1) ||
                             (CAUSAL && (j * 128 + 127 > q0 + (warp & 3) * 32 + p.causal_off));
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      float mx = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < 64; cc += 16) {  // 16-column chunks: half the live registers of a 32-column one
        uint32_t v[16];
        tmem_ld_32x32b_x16(tS_mine + cc, v);
        tmem_ld_wait();
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (k0 + cc + i <= vis) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[i]));
        }
        mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
      }
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch_mine), "f"(mx) : "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float mo;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mo) : "r"(xch_other) : "memory");
      const float m_new = fmaxf(m, fmaxf(mx, mo));
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = ex2_fast((m - m_use) * c);  // m = -inf -> 0
      const float mc = m_use * c;
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int cc = 0; cc < 64; cc += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tS_mine + cc, v);
        tmem_ld_wait();
        float pr[16];
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float e = ex2_fast(fmaf(__uint_as_float(v[i]), c, -mc));
            pr[i] = (k0 + cc + i <= vis) ? e : 0.f;
            s4[i & 3] += pr[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            pr[i] = ex2_fast(fmaf(__uint_as_float(v[i]), c, -mc));
            s4[i & 3] += pr[i];
          }
        }
        if (p.do_drop) {
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const uint32_t hsh = p.drop.hash2(drop_row, (uint32_t)(k0 + cc + i) >> 1);
            pr[i] = ((hsh & 0xffffu) >= thr16) ? pr[i] * keep_scale : 0.f;
            pr[i + 1] = ((hsh >> 16) >= thr16) ? pr[i + 1] * keep_scale : 0.f;
          }
        }
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(pr[2 * i], pr[2 * i + 1]);
        tmem_st_32x32b_x8(tP_mine + (cc >> 1), pk);   // 16 keys -> 8 packed columns
      }
      l = l * alpha + ((s4[0] + s4[1]) + (s4[2] + s4[3]));
      m = m_new;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p);
      mbar_wait(bar_o, j & 1);
      tc_fence_after();
      {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tO_mine, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = fmaf(o[i], alpha, __uint_as_float(v[i]));
      }
      tc_fence_before();
    }
    // combine the two half-row sums
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch_mine), "f"(l) : "memory");
    asm volatile("bar.sync 1, 256;" ::: "memory");
    float lo;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(lo) : "r"(xch_other) : "memory");
    l += lo;
    if (qi < p.Tq) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      __nv_bfloat16* op = p.out + (out_row0 + qi) * p.ld_out + h * AT_D + 32 * hf;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        const uint4 u = make_uint4(pack_bf16x2(o[i] * inv, o[i + 1] * inv), pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv),
                                   pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv), pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv));
        *reinterpret_cast<uint4*>(op + i) = u;
      }
      if (p.out_f32) {
        float* of = p.out_f32 + (out_row0 + qi) * (p.nh * AT_D) + h * AT_D + 32 * hf;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(of + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
      }
      if (p.lse && hf == 0)
        p.lse[stat_row0 + qi] = (m == -INFINITY ? 0.f : m) * p.scale + logf(l);
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
                             int64_t ld_out, float* out_f32, float* lse, const int* kv_lens, int B,
                             int nh, int Tq, int Tk, int head_dim, int causal, int causal_off,
                             float dropout_p, uint64_t seed, uint64_t offset, const ergm_pack* pack, int pack_kv,
                             void* stream) {
  if (!q || !k || !v || !out || B <= 0 || nh <= 0 || Tq <= 0 || Tk <= 0) return ERGM_ERR_ARG;
  if (pack && (!pack->cu_rows || (pack_kv && !pack->kv_lens))) return ERGM_ERR_ARG;
  // packed batch: Q (and K / V when pack_kv) are ONE row sequence of capacity B * T; sample b starts at cu_rows[b]
  const uint64_t q_rows = pack ? (uint64_t)B * Tq : (uint64_t)Tq, q_bat = pack ? 1 : (uint64_t)B;
  const uint64_t k_rows = (pack && pack_kv) ? (uint64_t)B * Tk : (uint64_t)Tk, k_bat = (pack && pack_kv) ? 1 : (uint64_t)B;
  if (head_dim != AT_D) return ERGM_ERR_UNSUPPORTED;
  if (ld_q % 8 || ld_k % 8 || ld_v % 8 || ld_out % 8 || q_col0 % 8 || k_col0 % 8 || v_col0 % 8)
    return ERGM_ERR_ARG;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = encode_tmap_3d(&tq, q, 2, (uint64_t)(q_col0 + nh * AT_D), q_rows, q_bat,
                           (uint64_t)ld_q * 2, q_rows * ld_q * 2, AT_D, 128, 1)))
    return rc;
  if ((rc = encode_tmap_3d(&tk, k, 2, (uint64_t)(k_col0 + nh * AT_D), k_rows, k_bat,
                           (uint64_t)ld_k * 2, k_rows * ld_k * 2, AT_D, 128, 1)))
    return rc;
  if ((rc = encode_tmap_3d(&tv, v, 2, (uint64_t)(v_col0 + nh * AT_D), k_rows, k_bat,
                           (uint64_t)ld_v * 2, k_rows * ld_v * 2, AT_D, 128, 1)))
    return rc;
  AttnFwdParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse; p.out_f32 = out_f32; p.kv_lens = kv_lens; p.ld_out = ld_out;
  p.cu_q = pack ? pack->cu_rows : nullptr;
  p.cu_k = (pack && pack_kv) ? pack->cu_rows : nullptr;
  if (pack && pack_kv) p.kv_lens = pack->kv_lens;
  p.Tq = Tq; p.Tk = Tk; p.nh = nh;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.causal_off = causal_off;
  p.scale = 1.0f / sqrtf((float)head_dim);
  p.drop = make_site(seed, offset, dropout_p, (uint32_t)Tk);
  p.do_drop = dropout_p > 0.f;
  {
    static std::atomic<uint64_t> done_mask{0};
    int dev = 0;
    ERGM_CUDA_TRY(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (!(done_mask.load(std::memory_order_acquire) & bit)) {
      ERGM_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
      ERGM_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
      // two CTAs per SM need 2 x 83 KB of shared memory: ask for the maximum shared-memory carve-out
      ERGM_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
      ERGM_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared));
      done_mask.fetch_or(bit, std::memory_order_release);
    }
  }
  dim3 grid(B, nh, (Tq + 127) / 128);
  if (causal)
    return (int)launch_pdl(attn_fwd_kernel<true>, grid, dim3(AT_THREADS), (size_t)AT_SMEM, (cudaStream_t)stream, 1, tq, tk, tv, p);
  return (int)launch_pdl(attn_fwd_kernel<false>, grid, dim3(AT_THREADS), (size_t)AT_SMEM, (cudaStream_t)stream, 1, tq, tk, tv, p);
}

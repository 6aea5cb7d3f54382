// ergm_attn_bwd — fused attention backward on tcgen05 / TMEM, head_dim 64.
//
// Gradient of GPT2Attention._attn (/root/reference/src/model.py:119-148) w.r.t. Q, K, V given
// dO, recomputing the probabilities from the saved log-sum-exp instead of storing the
// [B,nh,T,T] score / probability tensors the reference keeps for autograd.
//
// One CTA = one (batch, head, 128-key block j); it loops over the query blocks i that can see
// those keys.  Everything is computed TRANSPOSED (keys on TMEM lanes, queries on columns) so
// that P^T and dS^T come out of the softmax warps exactly in the layout the next MMAs need:
//   S^T  = K_j Q_i^T                      SS MMA  M128(kv) N128(q) K64      -> TMEM [0,128)
//   dP^T = V_j dO_i^T                     SS MMA                            -> TMEM [128,256)
//   P^T  = exp2(S^T c - lse_i log2e),  dS^T = P^T (dP^T - delta_i) scale    (8 compute warps)
//   dV_j += P^T dO_i                      SS MMA  M128(kv) N64 K128(q)      -> TMEM [256,320)
//   dK_j += dS^T Q_i                      SS MMA                            -> TMEM [320,384)
//   dQ_i  = dS K_j                        SS MMA  M128(q)  N64 K128(kv)     -> TMEM [384,448)
// The smem tile holding dS^T ([kv rows][q contiguous], 128B swizzle) is at the same time the
// K-major A operand of the dK product and the MN-major A operand of the dQ product, and the
// TMA tiles of Q_i / dO_i / K_j serve as K-major and MN-major B operands without any copy.
// dQ is accumulated across key blocks with red.global.add.v4.f32 into an fp32 buffer.
#include "../../include/ergm_b200.h"
#include "common.cuh"
#include "dropout.cuh"

namespace ergm {

// 4 helper warps (TMA producer, MMA issuer, TMEM allocator, spare) + 16 element-wise warps.  The first version had 8
// element-wise warps (64 score elements per thread and iteration, 168 registers): with one CTA per SM (512 TMEM
// columns) that left 2 warps per scheduler to hide the TMEM-load / MUFU / shared-store latencies of the
// P^T / dS^T stage, and ncu showed 18 % warps active, 21 % issue slots used.  16 warps halve each thread's share
// (one 32-column slab in two 16-column pieces, ~100 registers) and double the latency hiding.
constexpr int AB_CWARPS = 16;
constexpr int AB_THREADS = 128 + AB_CWARPS * 32;
constexpr int AB_TILE = 128 * 64 * 2;  // 16 KB
// K, V, 2x(Q,dO), PT (2 tiles), dST (2 tiles), lse/delta (2 stages x 2 x 128 floats)
constexpr int AB_SMEM = 2 * AB_TILE + 4 * AB_TILE + 2 * AB_TILE + 2 * AB_TILE + 2048 + 1024 + 256;

struct AttnBwdParams {
  const float* lse;     // [B, nh, Tq]
  const float* delta;   // [B, nh, Tq]
  float* dq_accum;      // fp32 [B*Tq, ld_dq], head h at columns [h*64, ...)
  __nv_bfloat16* dk;    // [B*Tk, ld_dk], head h at columns [dk_col0 + h*64, ...)
  __nv_bfloat16* dv;
  float* dk_colsum;     // nullable [nh*64]: += column sums of dK as stored (bias gradient of the K projection)
  float* dv_colsum;     // nullable [nh*64]
  const int* kv_lens;
  const int* cu_q;      // nullable [B+1]: packed batch (queries, dO, dQ rows of sample b start at cu_q[b])
  const int* cu_k;      // nullable [B+1]: keys / values / dK / dV packed the same way (self attention)
  int64_t ld_dq, ld_dk, ld_dv;
  int dk_col0, dv_col0;
  int Tq, Tk, nh;
  int q_col0, k_col0, v_col0;
  int causal_off;
  float scale;
  DropoutSite drop;
  int do_drop;
};

// Bias gradient of the K / V projection: column sums of this warp's [32 keys x 32 columns] slab of dK / dV,
// taken over the bf16-ROUNDED values that are stored (what a separate column-sum pass over dK / dV would
// see), one fp32 atomic per column per warp.  Replaces a full extra read of dK / dV per attention.
template <int N>
ERGM_DEVINL void colsum_fold(float (&x)[32], int lane) {
  // lanes with bit N set keep the upper N columns, the others the lower N; each adds what its partner held
  const bool upper = (lane & N) != 0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float send = upper ? x[i] : x[i + N];
    const float recv = __shfl_xor_sync(0xffffffffu, send, N);
    x[i] = (upper ? x[i + N] : x[i]) + recv;
  }
}
ERGM_DEVINL void attn_bwd_colsum(const uint32_t (&v)[32], bool row_ok, float* dst, int lane) {
  float x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = row_ok ? __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[i]))) : 0.f;
  // transpose-reduce butterfly: 16 + 8 + 4 + 2 + 1 = 31 shuffles; lane l ends up with column l's total
  colsum_fold<16>(x, lane); colsum_fold<8>(x, lane); colsum_fold<4>(x, lane); colsum_fold<2>(x, lane); colsum_fold<1>(x, lane);
  atomicAdd(dst + lane, x[0]);
}

// 16-column variant: this warp's [32 keys x 16 columns] slab -> lanes 0..15 add column totals
ERGM_DEVINL void attn_bwd_colsum16(const uint32_t (&v)[16], bool row_ok, float* dst, int lane) {
  float x[32];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = row_ok ? __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[i]))) : 0.f;
  // folds over lane bits 8, 4, 2, 1 leave column (lane & 15) summed over the 16 lanes sharing bit 16; one more
  // exchange joins the two halves
  colsum_fold<8>(x, lane); colsum_fold<4>(x, lane); colsum_fold<2>(x, lane); colsum_fold<1>(x, lane);
  const float tot = x[0] + __shfl_xor_sync(0xffffffffu, x[0], 16);
  if (lane < 16) atomicAdd(dst + lane, tot);
}

template <bool CAUSAL>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                const AttnBwdParams p_in) {
  extern __shared__ uint8_t smem_raw[];
  AttnBwdParams p = p_in;
  p.drop = p_in.drop.resolved();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = base, sV = base + AB_TILE;
  const uint32_t sQ = base + 2 * AB_TILE;    // 2 stages
  const uint32_t sDO = base + 4 * AB_TILE;   // 2 stages
  const uint32_t sPT = base + 6 * AB_TILE;   // 2 tiles (q chunks)
  const uint32_t sDS = base + 8 * AB_TILE;   // 2 tiles
  const uint32_t sStat = base + 10 * AB_TILE;  // [2 stages][lse 128 | delta 128] floats
  const uint32_t bars = sStat + 2048;
  const uint32_t bar_kv = bars, bar_sdp = bars + 8, bar_pds = bars + 16, bar_dq = bars + 24;
  auto qdo_full = [&](int s) { return bars + 32 + 8u * s; };
  auto qdo_empty = [&](int s) { return bars + 48 + 8u * s; };
  const uint32_t tmem_slot = bars + 64;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // heaviest CTAs first: causal key block jb is visited by the query blocks >= jb, so key block 0 has the most
  // work; blockIdx.z is the slowest-varying index of the block scheduler
  const int jb = blockIdx.z, h = blockIdx.y, b = blockIdx.x;
  const int k0 = jb * 128;
  // packed batches: per-sample extents and row bases (TMA batch coordinate 0: one long row sequence)
  int q_row0 = 0, q_bat = b, k_row0 = 0, k_bat = b;
  int64_t dq_row0 = (int64_t)b * p.Tq, dk_row0 = (int64_t)b * p.Tk;
  if (p.cu_q) { q_row0 = p.cu_q[b]; q_bat = 0; dq_row0 = q_row0; p.Tq = p.cu_q[b + 1] - q_row0; }
  if (p.cu_k) {
    k_row0 = p.cu_k[b]; k_bat = 0; dk_row0 = k_row0; p.Tk = p.cu_k[b + 1] - k_row0;
    if (k0 >= p.Tk) return;   // CTA-uniform, before any barrier / TMEM allocation: no key rows of this sample here
  }
  const int64_t stat_row0 = ((int64_t)b * p.nh + h) * p_in.Tq;   // lse / delta / dropout rows: padded [B, nh, T] indexing
  int kv_len = p.Tk;
  if (p.kv_lens) kv_len = min(kv_len, p.kv_lens[b]);
  const int n_q = (p.Tq + 127) / 128;
  int i_min = 0;
  if (CAUSAL) i_min = max(0, k0 - p.causal_off) / 128;
  const bool active = (k0 < kv_len) && (i_min < n_q);  // CTA-uniform
  const int n_iter = active ? n_q - i_min : 0;

  if (warp == 0 && lane == 0) {
    // the producer owns the load barriers and fires K / V and the first Q / dO stage BEFORE the block-wide
    // sync below: the first TMA round trip overlaps the 512-column TMEM allocation (one CTA per SM here,
    // so nothing else hides the prologue)
    mbar_init(bar_kv, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(qdo_full(s), 1); mbar_init(qdo_empty(s), 1); }
    fence_mbar_init();
    if (active) {
      mbar_expect_tx(bar_kv, 2 * AB_TILE);
      tma_load_3d(sK, &tm_k, bar_kv, p.k_col0 + h * 64, k_row0 + k0, k_bat);
      tma_load_3d(sV, &tm_v, bar_kv, p.v_col0 + h * 64, k_row0 + k0, k_bat);
      mbar_expect_tx(qdo_full(0), 2 * AB_TILE);
      tma_load_3d(sQ, &tm_q, qdo_full(0), p.q_col0 + h * 64, q_row0 + i_min * 128, q_bat);
      tma_load_3d(sDO, &tm_do, qdo_full(0), h * 64, q_row0 + i_min * 128, q_bat);
    }
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar_sdp, 1); mbar_init(bar_pds, AB_CWARPS * 32); mbar_init(bar_dq, 1);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tS = tmem, tDP = tmem + 128, tDV = tmem + 256, tDK = tmem + 320, tDQ = tmem + 384;

  if (warp == 0) {
    if (lane == 0 && active) {
      for (int it = 1; it < n_iter; ++it) {
        const int st = it & 1;
        mbar_wait(qdo_empty(st), ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(qdo_full(st), 2 * AB_TILE);
        tma_load_3d(sQ + st * AB_TILE, &tm_q, qdo_full(st), p.q_col0 + h * 64, q_row0 + (i_min + it) * 128, q_bat);
        tma_load_3d(sDO + st * AB_TILE, &tm_do, qdo_full(st), h * 64, q_row0 + (i_min + it) * 128, q_bat);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && active) {
      const uint32_t id_st = make_idesc_bf16(128, 128, 0, 0);   // K-major x K-major
      const uint32_t id_dkv = make_idesc_bf16(128, 64, 0, 1);   // A K-major, B MN-major
      const uint32_t id_dq = make_idesc_bf16(128, 64, 1, 1);    // A MN-major, B MN-major
      mbar_wait(bar_kv, 0);
      for (int it = 0; it < n_iter; ++it) {
        const int st = it & 1;
        const uint32_t q_t = sQ + st * AB_TILE, do_t = sDO + st * AB_TILE;
        mbar_wait(qdo_full(st), (it >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss(tS, make_smem_desc_sw128(sK + ks * 32, 16, 1024),
                  make_smem_desc_sw128(q_t + ks * 32, 16, 1024), id_st, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss(tDP, make_smem_desc_sw128(sV + ks * 32, 16, 1024),
                  make_smem_desc_sw128(do_t + ks * 32, 16, 1024), id_st, ks > 0);
        umma_commit(bar_sdp);
        mbar_wait(bar_pds, it & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t a_off = (ks >> 2) * AB_TILE + (ks & 3) * 32;
          umma_ss(tDV, make_smem_desc_sw128(sPT + a_off, 16, 1024),
                  make_smem_desc_sw128(do_t + ks * 2048, 8192, 1024), id_dkv, (it > 0 || ks > 0));
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t a_off = (ks >> 2) * AB_TILE + (ks & 3) * 32;
          umma_ss(tDK, make_smem_desc_sw128(sDS + a_off, 16, 1024),
                  make_smem_desc_sw128(q_t + ks * 2048, 8192, 1024), id_dkv, (it > 0 || ks > 0));
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_ss(tDQ, make_smem_desc_sw128(sDS + ks * 2048, AB_TILE, 1024),
                  make_smem_desc_sw128(sK + ks * 2048, 8192, 1024), id_dq, ks > 0);
        umma_commit(bar_dq);
        umma_commit(qdo_empty(st));
      }
    }
  } else if (warp >= 4 && active) {
    const int qr = warp & 3;             // TMEM lane quarter
    const int cg = (warp - 4) >> 2;      // column group: q columns [32*cg, 32*cg+32) of the 128-query block
    const int r = qr * 32 + lane;        // key row inside the block == TMEM lane
    const int kj = k0 + r;
    const bool key_ok = kj < kv_len;
    const uint32_t lane_addr = (uint32_t)(qr * 32) << 16;
    const float c = p.scale * 1.4426950408889634f;
    const uint32_t thr16 = p.drop.thr16();
    const float keep_scale = p.do_drop ? 65536.f / (65536.f - (float)thr16) : 1.f;
    const uint32_t drop_shift = (kj & 1) ? 16u : 0u;  // which 16-bit lane of hash2(row, kj >> 1) is ours
    const int ct = threadIdx.x - 128;    // 0..511 inside the compute group
    for (int it = 0; it < n_iter; ++it) {
      const int q0 = (i_min + it) * 128;
      // stage this query block's lse (pre-multiplied by log2 e) / delta in smem
      if (ct < 256) {
        const uint32_t dst = sStat + (it & 1) * 1024 + ct * 4;
        const int qi = q0 + (ct & 127);
        float val = 0.f;
        if (qi < p.Tq) {
          const int64_t idx = stat_row0 + qi;
          val = ct < 128 ? p.lse[idx] * 1.4426950408889634f : p.delta[idx];
        }
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(dst), "f"(val) : "memory");
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      // warp-uniform: does this (key rows, query block) tile need any masking?
      const bool need_mask = (k0 + qr * 32 + 31 > kv_len - 1) || (q0 + 127 > p.Tq - 1) ||
                             (CAUSAL && (k0 + qr * 32 + 31 > q0 + p.causal_off));
      mbar_wait(bar_sdp, it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 32 * cg; cc < 32 * cg + 32; cc += 16) {
        uint32_t sv[16], dv_[16];
        tmem_ld_32x32b_x16(tS + lane_addr + cc, sv);
        tmem_ld_32x32b_x16(tDP + lane_addr + cc, dv_);
        float ls[16], dl[16];
        const uint32_t stat = sStat + (it & 1) * 1024 + cc * 4;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(ls[i]), "=f"(ls[i + 1]), "=f"(ls[i + 2]), "=f"(ls[i + 3]) : "r"(stat + i * 4));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(dl[i]), "=f"(dl[i + 1]), "=f"(dl[i + 2]), "=f"(dl[i + 3]) : "r"(stat + 512 + i * 4));
        }
        tmem_ld_wait();
        float pt[16], ds[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float pr = ex2_fast(fminf(fmaf(__uint_as_float(sv[i]), c, -ls[i]), 0.f));
          if (need_mask) {
            const int qi = q0 + cc + i;
            bool vis = key_ok && qi < p.Tq;
            if (CAUSAL) vis = vis && (kj <= qi + p.causal_off);
            pr = vis ? pr : 0.f;
          }
          float dp = __uint_as_float(dv_[i]);
          float pd = pr;
          if (p.do_drop) {
            const uint32_t hsh = p.drop.hash2((uint32_t)(stat_row0 + q0 + cc + i), (uint32_t)kj >> 1);
            const bool keep = ((hsh >> drop_shift) & 0xffffu) >= thr16;
            dp = keep ? dp * keep_scale : 0.f;
            pd = keep ? pr * keep_scale : 0.f;
          }
          pt[i] = pd;
          ds[i] = pr * (dp - dl[i]) * p.scale;
        }
        const uint32_t rowoff = (cc >> 6) * AB_TILE + r * 128;
#pragma unroll
        for (int i = 0; i < 16; i += 8) {
          const uint32_t piece = (uint32_t)(((cc & 63) + i) >> 3);
          const uint32_t off = rowoff + ((piece ^ (uint32_t)(r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sPT + off),
                       "r"(pack_bf16x2(pt[i], pt[i + 1])), "r"(pack_bf16x2(pt[i + 2], pt[i + 3])),
                       "r"(pack_bf16x2(pt[i + 4], pt[i + 5])), "r"(pack_bf16x2(pt[i + 6], pt[i + 7]))
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sDS + off),
                       "r"(pack_bf16x2(ds[i], ds[i + 1])), "r"(pack_bf16x2(ds[i + 2], ds[i + 3])),
                       "r"(pack_bf16x2(ds[i + 4], ds[i + 5])), "r"(pack_bf16x2(ds[i + 6], ds[i + 7]))
                       : "memory");
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_pds);
      // dQ_i: TMEM lanes are queries here; this warp handles rows qr*32.., columns 16*cg..+16
      mbar_wait(bar_dq, it & 1);
      tc_fence_after();
      {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tDQ + lane_addr + 16 * cg, v);
        tmem_ld_wait();
        const int qi = q0 + r;
        if (qi < p.Tq) {
          float* dst = p.dq_accum + (dq_row0 + qi) * p.ld_dq + h * 64 + 16 * cg;
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i),
                         "f"(__uint_as_float(v[i])), "f"(__uint_as_float(v[i + 1])),
                         "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3]))
                         : "memory");
        }
      }
      tc_fence_before();
    }
    // dK_j, dV_j: lanes are keys; this warp writes rows qr*32.., columns 16*cg..+16.
    // tcgen05.ld is warp-collective (.sync.aligned): issue it unconditionally, guard the stores.
    {
      uint32_t v[16];
      tmem_ld_32x32b_x16(tDK + lane_addr + 16 * cg, v);
      tmem_ld_wait();
      if (kj < p.Tk) {
        __nv_bfloat16* dkp = p.dk + (dk_row0 + kj) * p.ld_dk + p.dk_col0 + h * 64 + 16 * cg;
#pragma unroll
        for (int i = 0; i < 16; i += 8)
          *reinterpret_cast<uint4*>(dkp + i) = make_uint4(
              pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
              pack_bf16x2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])),
              pack_bf16x2(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])),
              pack_bf16x2(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
      }
      if (p.dk_colsum) attn_bwd_colsum16(v, kj < p.Tk, p.dk_colsum + h * 64 + 16 * cg, lane);
      tmem_ld_32x32b_x16(tDV + lane_addr + 16 * cg, v);
      tmem_ld_wait();
      if (kj < p.Tk) {
        __nv_bfloat16* dvp = p.dv + (dk_row0 + kj) * p.ld_dv + p.dv_col0 + h * 64 + 16 * cg;
#pragma unroll
        for (int i = 0; i < 16; i += 8)
          *reinterpret_cast<uint4*>(dvp + i) = make_uint4(
              pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
              pack_bf16x2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])),
              pack_bf16x2(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])),
              pack_bf16x2(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
      }
      if (p.dv_colsum) attn_bwd_colsum16(v, kj < p.Tk, p.dv_colsum + h * 64 + 16 * cg, lane);
    }
  } else if (warp >= 4 && !active) {
    // keys that no query sees (or beyond kv_len): zero gradients
    const int r = (warp & 3) * 32 + lane, cg = (warp - 4) >> 2;
    const int kj = k0 + r;
    if (kj < p.Tk) {
      __nv_bfloat16* dkp = p.dk + (dk_row0 + kj) * p.ld_dk + p.dk_col0 + h * 64 + 16 * cg;
      __nv_bfloat16* dvp = p.dv + (dk_row0 + kj) * p.ld_dv + p.dv_col0 + h * 64 + 16 * cg;
#pragma unroll
      for (int i = 0; i < 16; i += 8) {
        *reinterpret_cast<uint4*>(dkp + i) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(dvp + i) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// delta[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d]   (one warp per row of the [B*Tq, nh*64] matrices)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ dout, int64_t ld_do,
                  const __nv_bfloat16* __restrict__ out, int64_t ld_o,
                  const float* __restrict__ out_f32, float* __restrict__ delta,
                  int rows, int Tq, int nh, const int* __restrict__ cu_q, const int* __restrict__ row_b,
                  const int* __restrict__ n_rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  int b = row / Tq, q = row - b * Tq;
  if (cu_q) {   // packed batch: row is a packed row; delta keeps the padded [B, nh, T] indexing
    if (row >= *n_rows) return;
    b = row_b[row];
    q = row - cu_q[b];
  }
  // lanes 0..15 cover one head (16 lanes x 4 elements), a warp covers 2 heads per step
  for (int h0 = 0; h0 < nh; h0 += 2) {
    const int h = h0 + (lane >> 4);
    float s = 0.f;
    if (h < nh) {
      const int col = h * 64 + (lane & 15) * 4;
      const uint2 a = *reinterpret_cast<const uint2*>(dout + (int64_t)row * ld_do + col);
      const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y);
      if (out_f32) {  // un-rounded forward output: keeps sum_j dS_ij = 0 to fp32 accuracy
        const float4 o = *reinterpret_cast<const float4*>(out_f32 + (int64_t)row * (nh * 64) + col);
        s = a0.x * o.x + a0.y * o.y + a1.x * o.z + a1.y * o.w;
      } else {
        const uint2 o = *reinterpret_cast<const uint2*>(out + (int64_t)row * ld_o + col);
        const float2 o0 = unpack_bf16x2(o.x), o1 = unpack_bf16x2(o.y);
        s = a0.x * o0.x + a0.y * o0.y + a1.x * o1.x + a1.y * o1.y;
      }
    }
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((lane & 15) == 0 && h < nh) delta[((int64_t)b * nh + h) * Tq + q] = s;
  }
}

}  // namespace ergm

using namespace ergm;

extern "C" int ergm_attn_bwd(const void* q, int64_t ld_q, int q_col0, const void* k, int64_t ld_k,
                             int k_col0, const void* v, int64_t ld_v, int v_col0, const void* out,
                             int64_t ld_out, const float* out_f32, const void* dout, int64_t ld_do, const float* lse,
                             float* delta, float* dq_accum, int64_t ld_dq, void* dk, int64_t ld_dk,
                             int dk_col0, void* dv, int64_t ld_dv, int dv_col0, float* dk_colsum, float* dv_colsum,
                             const int* kv_lens, int B, int nh, int Tq, int Tk, int head_dim, int causal, int causal_off,
                             float dropout_p, uint64_t seed, uint64_t offset, const ergm_pack* pack, int pack_kv,
                             void* stream) {
  if (!q || !k || !v || !out || !dout || !lse || !delta || !dq_accum || !dk || !dv) return ERGM_ERR_ARG;
  if (pack && (!pack->cu_rows || !pack->row_b || !pack->n_rows || (pack_kv && !pack->kv_lens))) return ERGM_ERR_ARG;
  const uint64_t q_rows = pack ? (uint64_t)B * Tq : (uint64_t)Tq, q_bat = pack ? 1 : (uint64_t)B;
  const uint64_t k_rows = (pack && pack_kv) ? (uint64_t)B * Tk : (uint64_t)Tk, k_bat = (pack && pack_kv) ? 1 : (uint64_t)B;
  if (B <= 0 || nh <= 0 || Tq <= 0 || Tk <= 0) return ERGM_ERR_ARG;
  if (head_dim != 64) return ERGM_ERR_UNSUPPORTED;
  if (ld_q % 8 || ld_k % 8 || ld_v % 8 || ld_out % 8 || ld_do % 8 || ld_dq % 4 || ld_dk % 8 || ld_dv % 8 ||
      q_col0 % 8 || k_col0 % 8 || v_col0 % 8 || dk_col0 % 8 || dv_col0 % 8)
    return ERGM_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  attn_delta_kernel<<<(B * Tq + 7) / 8, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(dout), ld_do,
                                                    reinterpret_cast<const __nv_bfloat16*>(out), ld_out,
                                                    out_f32, delta, B * Tq, Tq, nh, pack ? pack->cu_rows : nullptr,
                                                    pack ? pack->row_b : nullptr, pack ? pack->n_rows : nullptr);
  CUtensorMap tq, tk, tv, tdo;
  int rc;
  if ((rc = encode_tmap_3d(&tq, q, 2, (uint64_t)(q_col0 + nh * 64), q_rows, q_bat,
                           (uint64_t)ld_q * 2, q_rows * ld_q * 2, 64, 128, 1))) return rc;
  if ((rc = encode_tmap_3d(&tk, k, 2, (uint64_t)(k_col0 + nh * 64), k_rows, k_bat,
                           (uint64_t)ld_k * 2, k_rows * ld_k * 2, 64, 128, 1))) return rc;
  if ((rc = encode_tmap_3d(&tv, v, 2, (uint64_t)(v_col0 + nh * 64), k_rows, k_bat,
                           (uint64_t)ld_v * 2, k_rows * ld_v * 2, 64, 128, 1))) return rc;
  if ((rc = encode_tmap_3d(&tdo, dout, 2, (uint64_t)(nh * 64), q_rows, q_bat,
                           (uint64_t)ld_do * 2, q_rows * ld_do * 2, 64, 128, 1))) return rc;
  AttnBwdParams p;
  p.lse = lse; p.delta = delta; p.dq_accum = dq_accum;
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.dk_colsum = dk_colsum; p.dv_colsum = dv_colsum;
  p.kv_lens = kv_lens;
  p.cu_q = pack ? pack->cu_rows : nullptr;
  p.cu_k = (pack && pack_kv) ? pack->cu_rows : nullptr;
  if (pack && pack_kv) p.kv_lens = pack->kv_lens;
  p.ld_dq = ld_dq; p.ld_dk = ld_dk; p.ld_dv = ld_dv; p.dk_col0 = dk_col0; p.dv_col0 = dv_col0;
  p.Tq = Tq; p.Tk = Tk; p.nh = nh;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.causal_off = causal_off;
  p.scale = 1.0f / sqrtf((float)head_dim);
  p.drop = make_site(seed, offset, dropout_p, (uint32_t)Tk);
  p.do_drop = dropout_p > 0.f;
  ERGM_SET_SMEM_ATTR(attn_bwd_kernel<true>, AB_SMEM);
  ERGM_SET_SMEM_ATTR(attn_bwd_kernel<false>, AB_SMEM);
  dim3 grid(B, nh, (Tk + 127) / 128);
  if (causal)
    attn_bwd_kernel<true><<<grid, AB_THREADS, AB_SMEM, s>>>(tq, tk, tv, tdo, p);
  else
    attn_bwd_kernel<false><<<grid, AB_THREADS, AB_SMEM, s>>>(tq, tk, tv, tdo, p);
  return (int)cudaGetLastError();
}

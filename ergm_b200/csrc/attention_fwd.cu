// ergm_attn_fwd — fused flash-style attention forward on tcgen05 / TMEM for head_dim 64.
//
// Replaces GPT2Attention._attn (/root/reference/src/model.py:119-148): QK^T, /sqrt(hd),
// causal where-mask (self) or none (cross), softmax in fp32, attn_dropout, PV, and the
// _split_heads / _merge_heads permutes (:190-198) — Q/K/V are read straight out of the
// [rows, ld] projection outputs with 3-D TMA maps (col, row-in-sequence, batch) and the
// context is written merged as [B*Tq, nh*64].
//
// Persistent kernel: TWO CTAs are resident per SM (97 KB smem, 256 TMEM columns, <= 80 registers x 384 threads each),
// each walks a static list of (batch, head, 128-query block) work items, so that one CTA's exp phase overlaps the
// other's load / MMA / barrier latencies.  Warp roles: 0 = TMA producer (runs ahead across items), 1 = MMA issuer,
// 2 = TMEM allocator, 3 = output TMA store, 4..11 = softmax.  Two softmax threads share a query row (TMEM lane): each owns 64 of
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
// Q x2 | K x2 | V x2 | xch 1 KB | barriers  (the dynamic smem base is 1 KB aligned: no static smem in this kernel)
constexpr int AT_SMEM = 6 * AT_TILE + 1024 + 128;

struct AttnFwdParams {
  __nv_bfloat16* out;  // [B*Tq, ld_out], head h at columns [h*64, h*64+64)
  float* lse;          // [B, nh, Tq]
  float* out_f32;      // nullable [B*Tq, nh*64]: un-rounded context, used by backward's delta
  const int* kv_lens;  // nullable [B]: keys >= kv_lens[b] are masked
  const int* cu_q;     // nullable [B+1]: packed batch, queries (and outputs) of sample b are rows cu_q[b] .. cu_q[b+1]
  const int* cu_k;     // nullable [B+1]: keys / values packed the same way (self attention); NULL: rows b*Tk ..
  int64_t ld_out;
  int B, Tq, Tk, nh;
  int q_col0, k_col0, v_col0;
  int causal_off;      // query i may see key j iff j <= i + causal_off
  float scale;         // 1/sqrt(hd)
  DropoutSite drop;
  int do_drop;
#ifdef ERGM_ATTN_TRACE
  long long* trace;    // [ctas][64]: (event, %globaltimer) pairs of softmax thread 0 (profiling builds only)
#endif
};
#ifdef ERGM_ATTN_TRACE
long long* g_attn_fwd_trace = nullptr;
#endif

// One work item = one (batch, head, 128-query block).
struct AfItem {
  int b, h, q0;
  int q_row0, q_bat, k_row0, k_bat;   // TMA row base / batch coordinate (packed batches: one long row sequence)
  int64_t out_row0, stat_row0;
  int Tq, Tk, kv_len, n_kv;
  bool exists;
};

template <bool CAUSAL>
ERGM_DEVINL AfItem af_item(const AttnFwdParams& p, int w) {
  AfItem t;
  // heaviest items first: causal query block qb visits qb + 1 key blocks
  const int n_bh = p.B * p.nh;
  const int n_qb = (p.Tq + 127) / 128;
  const int qb = n_qb - 1 - w / n_bh, rem = w % n_bh;
  t.h = rem / p.B;
  t.b = rem - t.h * p.B;
  t.q0 = qb * 128;
  t.Tq = p.Tq; t.Tk = p.Tk;
  t.q_row0 = 0; t.q_bat = t.b; t.k_row0 = 0; t.k_bat = t.b;
  t.out_row0 = (int64_t)t.b * p.Tq;
  if (p.cu_q) { t.q_row0 = p.cu_q[t.b]; t.q_bat = 0; t.out_row0 = t.q_row0; t.Tq = p.cu_q[t.b + 1] - t.q_row0; }
  if (p.cu_k) { t.k_row0 = p.cu_k[t.b]; t.k_bat = 0; t.Tk = p.cu_k[t.b + 1] - t.k_row0; }
  t.exists = t.q0 < t.Tq;   // packed batches: this sample may have no query rows in this block
  t.stat_row0 = ((int64_t)t.b * p.nh + t.h) * p.Tq;   // lse / dropout rows keep the padded [B, nh, T] indexing
  t.kv_len = t.Tk;
  if (p.kv_lens) t.kv_len = min(t.kv_len, p.kv_lens[t.b]);
  t.n_kv = max(1, (t.kv_len + 127) / 128);
  if (CAUSAL) {
    const int last_key = min(t.q0 + 127, t.Tq - 1) + p.causal_off;  // largest visible key
    t.n_kv = max(1, min(t.n_kv, max(0, last_key) / 128 + 1));
  }
  return t;
}

// Persistent kernel, two CTAs per SM, each walking a static list of work items (snake order over the heaviest-first
// item list).  The first version launched one CTA per item (768 at the benchmarked shape): 8 us of CTA lifetime for
// ~2 us of MMA + softmax work, the rest TMEM allocation, the first TMA round trip, scattered 16-byte output stores,
// de-allocation and CTA turn-over.  Here the producer runs ahead across items (Q double-buffered, the K / V ring
// never drains), TMEM is allocated once, S of the next item is issued while the softmax warps still normalise and
// store the previous one, and the bf16 output leaves through the finished item's Q buffer and one TMA store issued by
// an otherwise idle warp.
template <bool CAUSAL>
__global__ void __launch_bounds__(AT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_out,
                const AttnFwdParams p_in) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  AttnFwdParams p = p_in;
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t sQ = base, sK = base + 2 * AT_TILE, sV = base + 4 * AT_TILE;   // Q x2 | K x2 | V x2
  const uint32_t sX = base + 6 * AT_TILE;  // float xch[2 halves][128]
  const uint32_t bars = sX + 1024;
  const uint32_t bar_s = bars, bar_p = bars + 8, bar_o = bars + 16;
  auto kv_full = [&](int s) { return bars + 32 + 8u * s; };
  auto kv_empty = [&](int s) { return bars + 48 + 8u * s; };
  auto q_full = [&](int s) { return bars + 64 + 8u * s; };
  auto q_empty = [&](int s) { return bars + 80 + 8u * s; };
  auto out_staged = [&](int s) { return bars + 96 + 8u * s; };
  const uint32_t tmem_slot = bars + 112;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef ERGM_ATTN_TRACE
  int n_log = 0;
#define AF_LOG(ev)                                                                     \
  do {                                                                                 \
    if (p.trace && threadIdx.x == 128 && n_log < 31) {                                 \
      long long t_;                                                                    \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                           \
      p.trace[blockIdx.x * 64 + 2 * n_log] = (ev);                                     \
      p.trace[blockIdx.x * 64 + 2 * n_log + 1] = t_;                                   \
      ++n_log;                                                                         \
    }                                                                                  \
  } while (0)
#else
#define AF_LOG(ev) do {} while (0)
#endif
  const int n_items = p.B * p.nh * ((p.Tq + 127) / 128);
  const int G = gridDim.x, cta = blockIdx.x;
  auto item_of = [&](int r) { return r * G + ((r & 1) ? G - 1 - cta : cta); };
  const int n_rounds = (n_items + G - 1) / G;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1);
      mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1);
      mbar_init(out_staged(s), 8);
    }
    mbar_init(bar_s, 1); mbar_init(bar_p, 8); mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();               // programmatic dependent launch: no global memory access above this line
  pdl_launch_dependents();
  p.drop = p_in.drop.resolved();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tS = tmem, tP = tmem + 128, tO = tmem + 192;

  if (warp == 0) {
    // ---------------- producer ----------------
    if (lane == 0) {
      int n = 0, gb = 0;
      for (int r = 0; r < n_rounds; ++r) {
        const int w = item_of(r);
        if (w >= n_items) continue;
        const AfItem t = af_item<CAUSAL>(p, w);
        if (!t.exists) continue;
        const int qs = n & 1;
        mbar_wait(q_empty(qs), ((n >> 1) & 1) ^ 1);   // the item before last has left this buffer (its output too)
        mbar_expect_tx(q_full(qs), AT_TILE);
        tma_load_3d(sQ + qs * AT_TILE, &tm_q, q_full(qs), p.q_col0 + t.h * AT_D, t.q_row0 + t.q0, t.q_bat);
        for (int j = 0; j < t.n_kv; ++j, ++gb) {
          const int st = gb & 1;
          mbar_wait(kv_empty(st), ((gb >> 1) & 1) ^ 1);
          mbar_expect_tx(kv_full(st), 2 * AT_TILE);
          tma_load_3d(sK + st * AT_TILE, &tm_k, kv_full(st), p.k_col0 + t.h * AT_D, t.k_row0 + j * 128, t.k_bat);
          tma_load_3d(sV + st * AT_TILE, &tm_v, kv_full(st), p.v_col0 + t.h * AT_D, t.k_row0 + j * 128, t.k_bat);
        }
        ++n;
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
      int n = 0, gb = 0;
      bool s_issued = false;   // S of block gb is already in flight (issued behind the previous block's PV)
      auto issue_s = [&](int qs, int g) {
        const uint32_t q_t = sQ + qs * AT_TILE, k_t = sK + (g & 1) * AT_TILE;
#pragma unroll
        for (int ks = 0; ks < AT_D / 16; ++ks)
          umma_ss(tS, make_smem_desc_sw128(q_t + ks * 32, 16, 1024),
                  make_smem_desc_sw128(k_t + ks * 32, 16, 1024), idesc_s, ks > 0);
        umma_commit(bar_s);
      };
      for (int r = 0; r < n_rounds; ++r) {
        const int w = item_of(r);
        if (w >= n_items) continue;
        const AfItem t = af_item<CAUSAL>(p, w);
        if (!t.exists) continue;
        const int qs = n & 1;
        if (!s_issued) {
          mbar_wait(q_full(qs), (n >> 1) & 1);
          mbar_wait(kv_full(gb & 1), (gb >> 1) & 1);
          tc_fence_after();
          issue_s(qs, gb);
        }
        s_issued = false;
        for (int j = 0; j < t.n_kv; ++j, ++gb) {
          const int st = gb & 1;
          const uint32_t v_t = sV + st * AT_TILE;
          mbar_wait(bar_p, gb & 1);  // P in TMEM, S consumed
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)   // A = P (TMEM, 8 columns = 16 bf16 per K step)
            umma_ts(tO, tP + ks * 8, make_smem_desc_sw128(v_t + ks * 2048, 8192, 1024), idesc_o, ks > 0);
          umma_commit(bar_o);
          umma_commit(kv_empty(st));
          if (j + 1 < t.n_kv) {
            mbar_wait(kv_full((gb + 1) & 1), ((gb + 1) >> 1) & 1);
            tc_fence_after();
            issue_s(qs, gb + 1);
          }
        }
        // S of the NEXT item's first key block, behind this item's last PV: its softmax starts as soon as the warps
        // have normalised and stored this item
        for (int r2 = r + 1; r2 < n_rounds; ++r2) {
          const int w2 = item_of(r2);
          if (w2 >= n_items) continue;
          const AfItem t2 = af_item<CAUSAL>(p, w2);
          if (!t2.exists) continue;
          mbar_wait(q_full((n + 1) & 1), ((n + 1) >> 1) & 1);
          mbar_wait(kv_full(gb & 1), (gb >> 1) & 1);
          tc_fence_after();
          issue_s((n + 1) & 1, gb);
          s_issued = true;
          break;
        }
        ++n;
      }
    }
  } else if (warp == 3) {
    // ---------------- output warp: TMA store of the staged bf16 tile, then the Q buffer goes back to the producer ----
    if (lane == 0) {
      int n = 0;
      for (int r = 0; r < n_rounds; ++r) {
        const int w = item_of(r);
        if (w >= n_items) continue;
        const AfItem t = af_item<CAUSAL>(p, w);
        if (!t.exists) continue;
        const int qs = n & 1;
        mbar_wait(out_staged(qs), (n >> 1) & 1);
        if (t.q0 + 128 <= t.Tq) {   // whole 128-row box inside the sample: staged in smem (else the threads stored it)
          tma_store_3d(&tm_out, sQ + qs * AT_TILE, t.h * AT_D, (int)t.q_row0 + t.q0, t.q_bat);
          tma_store_commit();
          tma_store_wait_read();
        }
        mbar_arrive(q_empty(qs));
        ++n;
      }
      tma_store_wait_all();   // smem must outlive the bulk stores
    }
  } else if (warp >= 4) {
    const int hf = (warp - 4) >> 2;          // which half of the key columns / output columns
    const int r = (warp & 3) * 32 + lane;    // query row inside the block == TMEM lane
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tS_mine = tS + lane_addr + 64 * hf;
    const uint32_t tP_mine = tP + lane_addr + 32 * hf;
    const uint32_t tO_mine = tO + lane_addr + 32 * hf;
    const float c = p.scale * 1.4426950408889634f;  // exp(x*scale) = exp2(x*c)
    const uint32_t thr_hi = p.drop.thr16() << 16;   // 16-bit lane >= thr16  <=>  (lane << 16) >= thr_hi
    const float keep_scale = p.do_drop ? p.drop.keep_scale() : 1.f;
    const uint32_t xch_mine = sX + (hf * 128 + r) * 4;
    const uint32_t xch_other = sX + ((hf ^ 1) * 128 + r) * 4;
    int n = 0, gb = 0;
    for (int rr = 0; rr < n_rounds; ++rr) {
      const int w = item_of(rr);
      if (w >= n_items) continue;
      const AfItem t = af_item<CAUSAL>(p, w);
      if (!t.exists) continue;
      const int h = t.h;
      const int qi = t.q0 + r;
      const uint32_t drop_row = (uint32_t)(t.stat_row0 + qi);
      const int vis = CAUSAL ? min(t.kv_len - 1, qi + p.causal_off) : t.kv_len - 1;  // last visible key
      float m = -INFINITY, l = 0.f;
      float o[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = 0.f;
      for (int j = 0; j < t.n_kv; ++j, ++gb) {
      const int k0 = j * 128 + 64 * hf;  // first key column this thread owns in block j
      // warp-uniform: does any row of this warp need masking in this block?
      const bool need_mask = (j * 128 + 127 > t.kv_len - 1) ||
                             (CAUSAL && (j * 128 + 127 > t.q0 + (warp & 3) * 32 + p.causal_off));
      AF_LOG(1);
      mbar_wait(bar_s, gb & 1);
      tc_fence_after();
      AF_LOG(2);
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
      AF_LOG(3);
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch_mine), "f"(mx) : "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      AF_LOG(4);
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
            // (the 1 / (1 - p) rescale of the kept probabilities is applied once per row, on the normalised output)
            pr[i] = ((hsh << 16) >= thr_hi) ? pr[i] : 0.f;
            pr[i + 1] = (hsh >= thr_hi) ? pr[i + 1] : 0.f;
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
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      AF_LOG(5);
      mbar_wait(bar_o, gb & 1);
      AF_LOG(6);
      tc_fence_after();
      {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tO_mine, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = fmaf(o[i], alpha, __uint_as_float(v[i]));
      }
      tc_fence_before();
      AF_LOG(7);
      }
      // combine the two half-row sums
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch_mine), "f"(l) : "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float lo;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(lo) : "r"(xch_other) : "memory");
      l += lo;
      const float inv = l > 0.f ? keep_scale / l : 0.f;   // l sums the un-dropped probabilities; kept ones carry 1 / (1 - p)
      const bool full = t.q0 + 128 <= t.Tq;   // CTA-uniform
      const int qs = n & 1;
      if (full) {
        // bf16 tile -> the finished item's Q buffer in the TMA (128B-swizzled) layout: 4 conflict-free 16-byte stores
        const uint32_t row = sQ + qs * AT_TILE + (uint32_t)r * 128u;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          const uint32_t ch = (uint32_t)(4 * hf + (i >> 3)) ^ (uint32_t)(r & 7);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (ch << 4)),
                       "r"(pack_bf16x2(o[i] * inv, o[i + 1] * inv)), "r"(pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv)),
                       "r"(pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv)), "r"(pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv))
                       : "memory");
        }
        fence_proxy_async();
      } else if (qi < t.Tq) {
        __nv_bfloat16* op = p.out + (t.out_row0 + qi) * p.ld_out + h * AT_D + 32 * hf;
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          const uint4 u = make_uint4(pack_bf16x2(o[i] * inv, o[i + 1] * inv), pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv),
                                     pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv), pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv));
          *reinterpret_cast<uint4*>(op + i) = u;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(out_staged(qs));
      if (qi < t.Tq) {
        if (p.out_f32) {
          float* of = p.out_f32 + (t.out_row0 + qi) * (p.nh * AT_D) + h * AT_D + 32 * hf;
          // 256-bit stores: a warp instruction fills 32 whole sectors (the float4 version touched 32 half sectors per
          // instruction and twice as many instructions: 3.3 us per item in the LSU, profiles/r2_attn_fwd_timeline.txt)
#pragma unroll
          for (int i = 0; i < 32; i += 8)
            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(of + i), "f"(o[i] * inv),
                         "f"(o[i + 1] * inv), "f"(o[i + 2] * inv), "f"(o[i + 3] * inv), "f"(o[i + 4] * inv),
                         "f"(o[i + 5] * inv), "f"(o[i + 6] * inv), "f"(o[i + 7] * inv)
                         : "memory");
        }
        if (p.lse && hf == 0)
          p.lse[t.stat_row0 + qi] = (m == -INFINITY ? 0.f : m) * p.scale + logf(l);
      }
      AF_LOG(8);
      ++n;
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
  CUtensorMap tq, tk, tv, tout;
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
  if ((reinterpret_cast<uintptr_t>(out) & 15) ||
      (rc = encode_tmap_3d(&tout, out, 2, (uint64_t)(nh * AT_D), q_rows, q_bat, (uint64_t)ld_out * 2, q_rows * ld_out * 2,
                           AT_D, 128, 1)))
    return rc ? rc : ERGM_ERR_ARG;
  AttnFwdParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse; p.out_f32 = out_f32; p.kv_lens = kv_lens; p.ld_out = ld_out;
  p.cu_q = pack ? pack->cu_rows : nullptr;
  p.cu_k = (pack && pack_kv) ? pack->cu_rows : nullptr;
  if (pack && pack_kv) p.kv_lens = pack->kv_lens;
  p.B = B; p.Tq = Tq; p.Tk = Tk; p.nh = nh;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.causal_off = causal_off;
  p.scale = 1.0f / sqrtf((float)head_dim);
  p.drop = make_site(seed, offset, dropout_p, (uint32_t)Tk);
  p.do_drop = dropout_p > 0.f;
#ifdef ERGM_ATTN_TRACE
  p.trace = g_attn_fwd_trace;
#endif
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
  const int n_items = B * nh * ((Tq + 127) / 128);
  dim3 grid(n_items < 2 * num_sms() ? n_items : 2 * num_sms());   // persistent: two CTAs per SM
  if (causal)
    return (int)launch_pdl(attn_fwd_kernel<true>, grid, dim3(AT_THREADS), (size_t)AT_SMEM, (cudaStream_t)stream, 1, tq, tk, tv, tout, p);
  return (int)launch_pdl(attn_fwd_kernel<false>, grid, dim3(AT_THREADS), (size_t)AT_SMEM, (cudaStream_t)stream, 1, tq, tk, tv, tout, p);
}

#ifdef ERGM_ATTN_TRACE
extern "C" int ergm_attn_fwd_set_trace(long long* dev_ptr) { ergm::g_attn_fwd_trace = dev_ptr; return ERGM_OK; }
#endif

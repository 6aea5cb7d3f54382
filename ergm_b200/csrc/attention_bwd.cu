// ergm_attn_bwd — fused attention backward on tcgen05 / TMEM, head_dim 64.
//
// Gradient of GPT2Attention._attn (/root/reference/src/model.py:119-148) w.r.t. Q, K, V given
// dO, recomputing the probabilities from the saved log-sum-exp instead of storing the
// [B,nh,T,T] score / probability tensors the reference keeps for autograd.
//
// Persistent kernel: one CTA per SM walks a list of work items.  An item is a whole (batch, head) - all its 128-key
// blocks j, each against the query blocks i that can see those keys - when the sequence has at most two query blocks
// (dQ then stays in TMEM), otherwise a single key block.  Everything is computed TRANSPOSED (keys on TMEM lanes,
// queries on columns) so that P^T and dS^T come out of the element-wise warps exactly in the layout the next MMAs need:
//   S^T  = K_j Q_i^T                      SS MMA  M128(kv) N128(q) K64      -> TMEM [0,128)
//   dP^T = V_j dO_i^T                     SS MMA                            -> TMEM [128,256)
//   P^T  = exp2(S^T c - lse_i log2e),  dS^T = P^T (dP^T - delta_i) scale    (8 compute warps)
//   dV_j += P^T dO_i                      SS MMA  M128(kv) N64 K128(q)      -> TMEM [256,320)
//   dK_j += dS^T Q_i                      SS MMA                            -> TMEM [320,384)
//   dQ_i  = dS K_j                        SS MMA  M128(q)  N64 K128(kv)     -> TMEM [384,448)
// The smem tile holding dS^T ([kv rows][q contiguous], 128B swizzle) is at the same time the
// K-major A operand of the dK product and the MN-major A operand of the dQ product, and the
// TMA tiles of Q_i / dO_i / K_j serve as K-major and MN-major B operands without any copy.
// dQ / dK / dV leave through smem tiles and TMA stores (whole-head items) or, for dQ of key-block items, one TMA
// reduction into an fp32 workspace; blocks that end inside a sample are stored by the threads.
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
// 2x(K,V), 2x(Q,dO), PT (2 tiles), dST (2 tiles), dK / dV store staging (2 tiles), lse/delta (2 stages x 2 x 128
// floats), barriers; the dynamic smem base is 1 KB aligned (no static smem in this kernel): 231,680 of 232,448 bytes
constexpr int AB_SMEM = 4 * AB_TILE + 4 * AB_TILE + 2 * AB_TILE + 2 * AB_TILE + 2 * AB_TILE + 2048 + 256;

struct AttnBwdParams {
  const float* lse;     // [B, nh, Tq]
  const float* delta;   // [B, nh, Tq]
  float* dq_accum;      // fp32 [B*Tq, nh*64] scratch, zeroed: key-block items add their dQ with red.global (Tq > 256)
  __nv_bfloat16* dq;    // [B*Tq, ld_dq], head h at columns [dq_col0 + h*64, ...): written by whole-head items
  __nv_bfloat16* dk;    // [B*Tk, ld_dk], head h at columns [dk_col0 + h*64, ...)
  __nv_bfloat16* dv;
  float* dq_colsum;     // nullable [nh*64]: += column sums of dQ as stored (bias gradient of the Q projection)
  float* dk_colsum;     // nullable [nh*64]: += column sums of dK as stored (bias gradient of the K projection)
  float* dv_colsum;     // nullable [nh*64]
  const int* kv_lens;
  const int* cu_q;      // nullable [B+1]: packed batch (queries, dO, dQ rows of sample b start at cu_q[b])
  const int* cu_k;      // nullable [B+1]: keys / values / dK / dV packed the same way (self attention)
  int64_t ld_dq, ld_dk, ld_dv;
  int dq_col0, dk_col0, dv_col0;
  int B, Tq, Tk, nh;
  int q_col0, k_col0, v_col0;
  int causal_off;
  float scale;
  DropoutSite drop;
  int do_drop;
  int whole_head;       // 1: a work item is a whole (batch, head): dQ accumulates in TMEM across key blocks and is STORED
                        //    (Tq <= 256: two 64-column accumulators); 0: item = one key block, dQ added with red.global
#ifdef ERGM_ATTN_TRACE
  long long* trace;     // [ctas][16] %globaltimer stamps (profiling builds only: scripts/trace_attn_bwd.py)
#endif
};

#ifdef ERGM_ATTN_TRACE
long long* g_attn_bwd_trace = nullptr;
#endif

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

// Element-wise stage of one thread: key row r (= TMEM lane), 32 query columns [col0, col0 + 32) of the 128-query block.
//   P^T = exp2(S^T c - lse log2e);  dS^T = P^T (dP^T - delta) scale;  with dropout (model.py:142) the stored P^T and
//   dP^T carry the keep mask / keep_scale.  The stage is instruction-issue bound (profiles/r2_attn_bwd_timeline_v1.txt),
//   so everything that does not depend on the element is hoisted: the Weyl counter of the dropout hash advances by one
//   add, the 16-bit keep test is one shift + one unsigned compare, delta arrives pre-multiplied by the softmax scale,
//   and masking / dropout are compile-time variants (the first version tested both per element: ~30 instructions,
//   now ~8 without and ~21 with dropout).
struct AbElemCtx {
  uint32_t tS, tDP;       // TMEM addresses (this warp's lane quarter)
  uint32_t stat;          // smem: [lse * log2e : 128][delta * scale : 128] of this query block
  uint32_t sPT, sDS;      // smem tiles (2 x [128 keys][64 queries], 128B swizzle)
  int r;                  // key row inside the block
  float c, scale, keep_scale;
  uint32_t thr_hi;        // dropout threshold << 16
  uint32_t lsh;           // 16 - (16-bit lane of the hash this key owns)
  uint32_t hash0, hash_a; // Weyl counter of query column 0 of the block, increment per query
  int kmin;               // causal: first query that sees this key (INT_MAX: key beyond kv_len)
  bool key_ok;
  int q0, Tq;
};

template <bool CAUSAL, bool MASK, bool DROP>
ERGM_DEVINL void ab_elementwise(const AbElemCtx& e, int col0) {
  const float ks_scale = e.keep_scale * e.scale;
#pragma unroll 1
  for (int cc = col0; cc < col0 + 32; cc += 16) {
    uint32_t sv[16], dv_[16];
    tmem_ld_32x32b_x16(e.tS + cc, sv);
    tmem_ld_32x32b_x16(e.tDP + cc, dv_);
    float ls[16], dl[16];
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(ls[i]), "=f"(ls[i + 1]), "=f"(ls[i + 2]), "=f"(ls[i + 3]) : "r"(e.stat + (cc + i) * 4));
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(dl[i]), "=f"(dl[i + 1]), "=f"(dl[i + 2]), "=f"(dl[i + 3]) : "r"(e.stat + 512 + (cc + i) * 4));
    }
    uint32_t hx = e.hash0 + (uint32_t)cc * e.hash_a;
    tmem_ld_wait();
    float pt[16], ds[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float pr = ex2_fast(fminf(fmaf(__uint_as_float(sv[i]), e.c, -ls[i]), 0.f));
      if (MASK) {
        const int qi = e.q0 + cc + i;
        bool vis = e.key_ok && qi < e.Tq;
        if (CAUSAL) vis = vis && qi >= e.kmin;
        pr = vis ? pr : 0.f;
      }
      const float dp = __uint_as_float(dv_[i]);
      if (DROP) {
        const bool keep = (DropoutSite::avalanche(hx) << e.lsh) >= e.thr_hi;
        hx += e.hash_a;
        const float t = fmaf(dp, ks_scale, -dl[i]);
        ds[i] = pr * (keep ? t : -dl[i]);
        pt[i] = keep ? pr * e.keep_scale : 0.f;
      } else {
        ds[i] = pr * fmaf(dp, e.scale, -dl[i]);
        pt[i] = pr;
      }
    }
    const uint32_t rowoff = (uint32_t)(cc >> 6) * AB_TILE + (uint32_t)e.r * 128u;
#pragma unroll
    for (int i = 0; i < 16; i += 8) {
      const uint32_t piece = (uint32_t)(((cc & 63) + i) >> 3);
      const uint32_t off = rowoff + ((piece ^ (uint32_t)(e.r & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(e.sPT + off),
                   "r"(pack_bf16x2(pt[i], pt[i + 1])), "r"(pack_bf16x2(pt[i + 2], pt[i + 3])),
                   "r"(pack_bf16x2(pt[i + 4], pt[i + 5])), "r"(pack_bf16x2(pt[i + 6], pt[i + 7]))
                   : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(e.sDS + off),
                   "r"(pack_bf16x2(ds[i], ds[i + 1])), "r"(pack_bf16x2(ds[i + 2], ds[i + 3])),
                   "r"(pack_bf16x2(ds[i + 4], ds[i + 5])), "r"(pack_bf16x2(ds[i + 6], ds[i + 7]))
                   : "memory");
    }
  }
}

// One work item = one (batch, head, 128-key block); `blocks` = the query blocks that see those keys.
struct AbItem {
  int b, h, k0;
  int q_row0, q_bat, k_row0, k_bat;   // TMA row base / batch coordinate (packed batches: one long row sequence)
  int64_t dq_row0, dk_row0, stat_row0;
  int Tq, Tk, kv_len, i_min, n_iter;
  bool exists, active;                // exists: has key rows to write; active: some query sees them
};

template <bool CAUSAL>
ERGM_DEVINL AbItem ab_item(const AttnBwdParams& p, int bh, int jb) {
  AbItem t;
  t.h = bh / p.B;
  t.b = bh - t.h * p.B;
  t.k0 = jb * 128;
  t.Tq = p.Tq; t.Tk = p.Tk;
  t.q_row0 = 0; t.q_bat = t.b; t.k_row0 = 0; t.k_bat = t.b;
  t.dq_row0 = (int64_t)t.b * p.Tq; t.dk_row0 = (int64_t)t.b * p.Tk;
  if (p.cu_q) { t.q_row0 = p.cu_q[t.b]; t.q_bat = 0; t.dq_row0 = t.q_row0; t.Tq = p.cu_q[t.b + 1] - t.q_row0; }
  if (p.cu_k) { t.k_row0 = p.cu_k[t.b]; t.k_bat = 0; t.dk_row0 = t.k_row0; t.Tk = p.cu_k[t.b + 1] - t.k_row0; }
  t.exists = t.k0 < t.Tk;    // packed batches: this sample may have no key rows in this block
  t.stat_row0 = ((int64_t)t.b * p.nh + t.h) * p.Tq;   // lse / delta / dropout rows: padded [B, nh, T] indexing
  t.kv_len = t.Tk;
  if (p.kv_lens) t.kv_len = min(t.kv_len, p.kv_lens[t.b]);
  const int n_q = (t.Tq + 127) / 128;
  t.i_min = 0;
  if (CAUSAL) t.i_min = max(0, t.k0 - p.causal_off) / 128;
  t.active = t.exists && (t.k0 < t.kv_len) && (t.i_min < n_q);
  t.n_iter = t.active ? n_q - t.i_min : 0;
  return t;
}

// Persistent kernel: one CTA per SM walks a static, weight-balanced list of work items (snake order over the
// heaviest-first item list).  The first version launched one CTA per item: 16 us of CTA lifetime for ~3 us of MMA +
// element-wise work - TMEM allocation, the first TMA round trip, the dQ / dK / dV drains, de-allocation and CTA
// turn-over were all exposed (profiles/r2_attn_bwd_timeline_v1.txt).  Here
//   * the producer warp runs ahead: K / V of the next item (double-buffered) and Q / dO + lse / delta of the next
//     query block (two stages) are in flight while the current block is computed;
//   * TMEM is allocated once; dQ is double-buffered (64 spare columns), so its drain (tcgen05.ld + red.global.add)
//     happens AFTER the next block's element-wise stage has been handed to the MMA warp, off the critical path;
//   * dK / dV of an item are drained while the MMA warp already computes S / dP of the next item.
template <bool CAUSAL>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_dq, const __grid_constant__ CUtensorMap tm_dk,
                const __grid_constant__ CUtensorMap tm_dv, const __grid_constant__ CUtensorMap tm_ws,
                const AttnBwdParams p_in) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  AttnBwdParams p = p_in;
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t sK = base, sV = base + 2 * AB_TILE;   // 2 item buffers each
  const uint32_t sQ = base + 4 * AB_TILE;    // 2 stages
  const uint32_t sDO = base + 6 * AB_TILE;   // 2 stages
  const uint32_t sPT = base + 8 * AB_TILE;   // 2 tiles (q chunks)
  const uint32_t sDS = base + 10 * AB_TILE;  // 2 tiles
  const uint32_t sOut = base + 12 * AB_TILE; // dK, dV tiles on their way to global memory (TMA store)
  const uint32_t sStat = base + 14 * AB_TILE;  // [2 stages][lse * log2e : 128 | delta * scale : 128] floats
  const uint32_t bars = sStat + 2048;
  const uint32_t bar_sdp = bars, bar_pds = bars + 8, bar_dkv_free = bars + 16, bar_item = bars + 24;
  auto qdo_full = [&](int s) { return bars + 32 + 8u * s; };
  auto qdo_empty = [&](int s) { return bars + 48 + 8u * s; };
  auto kv_full = [&](int s) { return bars + 64 + 8u * s; };
  auto kv_empty = [&](int s) { return bars + 80 + 8u * s; };
  const uint32_t tmem_slot = bars + 96;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_bh = p.B * p.nh;
  const int n_jb = (p.Tk + 127) / 128;
  const bool whole = p.whole_head != 0;
  // work items: whole heads, or single key blocks heaviest first (causal key block jb is visited by the query blocks
  // >= jb, so key block 0 has the most work)
  const int n_items = whole ? n_bh : n_bh * n_jb;
  const int G = gridDim.x, cta = blockIdx.x;
  // snake order: round r walks the CTAs forwards (even r) or backwards (odd r), so that the CTAs that got the
  // heaviest items of one round get the lightest of the next
  auto item_of = [&](int r) { return r * G + ((r & 1) ? G - 1 - cta : cta); };
  const int n_rounds = (n_items + G - 1) / G;
#ifdef ERGM_ATTN_TRACE
  int n_log = 0;
#define AB_LOG(ev)                                                                     \
  do {                                                                                 \
    if (p.trace && n_log < 30) {                                                       \
      long long t_;                                                                    \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                           \
      p.trace[cta * 64 + 2 * n_log] = (ev);                                            \
      p.trace[cta * 64 + 2 * n_log + 1] = t_;                                          \
      ++n_log;                                                                         \
    }                                                                                  \
  } while (0)
#define AB_MARK(slot)                                                                  \
  do {                                                                                 \
    if (p.trace) {                                                                     \
      long long t_;                                                                    \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                           \
      p.trace[cta * 64 + (slot)] = t_;                                                 \
    }                                                                                  \
  } while (0)
  if (threadIdx.x == 0) {
    AB_MARK(60);
    if (p.trace) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); p.trace[cta * 64 + 63] = sm; }
  }
#else
#define AB_LOG(ev) do {} while (0)
#define AB_MARK(slot) do {} while (0)
#endif

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(qdo_full(s), 1); mbar_init(qdo_empty(s), 1);
      mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1);
    }
    mbar_init(bar_sdp, 1); mbar_init(bar_pds, AB_CWARPS); mbar_init(bar_dkv_free, AB_CWARPS); mbar_init(bar_item, 1);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();               // programmatic dependent launch: no global memory access above this line
  pdl_launch_dependents();
  p.drop = p_in.drop.resolved();
  if (threadIdx.x == 0) AB_MARK(61);
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));
  const uint32_t tS = tmem, tDP = tmem + 128, tDV = tmem + 256, tDK = tmem + 320, tDQ = tmem + 384;  // dQ: 2 x 64

  if (warp == 0) {
    // ---------------- producer: TMA loads + lse / delta staging, runs ahead of the MMA warp ----------------
    int n_act = 0, g = 0;
    for (int r = 0; r < n_rounds; ++r) {
      const int w = item_of(r);
      if (w >= n_items) continue;
      const int jb_lo = whole ? 0 : w / n_bh, jb_hi = whole ? n_jb : jb_lo + 1, bh = whole ? w : w - jb_lo * n_bh;
      for (int jb = jb_lo; jb < jb_hi; ++jb) {
      const AbItem t = ab_item<CAUSAL>(p, bh, jb);
      if (!t.active) continue;
      const int kb = n_act & 1;
      if (lane == 0) {
        mbar_wait(kv_empty(kb), ((n_act >> 1) & 1) ^ 1);
        mbar_expect_tx(kv_full(kb), 2 * AB_TILE);
        tma_load_3d(sK + kb * AB_TILE, &tm_k, kv_full(kb), p.k_col0 + t.h * 64, t.k_row0 + t.k0, t.k_bat);
        tma_load_3d(sV + kb * AB_TILE, &tm_v, kv_full(kb), p.v_col0 + t.h * 64, t.k_row0 + t.k0, t.k_bat);
      }
      for (int it = 0; it < t.n_iter; ++it, ++g) {
        const int st = g & 1;
        const int q0 = (t.i_min + it) * 128;
        // lse (pre-multiplied by log2 e) and delta (pre-multiplied by the softmax scale) of this query block
        float sv_[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int e = j * 32 + lane;           // 0..127: lse, 128..255: delta
          const int qi = q0 + (e & 127);
          float val = 0.f;
          if (qi < t.Tq) val = e < 128 ? __ldg(p.lse + t.stat_row0 + qi) * 1.4426950408889634f
                                       : __ldg(p.delta + t.stat_row0 + qi) * p.scale;
          sv_[j] = val;
        }
        if (lane == 0) mbar_wait(qdo_empty(st), ((g >> 1) & 1) ^ 1);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(sStat + st * 1024 + (j * 32 + lane) * 4), "f"(sv_[j]) : "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_expect_tx(qdo_full(st), 2 * AB_TILE);
          tma_load_3d(sQ + st * AB_TILE, &tm_q, qdo_full(st), p.q_col0 + t.h * 64, t.q_row0 + q0, t.q_bat);
          tma_load_3d(sDO + st * AB_TILE, &tm_do, qdo_full(st), t.h * 64, t.q_row0 + q0, t.q_bat);
        }
      }
      ++n_act;
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t id_st = make_idesc_bf16(128, 128, 0, 0);   // K-major x K-major
      const uint32_t id_dkv = make_idesc_bf16(128, 64, 0, 1);   // A K-major, B MN-major
      const uint32_t id_dq = make_idesc_bf16(128, 64, 1, 1);    // A MN-major, B MN-major
      int n_act = 0, g = 0;
      for (int r = 0; r < n_rounds; ++r) {
        const int w = item_of(r);
        if (w >= n_items) continue;
        const int jb_lo = whole ? 0 : w / n_bh, jb_hi = whole ? n_jb : jb_lo + 1, bh = whole ? w : w - jb_lo * n_bh;
        uint32_t dq_seen = 0;   // whole-head items: query blocks whose dQ accumulator already holds a contribution
        for (int jb = jb_lo; jb < jb_hi; ++jb) {
        const AbItem t = ab_item<CAUSAL>(p, bh, jb);
        if (!t.active) continue;
        const int kb = n_act & 1;
        const uint32_t k_t = sK + kb * AB_TILE, v_t = sV + kb * AB_TILE;
        mbar_wait(kv_full(kb), (n_act >> 1) & 1);
        for (int it = 0; it < t.n_iter; ++it, ++g) {
          const int st = g & 1;
          const uint32_t q_t = sQ + st * AB_TILE, do_t = sDO + st * AB_TILE;
          mbar_wait(qdo_full(st), (g >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ss(tS, make_smem_desc_sw128(k_t + ks * 32, 16, 1024),
                    make_smem_desc_sw128(q_t + ks * 32, 16, 1024), id_st, ks > 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ss(tDP, make_smem_desc_sw128(v_t + ks * 32, 16, 1024),
                    make_smem_desc_sw128(do_t + ks * 32, 16, 1024), id_st, ks > 0);
          umma_commit(bar_sdp);
          mbar_wait(bar_pds, g & 1);
          if (it == 0 && n_act > 0) mbar_wait(bar_dkv_free, (n_act - 1) & 1);  // dK / dV of the previous item drained
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
          const int ib = t.i_min + it;
          const uint32_t dq_t = tDQ + 64 * (whole ? ib : (g & 1));
          const bool dq_acc = whole && ((dq_seen >> ib) & 1u);
          dq_seen |= 1u << ib;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_ss(dq_t, make_smem_desc_sw128(sDS + ks * 2048, AB_TILE, 1024),
                    make_smem_desc_sw128(k_t + ks * 2048, 8192, 1024), id_dq, dq_acc || ks > 0);
          umma_commit(qdo_empty(st));
          if (it == t.n_iter - 1) { umma_commit(kv_empty(kb)); umma_commit(bar_item); }
        }
        ++n_act;
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------- 16 element-wise / drain warps ----------------
    const int qr = warp & 3;             // TMEM lane quarter
    const int cg = (warp - 4) >> 2;      // column group
    const int r = qr * 32 + lane;        // key row inside the block == TMEM lane (query row for dQ)
    const uint32_t lane_addr = (uint32_t)(qr * 32) << 16;
    const float c = p.scale * 1.4426950408889634f;
    const uint32_t thr16 = p.drop.thr16();
    const float keep_scale = p.do_drop ? 65536.f / (65536.f - (float)thr16) : 1.f;
    const uint32_t drop_a = p.drop.ncol4 * 2u * 0x9E3779B1u;
    const int ct = threadIdx.x - 128;
    // dQ of one query block: TMEM lanes are queries; this warp handles rows qr*32.., columns 16*cg..+16
    auto drain_dq = [&](const AbItem& t, int q0, int buf) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(tDQ + 64 * buf + lane_addr + 16 * cg, v);
      tmem_ld_wait();
      const int qi = q0 + r;
      if (qi < t.Tq) {
        float* dst = p.dq_accum + (t.dq_row0 + qi) * (int64_t)(p.nh * 64) + t.h * 64 + 16 * cg;
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i),
                       "f"(__uint_as_float(v[i])), "f"(__uint_as_float(v[i + 1])),
                       "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3]))
                       : "memory");
      }
    };
    // 16 fp32 values of row r -> bf16: either into a [128 rows][64 bf16] smem tile in the TMA (128B-swizzled) layout, or,
    // for a block that ends inside the tensor (TMA would write rows of the next sample of a packed batch), guarded
    // 32-byte stores.  The smem route costs two conflict-free st.shared per thread; the direct route is what the first
    // version always did: 32 scattered half-sectors per warp instruction, 1.5 - 4 us per item in the LSU.
    auto put_rows = [&](const uint32_t (&v)[16], bool full, uint32_t tile, __nv_bfloat16* gdst, bool row_ok) {
      uint4 a = make_uint4(pack_bf16x2(__uint_as_float(v[0]), __uint_as_float(v[1])),
                           pack_bf16x2(__uint_as_float(v[2]), __uint_as_float(v[3])),
                           pack_bf16x2(__uint_as_float(v[4]), __uint_as_float(v[5])),
                           pack_bf16x2(__uint_as_float(v[6]), __uint_as_float(v[7])));
      uint4 b = make_uint4(pack_bf16x2(__uint_as_float(v[8]), __uint_as_float(v[9])),
                           pack_bf16x2(__uint_as_float(v[10]), __uint_as_float(v[11])),
                           pack_bf16x2(__uint_as_float(v[12]), __uint_as_float(v[13])),
                           pack_bf16x2(__uint_as_float(v[14]), __uint_as_float(v[15])));
      if (full) {   // CTA-uniform
        const uint32_t row = tile + (uint32_t)r * 128u;
        const uint32_t c0 = (uint32_t)(2 * cg) ^ (uint32_t)(r & 7), c1 = (uint32_t)(2 * cg + 1) ^ (uint32_t)(r & 7);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (c0 << 4)), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (c1 << 4)), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
      } else if (row_ok) {
        *reinterpret_cast<uint4*>(gdst) = a;
        *reinterpret_cast<uint4*>(gdst + 8) = b;
      }
    };
    // whole-head items: dQ of query block ib, complete in its TMEM accumulator (or zero: no key block touched it)
    auto store_dq = [&](const AbItem& t, int ib, bool touched, bool full) {
      uint32_t v[16];
      if (touched) {   // warp-uniform
        tmem_ld_32x32b_x16(tDQ + 64 * ib + lane_addr + 16 * cg, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0u;
      }
      const int qi = ib * 128 + r;
      put_rows(v, full, sPT + (uint32_t)ib * AB_TILE,
               p.dq + (t.dq_row0 + qi) * p.ld_dq + p.dq_col0 + t.h * 64 + 16 * cg, qi < t.Tq);
      if (p.dq_colsum && touched) attn_bwd_colsum16(v, qi < t.Tq, p.dq_colsum + t.h * 64 + 16 * cg, lane);
    };
    bool spt_busy = false;   // a TMA store may still be reading the dQ tiles staged in sPT
    // key-block items (Tq > 256): dQ of one query block -> fp32 tile in smem ([2 column halves][128 rows][32 fp32],
    // 128B-swizzled) that ONE TMA reduction adds into the fp32 workspace; the first version issued 4 scattered
    // red.global.add.v4 per thread (32 half-used sectors per warp instruction: 1.5 - 3 us per block in the LSU)
    auto stage_dq = [&](int buf, uint32_t tile) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(tDQ + 64 * buf + lane_addr + 16 * cg, v);
      tmem_ld_wait();
      const uint32_t row = tile + (uint32_t)(cg >> 1) * AB_TILE + (uint32_t)r * 128u;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t ch = (uint32_t)((cg & 1) * 4 + k) ^ (uint32_t)(r & 7);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (ch << 4)), "r"(v[4 * k]), "r"(v[4 * k + 1]),
                     "r"(v[4 * k + 2]), "r"(v[4 * k + 3]) : "memory");
      }
    };
    auto issue_dq_reduce = [&](const AbItem& t, int q0, uint32_t tile) {   // one thread
      tma_reduce_add_3d(&tm_ws, tile, t.h * 64, t.q_row0 + q0, t.q_bat);
      tma_reduce_add_3d(&tm_ws, tile + AB_TILE, t.h * 64 + 32, t.q_row0 + q0, t.q_bat);
    };
    int n_act = 0, g = 0;
    for (int rr = 0; rr < n_rounds; ++rr) {
      const int w = item_of(rr);
      if (w >= n_items) continue;
      const int jb_lo = whole ? 0 : w / n_bh, jb_hi = whole ? n_jb : jb_lo + 1, bh = whole ? w : w - jb_lo * n_bh;
      int last_act = -1;
      uint32_t dq_seen = 0;
      if (whole) {
        for (int jb = jb_lo; jb < jb_hi; ++jb) {
          const AbItem t = ab_item<CAUSAL>(p, bh, jb);
          if (t.active) {
            last_act = jb;
            for (int it = 0; it < t.n_iter; ++it) dq_seen |= 1u << (t.i_min + it);
          }
        }
        if (last_act < 0) {   // no key of this head is visible to any query: dQ = 0
          const AbItem t = ab_item<CAUSAL>(p, bh, 0);
          for (int ib = 0; ib * 128 < t.Tq; ++ib) store_dq(t, ib, false, false);
        }
      }
      for (int jb = jb_lo; jb < jb_hi; ++jb) {
      const AbItem t = ab_item<CAUSAL>(p, bh, jb);
      if (!t.exists) continue;
      const int kj = t.k0 + r;
      if (!t.active) {
        // keys that no query sees (or beyond kv_len): zero gradients
        if (kj < t.Tk) {
          __nv_bfloat16* dkp = p.dk + (t.dk_row0 + kj) * p.ld_dk + p.dk_col0 + t.h * 64 + 16 * cg;
          __nv_bfloat16* dvp = p.dv + (t.dk_row0 + kj) * p.ld_dv + p.dv_col0 + t.h * 64 + 16 * cg;
#pragma unroll
          for (int i = 0; i < 16; i += 8) {
            *reinterpret_cast<uint4*>(dkp + i) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(dvp + i) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        continue;
      }
      const bool key_ok = kj < t.kv_len;
      const uint32_t drop_shift = (kj & 1) ? 16u : 0u;  // which 16-bit lane of hash2(row, kj >> 1) is ours
      // hash2(row, col2) = avalanche((row * 2 ncol4 + col2) * W + key) = avalanche(row * drop_a + drop_b)
      const uint32_t drop_b = ((uint32_t)kj >> 1) * 0x9E3779B1u + p.drop.key;
      int pend_q0 = -1;
      for (int it = 0; it < t.n_iter; ++it, ++g) {
        const int st = g & 1;
        const int q0 = (t.i_min + it) * 128;
        // warp-uniform: does this (key rows, query block) tile need any masking?
        const bool need_mask = (t.k0 + qr * 32 + 31 > t.kv_len - 1) || (q0 + 127 > t.Tq - 1) ||
                               (CAUSAL && (t.k0 + qr * 32 + 31 > q0 + p.causal_off));
        if (spt_busy) {   // CTA-uniform: the previous item's dQ tiles must have left sPT before P^T is written there
          if (ct == 0) tma_store_wait_read();
          asm volatile("bar.sync 1, 512;" ::: "memory");
          spt_busy = false;
        }
        mbar_wait(qdo_full(st), (g >> 1) & 1);   // lse / delta stage written by the producer warp
        if (ct == 0) AB_LOG(8);
        mbar_wait(bar_sdp, g & 1);
        tc_fence_after();
        if (ct == 0) AB_LOG(1);
        // four straight-line variants of the element-wise stage, chosen once per (warp, query block)
        const AbElemCtx ec{tS + lane_addr, tDP + lane_addr, sStat + (uint32_t)st * 1024u, sPT, sDS, r, c,
                           p.scale, keep_scale, thr16 << 16, 16u - drop_shift,
                           (uint32_t)(t.stat_row0 + q0) * drop_a + drop_b, drop_a,
                           key_ok ? kj - p.causal_off : 0x7fffffff, key_ok, q0, t.Tq};
        if (p.do_drop) {
          if (need_mask) ab_elementwise<CAUSAL, true, true>(ec, 32 * cg);
          else ab_elementwise<CAUSAL, false, true>(ec, 32 * cg);
        } else {
          if (need_mask) ab_elementwise<CAUSAL, true, false>(ec, 32 * cg);
          else ab_elementwise<CAUSAL, false, false>(ec, 32 * cg);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pds);
        if (ct == 0) AB_LOG(2);
        // dQ of the PREVIOUS block (its MMAs completed before this block's S / dP did): off the critical path
        if (!whole && pend_q0 >= 0) {
          if (pend_q0 + 128 <= t.Tq) {   // CTA-uniform: whole 128-row box inside the sample -> staged TMA reduction
            if (ct == 0) tma_store_wait_read();          // sOut: its previous tile has been read by the TMA
            asm volatile("bar.sync 1, 512;" ::: "memory");
            stage_dq((g - 1) & 1, sOut);
            fence_proxy_async();
            asm volatile("bar.sync 1, 512;" ::: "memory");
            if (ct == 0) { issue_dq_reduce(t, pend_q0, sOut); tma_store_commit(); }
          } else {
            drain_dq(t, pend_q0, (g - 1) & 1);
          }
          if (ct == 0) AB_LOG(3);
        }
        pend_q0 = q0;
      }
      // item end: last dQ, then dK_j / dV_j (lanes are keys; this warp: rows qr*32.., columns 16*cg..+16)
      mbar_wait(bar_item, n_act & 1);
      tc_fence_after();
      if (ct == 0) AB_LOG(4);
      const bool kfull = t.k0 + 128 <= t.Tk;          // CTA-uniform: the whole 128-row box lies inside this sample
      const bool dq_now = whole && jb == last_act;
      const bool qfull0 = 128 <= t.Tq, qfull1 = 256 <= t.Tq;
      // staging tiles free again?  (dK / dV of the previous key block were handed to the TMA long ago)
      if (ct == 0) tma_store_wait_read();
      asm volatile("bar.sync 1, 512;" ::: "memory");
      const bool dq_staged = !whole && pend_q0 + 128 <= t.Tq;   // last dQ of a key-block item goes through sPT
      if (!whole) {
        if (dq_staged) stage_dq((g - 1) & 1, sPT);
        else drain_dq(t, pend_q0, (g - 1) & 1);
      } else if (dq_now) {
        store_dq(t, 0, dq_seen & 1u, qfull0);
        if (t.Tq > 128) store_dq(t, 1, (dq_seen >> 1) & 1u, qfull1);
      }
      if (ct == 0) AB_LOG(6);
      {
        // tcgen05.ld is warp-collective (.sync.aligned): issue it unconditionally, guard the stores
        uint32_t v[16], u[16];
        tmem_ld_32x32b_x16(tDK + lane_addr + 16 * cg, v);
        tmem_ld_32x32b_x16(tDV + lane_addr + 16 * cg, u);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dkv_free);   // accumulators are in registers: the next item may overwrite them
        if (ct == 0) AB_LOG(7);
        put_rows(v, kfull, sOut, p.dk + (t.dk_row0 + kj) * p.ld_dk + p.dk_col0 + t.h * 64 + 16 * cg, kj < t.Tk);
        put_rows(u, kfull, sOut + AB_TILE, p.dv + (t.dk_row0 + kj) * p.ld_dv + p.dv_col0 + t.h * 64 + 16 * cg, kj < t.Tk);
        if (p.dk_colsum) attn_bwd_colsum16(v, kj < t.Tk, p.dk_colsum + t.h * 64 + 16 * cg, lane);
        if (p.dv_colsum) attn_bwd_colsum16(u, kj < t.Tk, p.dv_colsum + t.h * 64 + 16 * cg, lane);
      }
      if (kfull || (dq_now && qfull0) || dq_staged) {   // CTA-uniform
        fence_proxy_async();
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (ct == 0) {
          if (dq_staged) issue_dq_reduce(t, pend_q0, sPT);
          if (kfull) {
            tma_store_3d(&tm_dk, sOut, p.dk_col0 + t.h * 64, t.k_row0 + t.k0, t.k_bat);
            tma_store_3d(&tm_dv, sOut + AB_TILE, p.dv_col0 + t.h * 64, t.k_row0 + t.k0, t.k_bat);
          }
          if (dq_now && qfull0) tma_store_3d(&tm_dq, sPT, p.dq_col0 + t.h * 64, t.q_row0, t.q_bat);
          if (dq_now && qfull1) tma_store_3d(&tm_dq, sPT + AB_TILE, p.dq_col0 + t.h * 64, t.q_row0 + 128, t.q_bat);
          tma_store_commit();
        }
        spt_busy = (dq_now && qfull0) || dq_staged;
      }
      if (ct == 0) AB_LOG(5);
      ++n_act;
      }
    }
    if (ct == 0) tma_store_wait_all();   // smem must outlive the bulk stores
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) AB_MARK(62);
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
  pdl_wait();
  pdl_launch_dependents();
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

extern "C" int ergm_attn_bwd_workspace_bytes(int B, int nh, int Tq, int64_t* bytes) {
  if (B <= 0 || nh <= 0 || Tq <= 0 || !bytes) return ERGM_ERR_ARG;
  // Tq <= 256: a CTA owns a whole (batch, head) and keeps dQ in TMEM; longer sequences: key-block items add their
  // dQ contributions into an fp32 scratch that is cast (and column-summed) afterwards
  *bytes = Tq <= 256 ? 0 : (int64_t)B * Tq * nh * 64 * (int64_t)sizeof(float);
  return ERGM_OK;
}

extern "C" int ergm_attn_bwd(const void* q, int64_t ld_q, int q_col0, const void* k, int64_t ld_k,
                             int k_col0, const void* v, int64_t ld_v, int v_col0, const void* out,
                             int64_t ld_out, const float* out_f32, const void* dout, int64_t ld_do, const float* lse,
                             float* delta, void* dq, int64_t ld_dq, int dq_col0, void* dk, int64_t ld_dk,
                             int dk_col0, void* dv, int64_t ld_dv, int dv_col0, float* dq_colsum, float* dk_colsum,
                             float* dv_colsum, const int* kv_lens, int B, int nh, int Tq, int Tk, int head_dim,
                             int causal, int causal_off, float dropout_p, uint64_t seed, uint64_t offset,
                             const ergm_pack* pack, int pack_kv, void* workspace, int64_t workspace_bytes,
                             void* stream) {
  if (!q || !k || !v || !out || !dout || !lse || !delta || !dq || !dk || !dv) return ERGM_ERR_ARG;
  if (pack && (!pack->cu_rows || !pack->row_b || !pack->n_rows || (pack_kv && !pack->kv_lens))) return ERGM_ERR_ARG;
  const uint64_t q_rows = pack ? (uint64_t)B * Tq : (uint64_t)Tq, q_bat = pack ? 1 : (uint64_t)B;
  const uint64_t k_rows = (pack && pack_kv) ? (uint64_t)B * Tk : (uint64_t)Tk, k_bat = (pack && pack_kv) ? 1 : (uint64_t)B;
  if (B <= 0 || nh <= 0 || Tq <= 0 || Tk <= 0) return ERGM_ERR_ARG;
  if (head_dim != 64) return ERGM_ERR_UNSUPPORTED;
  if (ld_q % 8 || ld_k % 8 || ld_v % 8 || ld_out % 8 || ld_do % 8 || ld_dq % 8 || ld_dk % 8 || ld_dv % 8 ||
      q_col0 % 8 || k_col0 % 8 || v_col0 % 8 || dq_col0 % 8 || dk_col0 % 8 || dv_col0 % 8)
    return ERGM_ERR_ARG;
  int64_t need = 0;
  ergm_attn_bwd_workspace_bytes(B, nh, Tq, &need);
  if (need > 0 && (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 15))) return ERGM_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  ERGM_CUDA_TRY(launch_pdl(attn_delta_kernel, dim3((unsigned)((B * Tq + 7) / 8)), dim3(256), 0, s, 1,
                           reinterpret_cast<const __nv_bfloat16*>(dout), ld_do, reinterpret_cast<const __nv_bfloat16*>(out),
                           ld_out, (const float*)out_f32, delta, B * Tq, Tq, nh, pack ? pack->cu_rows : (const int*)nullptr,
                           pack ? pack->row_b : (const int*)nullptr, pack ? pack->n_rows : (const int*)nullptr));
  CUtensorMap tq, tk, tv, tdo, tdq, tdk, tdv, tws;
  int rc;
  if ((rc = encode_tmap_3d(&tq, q, 2, (uint64_t)(q_col0 + nh * 64), q_rows, q_bat,
                           (uint64_t)ld_q * 2, q_rows * ld_q * 2, 64, 128, 1))) return rc;
  if ((rc = encode_tmap_3d(&tk, k, 2, (uint64_t)(k_col0 + nh * 64), k_rows, k_bat,
                           (uint64_t)ld_k * 2, k_rows * ld_k * 2, 64, 128, 1))) return rc;
  if ((rc = encode_tmap_3d(&tv, v, 2, (uint64_t)(v_col0 + nh * 64), k_rows, k_bat,
                           (uint64_t)ld_v * 2, k_rows * ld_v * 2, 64, 128, 1))) return rc;
  if ((rc = encode_tmap_3d(&tdo, dout, 2, (uint64_t)(nh * 64), q_rows, q_bat,
                           (uint64_t)ld_do * 2, q_rows * ld_do * 2, 64, 128, 1))) return rc;
  // outputs (TMA stores of whole 128-row blocks; blocks that end inside a sample are stored by the threads)
  if ((rc = encode_tmap_3d(&tdq, dq, 2, (uint64_t)(dq_col0 + nh * 64), q_rows, q_bat,
                           (uint64_t)ld_dq * 2, q_rows * ld_dq * 2, 64, 128, 1))) return rc;
  if ((rc = encode_tmap_3d(&tdk, dk, 2, (uint64_t)(dk_col0 + nh * 64), k_rows, k_bat,
                           (uint64_t)ld_dk * 2, k_rows * ld_dk * 2, 64, 128, 1))) return rc;
  if ((rc = encode_tmap_3d(&tdv, dv, 2, (uint64_t)(dv_col0 + nh * 64), k_rows, k_bat,
                           (uint64_t)ld_dv * 2, k_rows * ld_dv * 2, 64, 128, 1))) return rc;
  tws = tdq;   // (unused by whole-head items)
  if (need > 0 && (rc = encode_tmap_3d(&tws, workspace, 4, (uint64_t)(nh * 64), q_rows, q_bat, (uint64_t)nh * 64 * 4,
                                       q_rows * (uint64_t)nh * 64 * 4, 32, 128, 1))) return rc;
  AttnBwdParams p;
  p.lse = lse; p.delta = delta; p.dq_accum = reinterpret_cast<float*>(workspace);
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq);
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.dq_colsum = dq_colsum; p.dk_colsum = dk_colsum; p.dv_colsum = dv_colsum;
  p.kv_lens = kv_lens;
  p.cu_q = pack ? pack->cu_rows : nullptr;
  p.cu_k = (pack && pack_kv) ? pack->cu_rows : nullptr;
  if (pack && pack_kv) p.kv_lens = pack->kv_lens;
  p.ld_dq = ld_dq; p.ld_dk = ld_dk; p.ld_dv = ld_dv; p.dq_col0 = dq_col0; p.dk_col0 = dk_col0; p.dv_col0 = dv_col0;
  p.B = B; p.Tq = Tq; p.Tk = Tk; p.nh = nh;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.causal_off = causal_off;
  p.scale = 1.0f / sqrtf((float)head_dim);
  p.drop = make_site(seed, offset, dropout_p, (uint32_t)Tk);
  p.do_drop = dropout_p > 0.f;
  p.whole_head = need == 0;
#ifdef ERGM_ATTN_TRACE
  p.trace = g_attn_bwd_trace;
#endif
  if (!p.whole_head) {
    cudaError_t ce = cudaMemsetAsync(workspace, 0, (size_t)need, s);
    if (ce != cudaSuccess) return (int)ce;
  }
  ERGM_SET_SMEM_ATTR(attn_bwd_kernel<true>, AB_SMEM);
  ERGM_SET_SMEM_ATTR(attn_bwd_kernel<false>, AB_SMEM);
  const int n_items = p.whole_head ? B * nh : B * nh * ((Tk + 127) / 128);
  dim3 grid(n_items < num_sms() ? n_items : num_sms());
  if (causal)
    ERGM_CUDA_TRY(launch_pdl(attn_bwd_kernel<true>, grid, dim3(AB_THREADS), (size_t)AB_SMEM, s, 1, tq, tk, tv, tdo, tdq, tdk, tdv, tws, p));
  else
    ERGM_CUDA_TRY(launch_pdl(attn_bwd_kernel<false>, grid, dim3(AB_THREADS), (size_t)AB_SMEM, s, 1, tq, tk, tv, tdo, tdq, tdk, tdv, tws, p));
  if (!p.whole_head) {
    // dQ contributions of the key blocks -> bf16 dQ (+ its column sums: the bias gradient of the Q projection)
    return ergm_cast_f32_bf16_2d(reinterpret_cast<const float*>(workspace), (int64_t)nh * 64,
                                 reinterpret_cast<__nv_bfloat16*>(dq) + dq_col0, ld_dq, B * Tq, nh * 64, dq_colsum,
                                 pack ? pack->n_rows : nullptr, stream);
  }
  return ERGM_OK;
}

#ifdef ERGM_ATTN_TRACE
extern "C" int ergm_attn_bwd_set_trace(long long* dev_ptr) { ergm::g_attn_bwd_trace = dev_ptr; return ERGM_OK; }
#endif

// ergm_decode_layers - the whole transformer stack of one decode step as ONE persistent kernel.
//
// profiles/r1_decode.md: as a chain of ~60 dependent launches the decode step is bound by the
// dependent-launch floor (~4 us per kernel), not by HBM: every kernel moves only 1-19 MB.  Here the
// L x {LN1+QKV, paged attention, out-proj +=, [LN+q, cross attention, out-proj +=], LN2+FC+gelu,
// MLP-proj +=} phases (model.py:286-341 per block) run inside one cooperative kernel, 148 CTAs x 512
// threads, separated by grid barriers (one atomic + acquire spin, ~1 us) instead of kernel boundaries:
//   * the GEMM phases are the slab kernels of decode_gemm.cu (same packed weights, same fragment
//     layout, A fragments register-resident, cross-warp K reduction), with a virtual (slab, k-split)
//     grid per phase;
//   * the weight slabs of the NEXT GEMM phase are bulk-copied into the shared-memory ring before the
//     grid barrier (weights never depend on the step's data), so their DRAM latency hides behind the
//     barrier and the attention phase;
//   * the attention phase runs four independent 128-thread groups per CTA (named barriers), each one
//     (head, sequence) item at a time - the flash-decoding body of decode.cu.
// Data written by other CTAs in an earlier phase is always read with ordinary loads (generic proxy)
// after the barrier's fence; only the immutable weights use the async (bulk-copy) proxy.
#include "../../include/ergm_b200.h"
#include "common.cuh"
#include <cstdlib>

namespace ergm {

constexpr int MG_THREADS = 512;
constexpr int MG_M = 64;
constexpr int MG_MAX_STAGES = 4;
constexpr int MG_PAGE = 16;
constexpr int MG_UNR = 4;   // tokens per 8-lane group per pass (8 spilled at the 128-register cap of 512 threads)

struct MegaLayer {  // one per transformer block, device resident (built once per generation batch)
  const __nv_bfloat16* w_qkv; const float* b_qkv;  // ln_1 folded in
  const __nv_bfloat16* w_o;   const float* b_o;
  const __nv_bfloat16* w_q2;  const float* b_q2;   // ln_cross_attn folded in (cross attention, nullable)
  const __nv_bfloat16* w_o2;  const float* b_o2;
  const __nv_bfloat16* w_fc;  const float* b_fc;   // ln_2 folded in
  const __nv_bfloat16* w_p2;  const float* b_p2;
  __nv_bfloat16* pool;                              // paged self-attention K/V of this layer
  const __nv_bfloat16* kv2;                         // cached cross-attention K/V [B*Tc, 2H] (nullable)
};

struct MegaParams {
  const MegaLayer* layers;
  int L, H, I, nh, B;
  float* x;              // [B, H] fp32 residual stream (in: embeddings, out: last block's output)
  __nv_bfloat16* qkv;    // [B, 3H]
  __nv_bfloat16* ctx;    // [B, H]
  __nv_bfloat16* q2;     // [B, H]
  __nv_bfloat16* g;      // [B, I]
  const int* block_table;
  const int* seq_lens;
  int max_pages, Tc;
  float eps, scale;
  unsigned int* sync_ctr;  // zeroed by the host before every launch
  int a_bytes;             // shared-memory layout: [128 B barriers][a_bytes A / reduction][stages x slab_stride]
  int slab_stride;
  int stages;
  int dbg;                 // timing experiments (ERGM_MEGA_DBG): 1 skip attention, 2 skip GEMM phases, 4 skip grid barriers
};

ERGM_DEVINL void mg_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
ERGM_DEVINL void mg_ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
ERGM_DEVINL void mg_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
ERGM_DEVINL unsigned int mg_ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// grid-wide barrier: every CTA arrives once per call; `epoch` counts the calls (identical in all threads)
ERGM_DEVINL void mg_grid_sync(unsigned int* ctr, unsigned int& epoch, int dbg = 0) {
  if (dbg & 4) { __syncthreads(); return; }
  __syncthreads();
  ++epoch;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");  // no round trip before the spin
    const unsigned int target = epoch * gridDim.x;
    while (mg_ld_acquire(ctr) < target) {}
    __threadfence();
  }
  __syncthreads();
}

// Geometry of one GEMM phase on the persistent grid.
struct MgGemm {
  const __nv_bfloat16* w;   // packed slabs
  const float* bias;
  int K, N, NB;             // NB = slabs
  int KBc;                  // k16 blocks per CTA k-range
  int gx, gy;               // virtual grid: gx slab lanes x gy k-splits (gx * gy <= gridDim.x)
};

ERGM_DEVINL MgGemm mg_make_gemm(const __nv_bfloat16* w, const float* bias, int K, int N, int Kc, int G) {
  MgGemm g;
  g.w = w; g.bias = bias; g.K = K; g.N = N; g.NB = (N + 15) / 16;
  g.KBc = Kc / 16;
  g.gy = K / Kc;
  g.gx = min(g.NB, G / g.gy);
  return g;
}

// producer side of a GEMM phase: thread 0 arms the ring with this CTA's first slabs.  May be called
// before the grid barrier that precedes the phase (weights are immutable).
ERGM_DEVINL void mg_gemm_prefetch(const MgGemm& g, const MegaParams& p, unsigned char* smem) {
  const int c = blockIdx.x;
  if (threadIdx.x != 0 || c >= g.gx * g.gy) return;
  const uint32_t bars = smem_u32(smem);
  unsigned char* ring = smem + 128 + p.a_bytes;
  const int vx = c % g.gx, ks = c / g.gx;
  const uint32_t slab_bytes = (uint32_t)g.KBc * 512u;
  const int KB = g.K / 16;
  int slab = vx;
  for (int s = 0; s < p.stages && slab < g.NB; ++s, slab += g.gx) {
    mbar_expect_tx(bars + 8 * s, slab_bytes);
    mg_bulk_g2s(smem_u32(ring + (size_t)s * p.slab_stride), g.w + ((int64_t)slab * KB + (int64_t)ks * g.KBc) * 256,
                slab_bytes, bars + 8 * s);
  }
}

// consumer side.  LN: A = (x - mean) * rstd of the fp32 residual stream (gamma / beta live in the packed
// weight / bias); otherwise A = a[:, ks*Kc : (ks+1)*Kc] (bf16).  KS = k16 steps per warp (KBc / 8).
// out_mode 0: bf16 store (+gelu), 2: fp32 += .  `par`: per-stage mbarrier parity bits (all threads).
template <bool LN, int KS>
ERGM_DEVINL void mg_gemm_compute(const MgGemm& g, const MegaParams& p, unsigned char* smem, const float* x,
                                 const __nv_bfloat16* a, int64_t lda, void* out, int64_t ldo, int out_mode, int gelu,
                                 uint32_t& par) {
  const int c = blockIdx.x;
  if (c >= g.gx * g.gy) return;
  if (p.dbg & 2) {  // timing experiment: consume the armed stages, no math
    int it = 0;
    for (int slab = c % g.gx; slab < g.NB; slab += g.gx, ++it) {
      const int stage = it % p.stages;
      mbar_wait(smem_u32(smem) + 8 * stage, (par >> stage) & 1u);
      par ^= 1u << stage;
      __syncthreads();
      const int next = slab + p.stages * g.gx;
      if (threadIdx.x == 0 && next < g.NB) {
        mbar_expect_tx(smem_u32(smem) + 8 * stage, (uint32_t)g.KBc * 512u);
        mg_bulk_g2s(smem_u32(smem + 128 + p.a_bytes + (size_t)stage * p.slab_stride),
                    g.w + ((int64_t)next * (g.K / 16) + (int64_t)(c / g.gx) * g.KBc) * 256, (uint32_t)g.KBc * 512u,
                    smem_u32(smem) + 8 * stage);
      }
    }
    return;
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int vx = c % g.gx, ks = c / g.gx;
  const int Kc = g.KBc * 16;
  const int a_stride = Kc * 2 + 16;
  const uint32_t slab_bytes = (uint32_t)g.KBc * 512u;
  const uint32_t bars = smem_u32(smem);
  unsigned char* a_sm = smem + 128;
  float* red = reinterpret_cast<float*>(a_sm);
  unsigned char* ring = smem + 128 + p.a_bytes;
  const int KB = g.K / 16;
  const int M = p.B;
  // ---- A operand -> smem ----
  if constexpr (LN) {
    constexpr int NV = KS;  // K / 128
    const float invH = 1.f / (float)(NV * 128);
    for (int rr = 0; rr < 4; rr += 2) {  // 16 warps x 4 rows, two rows at a time
      const int row0 = warp * 4 + rr;
      float4 v[2][NV];
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int i = 0; i < NV; ++i)
          v[q][i] = (row0 + q) < M ? __ldcg(reinterpret_cast<const float4*>(x + (int64_t)(row0 + q) * lda) + lane + 32 * i)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int row = row0 + q;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) s += (v[q][i].x + v[q][i].y) + (v[q][i].z + v[q][i].w);
        const float mean = warp_sum(s) * invH;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float dx = v[q][i].x - mean, dy = v[q][i].y - mean, dz = v[q][i].z - mean, dw = v[q][i].w - mean;
          ss += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
        const float rstd = rsqrtf(warp_sum(ss) * invH + p.eps);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          uint2 o = make_uint2(0u, 0u);
          if (row < M)
            o = make_uint2(pack_bf16x2((v[q][i].x - mean) * rstd, (v[q][i].y - mean) * rstd),
                           pack_bf16x2((v[q][i].z - mean) * rstd, (v[q][i].w - mean) * rstd));
          *reinterpret_cast<uint2*>(a_sm + (size_t)row * a_stride + (lane + 32 * i) * 8) = o;
        }
      }
    }
  } else {
    const int chunks = Kc / 8;  // 16-byte chunks per row
    const __nv_bfloat16* src = a + (int64_t)ks * Kc;
    const int total = MG_M * chunks;
    for (int base = 0; base < total; base += 4 * MG_THREADS) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = base + u * MG_THREADS + tid;
        const int r = idx / chunks, cc = idx - r * chunks;
        v[u] = (idx < total && r < M) ? __ldcg(reinterpret_cast<const uint4*>(src + (int64_t)r * lda + cc * 8))
                                      : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = base + u * MG_THREADS + tid;
        if (idx < total) {
          const int r = idx / chunks, cc = idx - r * chunks;
          *reinterpret_cast<uint4*>(a_sm + (size_t)r * a_stride + cc * 16) = v[u];
        }
      }
    }
  }
  __syncthreads();
  // ---- A fragments -> registers ----
  const int mh = warp & 1, kw = warp >> 1;
  uint32_t afr[2][KS][4];
  {
    const uint32_t a_base = smem_u32(a_sm) + (uint32_t)(mh * 32 + (lane & 7) + ((lane >> 3) & 1) * 8) * a_stride +
                            (uint32_t)(lane >> 4) * 16 + (uint32_t)(kw * KS) * 32;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int k = 0; k < KS; ++k)
        mg_ldmatrix_x4(a_base + (uint32_t)(mi * 16) * a_stride + k * 32, afr[mi][k][0], afr[mi][k][1], afr[mi][k][2],
                       afr[mi][k][3]);
  }
  __syncthreads();  // the A region now serves as the reduction buffer
  const int rmh = tid >= 256 ? 1 : 0;
  int it = 0;
  for (int slab = vx; slab < g.NB; slab += g.gx, ++it) {
    const int stage = it % p.stages;
    const int e2 = tid - rmh * 256;
    const int col = slab * 16 + ((e2 >> 6) & 1) * 8 + ((e2 >> 1) & 3) * 2;
    float2 bias_v = make_float2(0.f, 0.f);
    if (g.bias && (out_mode != 2 || ks == 0)) {
      if (col < g.N) bias_v.x = __ldg(g.bias + col);
      if (col + 1 < g.N) bias_v.y = __ldg(g.bias + col + 1);
    }
    mbar_wait(bars + 8 * stage, (par >> stage) & 1u);
    par ^= 1u << stage;
    const unsigned char* ws = ring + (size_t)stage * p.slab_stride + (size_t)(kw * KS) * 512 + lane * 16;
    float acc[2][2][4];
#pragma unroll
    for (int i0 = 0; i0 < 2; ++i0)
#pragma unroll
      for (int i1 = 0; i1 < 2; ++i1)
#pragma unroll
        for (int i2 = 0; i2 < 4; ++i2) acc[i0][i1][i2] = 0.f;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const uint4 b = *reinterpret_cast<const uint4*>(ws + (size_t)k * 512);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        mg_mma(acc[mi][0], afr[mi][k][0], afr[mi][k][1], afr[mi][k][2], afr[mi][k][3], b.x, b.y);
        mg_mma(acc[mi][1], afr[mi][k][0], afr[mi][k][1], afr[mi][k][2], afr[mi][k][3], b.z, b.w);
      }
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int n8 = 0; n8 < 2; ++n8)
        *reinterpret_cast<float4*>(red + ((warp * 4 + mi * 2 + n8) * 32 + lane) * 4) =
            make_float4(acc[mi][n8][0], acc[mi][n8][1], acc[mi][n8][2], acc[mi][n8][3]);
    __syncthreads();
    {
      const int q = e2 >> 6, l2 = (e2 >> 1) & 31, sub = (e2 & 1) * 2;
      float v0 = 0.f, v1 = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const float2 t2 = *reinterpret_cast<const float2*>(red + (((w * 2 + rmh) * 4 + q) * 32 + l2) * 4 + sub);
        v0 += t2.x; v1 += t2.y;
      }
      const int row = rmh * 32 + (q >> 1) * 16 + (l2 >> 2) + (sub ? 8 : 0);
      const bool c0 = col < g.N, c1 = col + 1 < g.N;
      if (row < M && c0) {
        v0 += bias_v.x; v1 += bias_v.y;
        if (gelu) { v0 = gelu_new<false>(v0); v1 = gelu_new<false>(v1); }
        if (out_mode == 0) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + (int64_t)row * ldo + col;
          if (c1) *reinterpret_cast<uint32_t*>(o) = pack_bf16x2(v0, v1);
          else *o = __float2bfloat16_rn(v0);
        } else {
          float* o = reinterpret_cast<float*>(out) + (int64_t)row * ldo + col;
          atomicAdd(o, v0);
          if (c1) atomicAdd(o + 1, v1);
        }
      }
    }
    __syncthreads();  // `red` and this ring stage are free again
    const int next = slab + p.stages * g.gx;
    if (tid == 0 && next < g.NB) {
      mbar_expect_tx(bars + 8 * stage, slab_bytes);
      mg_bulk_g2s(smem_u32(ring + (size_t)stage * p.slab_stride), g.w + ((int64_t)next * KB + (int64_t)ks * g.KBc) * 256,
                  slab_bytes, bars + 8 * stage);
    }
  }
}

ERGM_DEVINL float mg_dot8(const uint4 a, const uint4 b) {
  const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
  const float2 b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y), b2 = unpack_bf16x2(b.z), b3 = unpack_bf16x2(b.w);
  return a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
}

ERGM_DEVINL void mg_group_sync(int grp4) {
  asm volatile("bar.sync %0, 128;" ::"r"(grp4 + 1) : "memory");
}

// One-query attention over the cached context, four independent 128-thread groups per CTA, one
// (head, sequence) item per group at a time (decode.cu's flash-decoding body).  PAGED: self attention
// over the page pool, appends the new token's K / V; otherwise cross attention over kv2.
template <bool PAGED>
ERGM_DEVINL void mg_attention(const MegaParams& p, unsigned char* smem, const __nv_bfloat16* q, int64_t ld_q, int q_col0,
                              int k_col0, int v_col0, __nv_bfloat16* pool, const __nv_bfloat16* kc, int64_t ld_k) {
  const int grp4 = threadIdx.x >> 7, t = threadIdx.x & 127;
  const int grp = t >> 3, gl = t & 7, lane = t & 31;
  // per-group scratch inside the (currently unused) A region
  unsigned char* base = smem + 128 + (size_t)grp4 * (16 * 64 * 4 + 2 * 16 * 4);
  float* s_acc = reinterpret_cast<float*>(base);             // [16][64]
  float* s_m = s_acc + 16 * 64;                                // [16]
  float* s_l = s_m + 16;                                       // [16]
  // block table + sequence lengths of the whole batch: staged once per kernel behind the slab ring
  const int* s_tab = reinterpret_cast<const int*>(smem + 128 + p.a_bytes + (size_t)p.stages * p.slab_stride);
  const int* s_len = s_tab + p.B * p.max_pages;
  const int n_items = p.nh * p.B;
  for (int item = blockIdx.x * 4 + grp4; item < n_items; item += gridDim.x * 4) {
    const int h = item % p.nh, b = item / p.nh;
    int n_old, pos = 0;
    if (PAGED) {
      pos = s_len[b];
      n_old = pos;
    } else {
      n_old = p.Tc;
    }
    const int* s_bt = s_tab + b * p.max_pages;
    mg_group_sync(grp4);  // previous item's scratch reads are done
    auto kv_row = [&](int tok, int which) -> const uint4* {
      if (PAGED) {
        const int page = s_bt[tok / MG_PAGE];
        return reinterpret_cast<const uint4*>(pool + ((((int64_t)page * 2 + which) * p.nh + h) * MG_PAGE + tok % MG_PAGE) * 64) + gl;
      }
      return reinterpret_cast<const uint4*>(kc + ((int64_t)b * p.Tc + tok) * ld_k + (which ? v_col0 : k_col0) + h * 64) + gl;
    };
    const uint4 qv = __ldcg(reinterpret_cast<const uint4*>(q + (int64_t)b * ld_q + q_col0 + h * 64 + gl * 8));
    uint4 kn = make_uint4(0u, 0u, 0u, 0u), vn = kn;
    if (PAGED && t < 32) {
      kn = __ldcg(reinterpret_cast<const uint4*>(q + (int64_t)b * ld_q + k_col0 + h * 64 + gl * 8));
      vn = __ldcg(reinterpret_cast<const uint4*>(q + (int64_t)b * ld_q + v_col0 + h * 64 + gl * 8));
    }
    float m_run = -INFINITY, l_run = 0.f;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int tb = 0; tb < n_old; tb += 16 * MG_UNR) {
      uint4 kk[MG_UNR], vv[MG_UNR];
#pragma unroll
      for (int u = 0; u < MG_UNR; ++u) {
        const int tok = tb + u * 16 + grp;
        const bool ok = tok < n_old;
        kk[u] = ok ? *kv_row(tok, 0) : make_uint4(0u, 0u, 0u, 0u);
        vv[u] = ok ? *kv_row(tok, 1) : make_uint4(0u, 0u, 0u, 0u);
      }
      float sc[MG_UNR];
      float m_new = m_run;
#pragma unroll
      for (int u = 0; u < MG_UNR; ++u) {
        float s = mg_dot8(qv, kk[u]);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        sc[u] = (tb + u * 16 + grp < n_old) ? s * p.scale : -INFINITY;
        m_new = fmaxf(m_new, sc[u]);
      }
      if (m_new > -INFINITY) {
        const float corr = __expf(m_run - m_new);
        l_run *= corr;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] *= corr;
#pragma unroll
        for (int u = 0; u < MG_UNR; ++u) {
          const float w = __expf(sc[u] - m_new);
          l_run += w;
          const float2 v0 = unpack_bf16x2(vv[u].x), v1 = unpack_bf16x2(vv[u].y), v2 = unpack_bf16x2(vv[u].z),
                       v3 = unpack_bf16x2(vv[u].w);
          acc[0] += w * v0.x; acc[1] += w * v0.y; acc[2] += w * v1.x; acc[3] += w * v1.y;
          acc[4] += w * v2.x; acc[5] += w * v2.y; acc[6] += w * v3.x; acc[7] += w * v3.y;
        }
        m_run = m_new;
      }
    }
    if (PAGED && t < 32) {
      // the new token: from the projection output; the group's warp 0 scores it, 8-lane group 0 folds it into
      // its state, groups 0 / 1 append K / V to the page pool
      float s = mg_dot8(qv, kn);
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
        const int page = s_bt[pos / MG_PAGE];
        __nv_bfloat16* dst = pool + ((((int64_t)page * 2 + grp) * p.nh + h) * MG_PAGE + pos % MG_PAGE) * 64 + gl * 8;
        *reinterpret_cast<uint4*>(dst) = grp == 0 ? kn : vn;
      }
    }
    if (gl == 0) { s_m[grp] = m_run; s_l[grp] = l_run; }
#pragma unroll
    for (int i = 0; i < 8; ++i) s_acc[grp * 64 + gl * 8 + i] = acc[i];
    mg_group_sync(grp4);
    if (t < 64) {
      float Mx = -INFINITY;
#pragma unroll
      for (int g2 = 0; g2 < 16; ++g2) Mx = fmaxf(Mx, s_m[g2]);
      float Lsum = 0.f, o = 0.f;
#pragma unroll
      for (int g2 = 0; g2 < 16; ++g2) {
        const float w = s_m[g2] > -INFINITY ? __expf(s_m[g2] - Mx) : 0.f;
        Lsum += s_l[g2] * w;
        o += s_acc[g2 * 64 + t] * w;
      }
      p.ctx[(int64_t)b * p.H + h * 64 + t] = __float2bfloat16_rn(Lsum > 0.f ? o / Lsum : 0.f);
    }
    (void)lane;
  }
}

// KSL = H / 128 (LayerNorm-fused phases), direct phases use K ranges of 256 (out-proj) / 512 (MLP proj).
template <int KSL>
__global__ void __launch_bounds__(MG_THREADS, 1) decode_layers_kernel(const MegaParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int G = gridDim.x;
  unsigned int epoch = 0;
  uint32_t par = 0;  // mbarrier parities of the ring stages
  if (threadIdx.x == 0) {
    for (int s = 0; s < MG_MAX_STAGES; ++s) mbar_init(smem_u32(smem) + 8 * s, 1);
    fence_mbar_init();
  }
  {
    int* s_tab = reinterpret_cast<int*>(smem + 128 + p.a_bytes + (size_t)p.stages * p.slab_stride);
    const int nt = p.B * p.max_pages;
    for (int i = threadIdx.x; i < nt; i += MG_THREADS) s_tab[i] = p.block_table[i];
    for (int i = threadIdx.x; i < p.B; i += MG_THREADS) s_tab[nt + i] = p.seq_lens[i];
  }
  __syncthreads();
  const int H = p.H, I = p.I;
  const bool cross = p.Tc > 0;
  {
    const MegaLayer& l0 = p.layers[0];
    mg_gemm_prefetch(mg_make_gemm(l0.w_qkv, l0.b_qkv, H, 3 * H, H, G), p, smem);
  }
  for (int l = 0; l < p.L; ++l) {
    const MegaLayer ly = p.layers[l];
    const MgGemm g_qkv = mg_make_gemm(ly.w_qkv, ly.b_qkv, H, 3 * H, H, G);
    const int kc_o = H >= 256 ? 256 : 128;
    const MgGemm g_o = mg_make_gemm(ly.w_o, ly.b_o, H, H, kc_o, G);
    const MgGemm g_fc = mg_make_gemm(ly.w_fc, ly.b_fc, H, I, H, G);
    const MgGemm g_p2 = mg_make_gemm(ly.w_p2, ly.b_p2, I, H, 512, G);
    // ---- LN1 + QKV (weights prefetched before the previous barrier) ----
    mg_gemm_compute<true, KSL>(g_qkv, p, smem, p.x, nullptr, H, p.qkv, 3 * H, 0, 0, par);
    mg_gemm_prefetch(g_o, p, smem);
    mg_grid_sync(p.sync_ctr, epoch, p.dbg);
    // ---- paged self attention (+ append) ----
    if (!(p.dbg & 1)) mg_attention<true>(p, smem, p.qkv, 3 * H, 0, H, 2 * H, ly.pool, nullptr, 0);
    mg_grid_sync(p.sync_ctr, epoch, p.dbg);
    // ---- out-proj += residual ----
    if (H >= 256) mg_gemm_compute<false, 2>(g_o, p, smem, nullptr, p.ctx, H, p.x, H, 2, 0, par);
    else mg_gemm_compute<false, 1>(g_o, p, smem, nullptr, p.ctx, H, p.x, H, 2, 0, par);
    if (cross) {
      const MgGemm g_q2 = mg_make_gemm(ly.w_q2, ly.b_q2, H, H, H, G);
      const MgGemm g_o2 = mg_make_gemm(ly.w_o2, ly.b_o2, H, H, kc_o, G);
      mg_gemm_prefetch(g_q2, p, smem);
      mg_grid_sync(p.sync_ctr, epoch, p.dbg);
      mg_gemm_compute<true, KSL>(g_q2, p, smem, p.x, nullptr, H, p.q2, H, 0, 0, par);
      mg_gemm_prefetch(g_o2, p, smem);
      mg_grid_sync(p.sync_ctr, epoch, p.dbg);
      if (!(p.dbg & 1)) mg_attention<false>(p, smem, p.q2, H, 0, 0, H, nullptr, ly.kv2, 2 * H);
      mg_grid_sync(p.sync_ctr, epoch, p.dbg);
      if (H >= 256) mg_gemm_compute<false, 2>(g_o2, p, smem, nullptr, p.ctx, H, p.x, H, 2, 0, par);
      else mg_gemm_compute<false, 1>(g_o2, p, smem, nullptr, p.ctx, H, p.x, H, 2, 0, par);
    }
    mg_gemm_prefetch(g_fc, p, smem);
    mg_grid_sync(p.sync_ctr, epoch, p.dbg);
    // ---- LN2 + FC + gelu ----
    mg_gemm_compute<true, KSL>(g_fc, p, smem, p.x, nullptr, H, p.g, I, 0, 1, par);
    mg_gemm_prefetch(g_p2, p, smem);
    mg_grid_sync(p.sync_ctr, epoch, p.dbg);
    // ---- MLP proj += residual ----
    mg_gemm_compute<false, 4>(g_p2, p, smem, nullptr, p.g, I, p.x, H, 2, 0, par);
    if (l + 1 < p.L) {
      const MegaLayer& nx = p.layers[l + 1];
      mg_gemm_prefetch(mg_make_gemm(nx.w_qkv, nx.b_qkv, H, 3 * H, H, G), p, smem);
      mg_grid_sync(p.sync_ctr, epoch, p.dbg);
    }
  }
}

}  // namespace ergm

using namespace ergm;

extern "C" int ergm_decode_layers(const void* layer_table, int L, int H, int I, int nh, int B, float* x, void* qkv,
                                  void* ctx, void* q2, void* g, const int* block_table, const int* seq_lens,
                                  int max_pages, int Tc, float eps, unsigned int* sync_ctr, void* stream) {
  if (!layer_table || !x || !qkv || !ctx || !g || !block_table || !seq_lens || !sync_ctr) return ERGM_ERR_ARG;
  if (L <= 0 || B <= 0 || B > MG_M || nh <= 0 || H != nh * 64) return ERGM_ERR_ARG;
  if ((H != 768 && H != 1024 && H != 128 && H != 256 && H != 512) || I % 512) return ERGM_ERR_UNSUPPORTED;
  if (Tc > 0 && !q2) return ERGM_ERR_ARG;
  MegaParams p{};
  p.layers = reinterpret_cast<const MegaLayer*>(layer_table);
  p.L = L; p.H = H; p.I = I; p.nh = nh; p.B = B;
  p.x = x;
  p.qkv = reinterpret_cast<__nv_bfloat16*>(qkv); p.ctx = reinterpret_cast<__nv_bfloat16*>(ctx);
  p.q2 = reinterpret_cast<__nv_bfloat16*>(q2); p.g = reinterpret_cast<__nv_bfloat16*>(g);
  p.block_table = block_table; p.seq_lens = seq_lens; p.max_pages = max_pages; p.Tc = Tc;
  p.eps = eps; p.scale = 0.125f;
  p.sync_ctr = sync_ctr;
  // smem: A region sized for the largest phase (LN: 64 x (2H + 16); MLP proj: 64 x (1024 + 16); reduction 32 KB;
  // attention scratch 4 x 5 KB), ring of up to 4 slabs of the largest slab (H/16 x 512 B)
  int a_bytes = MG_M * (2 * H + 16);
  if (a_bytes < MG_M * (1024 + 16)) a_bytes = MG_M * (1024 + 16);
  a_bytes = (a_bytes + 127) & ~127;
  p.a_bytes = a_bytes;
  p.slab_stride = (H / 16) * 512;
  if (p.slab_stride < 32 * 512) p.slab_stride = 32 * 512;  // MLP proj slabs: 512 / 16 = 32 k-blocks
  int stages = (232448 - 128 - a_bytes) / p.slab_stride;
  if (stages > MG_MAX_STAGES) stages = MG_MAX_STAGES;
  if (stages < 2) return ERGM_ERR_UNSUPPORTED;
  const int tab_bytes = ((B * max_pages + B) * 4 + 127) & ~127;
  while (stages > 2 && 128 + a_bytes + stages * p.slab_stride + tab_bytes > 232448) --stages;
  if (128 + a_bytes + stages * p.slab_stride + tab_bytes > 232448) return ERGM_ERR_UNSUPPORTED;
  p.stages = stages;
  { const char* e = getenv("ERGM_MEGA_DBG"); p.dbg = e ? atoi(e) : 0; }
  const size_t smem = 128 + (size_t)a_bytes + (size_t)stages * p.slab_stride + (size_t)tab_bytes;
  const int grid = num_sms();
  cudaStream_t st = (cudaStream_t)stream;
  ERGM_CUDA_TRY(cudaMemsetAsync(sync_ctr, 0, sizeof(unsigned int), st));
  void* args[] = {(void*)&p};
  auto launch = [&](auto kern) -> int {
    ERGM_SET_SMEM_ATTR(kern, 232448);
    return (int)cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)grid), dim3(MG_THREADS), args, smem, st);
  };
  switch (H / 128) {
    case 1: return launch(decode_layers_kernel<1>);
    case 2: return launch(decode_layers_kernel<2>);
    case 4: return launch(decode_layers_kernel<4>);
    case 6: return launch(decode_layers_kernel<6>);
    case 8: return launch(decode_layers_kernel<8>);
  }
  return ERGM_ERR_UNSUPPORTED;
}

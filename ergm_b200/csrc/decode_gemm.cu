// Decode-step GEMMs (one new token per sequence, M = batch <= 64 rows): weight streaming.
//
// At M = 64 every projection of the block (model.py:222,244,218-219,263,265) and the LM head
// (model.py:698) is bound by reading its weight matrix once from HBM (247 MB of bf16
// weights per decode step vs 15.8 GFLOP), so the tcgen05 tile kernel of gemm_sm100.cu — 128-row
// tiles, a handful of CTAs per small-N problem — is the wrong tool.  Here:
//   * weights are re-packed once per weight version into "slabs": 16 output columns x all K, stored
//     in mma.m16n8k16 B-fragment order (one 512-byte block per 16x16 (k, n) tile, 16 bytes per lane),
//     so a CTA fetches a slab (or a K-range of it) with ONE contiguous cp.async.bulk and feeds the
//     tensor cores with conflict-free 8-byte shared loads, no ldmatrix / transposes;
//   * every SM streams: LayerNorm-fused problems (QKV, MLP up-projection, LM head) give each CTA
//     whole slabs (a ring of slabs per CTA for the 3142-slab LM head); residual problems (attention /
//     MLP output projections, fp32 residual stream) are additionally split along K and accumulated
//     with fp32 reductions straight into the residual stream, so 48 slabs still occupy > 500 CTAs;
//   * LayerNorm (model.py:298,318,332,578) is fused as the A-operand prologue: each CTA normalises the
//     (at most 64) rows itself from the fp32 residual stream (L2-resident, 196 KB) into bf16 smem;
//   * bias, gelu_new and the residual add are fused in the epilogue.
// The weight bulk loads are issued BEFORE anything that depends on the previous kernel's output.
#include "../../include/ergm_b200.h"
#include "common.cuh"
#include <cstdlib>

namespace ergm {

constexpr int DG_M = 64;          // rows per CTA tile (sequences)
constexpr int DG_MAX_STAGES = 6;
constexpr int DG_A_OFF = 128;     // smem: [0,128) mbarriers (ring stages + A), then A / reduction buffer, then the slab ring

ERGM_DEVINL void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

ERGM_DEVINL void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

ERGM_DEVINL void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                           uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------------------
// weight packing: logical W[K, N] (out = A @ W) -> [N/16 slabs][K/16][32 lanes][8 bf16]
//   lane l, element j:  k = kb*16 + (l%4)*2 + (j&1) + ((j&2) ? 8 : 0),  n = nb*16 + l/4 + ((j&4) ? 8 : 0)
// The LayerNorm affine parameters are folded in here, once per weight version:
//   LN(x) @ W + b = ((x - mean) * rstd) @ (diag(gamma) W) + (beta @ W + b)
// so the decode kernels only normalise (no gamma / beta loads on the critical path).
// ------------------------------------------------------------------------------------------
__global__ void dec_pack_kernel(const float* __restrict__ w, int64_t ld, int K, int N, int w_is_nk,
                                const float* __restrict__ gamma, __nv_bfloat16* __restrict__ packed) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int KB = K / 16, NB = (N + 15) / 16;
  if (idx >= (int64_t)NB * KB * 32) return;
  const int lane = (int)(idx & 31);
  const int64_t t = idx >> 5;
  const int kb = (int)(t % KB), nb = (int)(t / KB);
  __align__(16) __nv_bfloat16 v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = kb * 16 + (lane & 3) * 2 + (j & 1) + ((j & 2) ? 8 : 0);
    const int n = nb * 16 + (lane >> 2) + ((j & 4) ? 8 : 0);
    float x = 0.f;
    if (n < N) {
      x = w_is_nk ? w[(int64_t)n * ld + k] : w[(int64_t)k * ld + n];
      if (gamma) x *= gamma[k];
    }
    v[j] = __float2bfloat16_rn(x);
  }
  reinterpret_cast<uint4*>(packed)[idx] = *reinterpret_cast<const uint4*>(v);
}

// bias_out[n] = bias_in[n] + sum_k beta[k] * W[k, n]   (fixed summation order: deterministic)
__global__ void dec_fold_bias_kernel(const float* __restrict__ w, int64_t ld, int K, int N, int w_is_nk,
                                     const float* __restrict__ beta, const float* __restrict__ bias_in,
                                     float* __restrict__ bias_out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  if (beta) {
    if (w_is_nk) {
      const float4* row = reinterpret_cast<const float4*>(w + (int64_t)n * ld);
      for (int k4 = 0; k4 < K / 4; ++k4) {
        const float4 a = __ldg(row + k4), b = __ldg(reinterpret_cast<const float4*>(beta) + k4);
        s += (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
      }
    } else {
      for (int k = 0; k < K; ++k) s += __ldg(beta + k) * w[(int64_t)k * ld + n];
    }
  }
  bias_out[n] = s + (bias_in ? bias_in[n] : 0.f);
}

// ------------------------------------------------------------------------------------------
struct DecGemmParams {
  const float* x;               // LN mode: fp32 residual stream [M, lda]
  const __nv_bfloat16* a;       // direct mode: bf16 [M, lda]
  int64_t lda;                  // leading dimension of x / a (elements)
  float eps;
  const __nv_bfloat16* w;       // packed
  int K, N, NB;                 // NB = slabs
  int KBc;                      // k16 blocks per CTA k-range (K/16 / gridDim.y)
  int stages;
  const float* bias;
  void* out;
  int64_t ldo;
  int out_mode;                 // 0: bf16 store, 1: fp32 store, 2: fp32 reduction (+=)
  int gelu;
  int M;
  int cluster;                  // LN mode: CTAs per cluster sharing the LayerNorm work (1, 2 or 4)
};

// NV = float4 per lane per row of the LayerNorm prologue (K / 128); 0 = direct bf16 A operand.
// KS = k16 steps per warp.  Warps = 2 row halves x KW k-ways (KW = warps / 2, KS = KBc / KW): a warp
// keeps its A fragments (32 rows x KS*16 k) in REGISTERS for the whole kernel, so the per-slab shared
// memory traffic is one read of the 16-column weight slab per row half plus the cross-warp reduction
// of the [64, 16] partial tiles - the first version re-read all of A from smem for every slab and was
// bound by the shared-memory pipe (mio_throttle), 3 us per LM-head slab.
// LN mode runs as clusters of CS CTAs: every CTA of a cluster needs the same normalised [64, K] bf16
// operand, so CTA r normalises rows [r*64/CS, (r+1)*64/CS) only and writes them into the shared
// memory of all CS CTAs (DSMEM stores).
template <int NV, int KS>
__global__ void __launch_bounds__(NV > 0 ? 512 : 256) dec_gemm_kernel(const DecGemmParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Kc = p.KBc * 16;
  const int a_stride = Kc * 2 + 16;  // bytes; +16 keeps ldmatrix rows on distinct banks
  const uint32_t slab_bytes = (uint32_t)p.KBc * 512u;
  const uint32_t bars = smem_u32(smem);
  const uint32_t abar = bars + 8 * DG_MAX_STAGES;
  unsigned char* a_sm = smem + DG_A_OFF;
  float* red = reinterpret_cast<float*>(a_sm);  // aliases A once the fragments are in registers
  const int ring_off = (DG_A_OFF + max(DG_M * a_stride, (int)(blockDim.x >> 5) * 2048) + 127) & ~127;
  unsigned char* ring = smem + ring_off;
  const int ks = blockIdx.y;
  const int KB = p.K / 16;

  pdl_launch_dependents();
  if (NV > 0 && p.cluster > 1) cluster_arrive();  // "this CTA is running": peers may write its smem after the wait
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) mbar_init(bars + 8 * s, 1);
    mbar_init(abar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  // ---- weights first: they do not depend on the previous kernel ----
  if (tid == 0) {
    int slab = blockIdx.x;
    for (int s = 0; s < p.stages && slab < p.NB; ++s, slab += gridDim.x) {
      mbar_expect_tx(bars + 8 * s, slab_bytes);
      bulk_g2s(smem_u32(ring + (size_t)s * slab_bytes), p.w + ((int64_t)slab * KB + (int64_t)ks * p.KBc) * 256,
               slab_bytes, bars + 8 * s);
    }
  }
  pdl_wait();  // everything below reads the previous kernels' outputs
  // ---- A operand ----
  if constexpr (NV > 0) {
    const int CS = p.cluster;
    const uint32_t rank = CS > 1 ? cluster_ctarank() : 0u;
    const int rows_per_cta = DG_M / CS;
    const int rows_per_warp = rows_per_cta >> 4;  // 16 warps
    const float invH = 1.f / (float)(NV * 128);
    const uint32_t a_sm32 = smem_u32(a_sm);
    bool waited = CS == 1;
    for (int rr = 0; rr < rows_per_warp; rr += 2) {
      const int row0 = (int)rank * rows_per_cta + warp * rows_per_warp + rr;
      const int nq = rows_per_warp - rr >= 2 ? 2 : 1;
      float4 v[2][NV];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int row = row0 + q;
#pragma unroll
        for (int i = 0; i < NV; ++i)
          v[q][i] = (q < nq && row < p.M) ? reinterpret_cast<const float4*>(p.x + (int64_t)row * p.lda)[lane + 32 * i]
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float mean[2], rstd[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) s += (v[q][i].x + v[q][i].y) + (v[q][i].z + v[q][i].w);
        mean[q] = warp_sum(s) * invH;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const float dx = v[q][i].x - mean[q], dy = v[q][i].y - mean[q], dz = v[q][i].z - mean[q],
                      dw = v[q][i].w - mean[q];
          ss += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
        rstd[q] = rsqrtf(warp_sum(ss) * invH + p.eps);
      }
      if (!waited) { cluster_wait(); waited = true; }  // all CTAs of the cluster are resident now
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (q >= nq) break;
        const int row = row0 + q;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          uint2 o = make_uint2(0u, 0u);
          if (row < p.M)
            o = make_uint2(pack_bf16x2((v[q][i].x - mean[q]) * rstd[q], (v[q][i].y - mean[q]) * rstd[q]),
                           pack_bf16x2((v[q][i].z - mean[q]) * rstd[q], (v[q][i].w - mean[q]) * rstd[q]));
          const uint32_t off = (uint32_t)row * a_stride + (uint32_t)(lane + 32 * i) * 8;
          if (CS == 1) {
            *reinterpret_cast<uint2*>(a_sm + off) = o;
          } else {
            for (int c = 0; c < CS; ++c) st_cluster_v2(mapa_cluster(a_sm32 + off, (uint32_t)c), o.x, o.y);
          }
        }
      }
    }
    if (CS > 1) {
      if (!waited) cluster_wait();
      cluster_arrive();  // release: my rows are written everywhere
      cluster_wait();    // acquire: everybody's rows are here
    } else {
      __syncthreads();
    }
  } else {
    // one bulk copy per row (Kc * 2 bytes), all in flight at once
    const uint32_t row_bytes = (uint32_t)Kc * 2u;
    if (warp == 0) {
      if (lane == 0) mbar_expect_tx(abar, row_bytes * (uint32_t)p.M);
      __syncwarp();
      const __nv_bfloat16* src = p.a + (int64_t)ks * Kc;
      for (int r = lane; r < p.M; r += 32)
        bulk_g2s(smem_u32(a_sm + (size_t)r * a_stride), src + (int64_t)r * p.lda, row_bytes, abar);
    } else {
      const int chunks = Kc / 8;
      for (int idx = tid - 32; idx < (DG_M - p.M) * chunks; idx += blockDim.x - 32) {
        const int r = p.M + idx / chunks, c = idx % chunks;
        *reinterpret_cast<uint4*>(a_sm + (size_t)r * a_stride + c * 16) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    __syncthreads();
    mbar_wait(abar, 0);
  }

  // ---- A fragments -> registers (once) ----
  const int mh = warp & 1, kw = warp >> 1;
  uint32_t afr[2][KS][4];
  {
    const uint32_t a_base = smem_u32(a_sm) + (uint32_t)(mh * 32 + (lane & 7) + ((lane >> 3) & 1) * 8) * a_stride +
                            (uint32_t)(lane >> 4) * 16 + (uint32_t)(kw * KS) * 32;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int k = 0; k < KS; ++k)
        ldmatrix_x4(a_base + (uint32_t)(mi * 16) * a_stride + k * 32, afr[mi][k][0], afr[mi][k][1], afr[mi][k][2], afr[mi][k][3]);
  }
  __syncthreads();  // A region is free: it becomes the cross-warp reduction buffer

  // ---- main loop over this CTA's slabs ----
  const int half_threads = blockDim.x >> 1;
  const int KW = blockDim.x >> 6;
  const int rmh = tid >= half_threads ? 1 : 0;
  int it = 0;
  for (int slab = blockIdx.x; slab < p.NB; slab += gridDim.x, ++it) {
    const int stage = it % p.stages;
    // this thread's epilogue bias (2 adjacent columns per owned element pair): fetched now, used after
    // the reduction, so its L2 / DRAM latency hides behind the slab wait and the MMAs
    float2 bias_v[NV > 0 ? 1 : 2];
#pragma unroll
    for (int j = 0; j < (NV > 0 ? 1 : 2); ++j) {
      const int e2 = tid - rmh * half_threads + j * half_threads;
      const int col = slab * 16 + ((e2 >> 6) & 1) * 8 + ((e2 >> 1) & 3) * 2;
      bias_v[j] = make_float2(0.f, 0.f);
      if (p.bias && (p.out_mode != 2 || ks == 0)) {
        if (col < p.N) bias_v[j].x = __ldg(p.bias + col);
        if (col + 1 < p.N) bias_v[j].y = __ldg(p.bias + col + 1);
      }
    }
    mbar_wait(bars + 8 * stage, (uint32_t)((it / p.stages) & 1));
    const unsigned char* ws = ring + (size_t)stage * slab_bytes + (size_t)(kw * KS) * 512 + lane * 16;
    float acc[2][2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const uint4 b = *reinterpret_cast<const uint4*>(ws + (size_t)k * 512);
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        mma_16816(acc[mi][0], afr[mi][k][0], afr[mi][k][1], afr[mi][k][2], afr[mi][k][3], b.x, b.y);
        mma_16816(acc[mi][1], afr[mi][k][0], afr[mi][k][1], afr[mi][k][2], afr[mi][k][3], b.z, b.w);
      }
    }
    // partial [32 rows x 16 cols] of this warp -> red[warp][q][lane] (float4 q = mi*2 + n8)
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int n8 = 0; n8 < 2; ++n8)
        *reinterpret_cast<float4*>(red + ((warp * 4 + mi * 2 + n8) * 32 + lane) * 4) =
            make_float4(acc[mi][n8][0], acc[mi][n8][1], acc[mi][n8][2], acc[mi][n8][3]);
    __syncthreads();
    // reduce over the KW k-ways and run the epilogue: every thread owns 2 adjacent columns of one row
#pragma unroll
    for (int j = 0; j < (NV > 0 ? 1 : 2); ++j) {
      const int e2 = tid - rmh * half_threads + j * half_threads;
      const int q = e2 >> 6, l2 = (e2 >> 1) & 31, sub = (e2 & 1) * 2;
      float v0 = 0.f, v1 = 0.f;
      for (int w = 0; w < KW; ++w) {
        const float2 t2 = *reinterpret_cast<const float2*>(red + (((w * 2 + rmh) * 4 + q) * 32 + l2) * 4 + sub);
        v0 += t2.x; v1 += t2.y;
      }
      const int row = rmh * 32 + (q >> 1) * 16 + (l2 >> 2) + (sub ? 8 : 0);
      const int col = slab * 16 + (q & 1) * 8 + (l2 & 3) * 2;
      const bool c0 = col < p.N, c1 = col + 1 < p.N;
      if (row >= p.M || !c0) continue;
      v0 += bias_v[j].x; v1 += bias_v[j].y;
      if (p.gelu) { v0 = gelu_new<false>(v0); v1 = gelu_new<false>(v1); }
      if (p.out_mode == 0) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t)row * p.ldo + col;
        if (c1) *reinterpret_cast<uint32_t*>(o) = pack_bf16x2(v0, v1);
        else *o = __float2bfloat16_rn(v0);
      } else if (p.out_mode == 1) {
        float* o = reinterpret_cast<float*>(p.out) + (int64_t)row * p.ldo + col;
        if (c1) *reinterpret_cast<float2*>(o) = make_float2(v0, v1);
        else *o = v0;
      } else {
        float* o = reinterpret_cast<float*>(p.out) + (int64_t)row * p.ldo + col;
        atomicAdd(o, v0);
        if (c1) atomicAdd(o + 1, v1);
      }
    }
    if (slab + (int)gridDim.x < p.NB) __syncthreads();  // `red` and this ring stage are free again
    const int next = slab + p.stages * gridDim.x;
    if (tid == 0 && next < p.NB) {
      mbar_expect_tx(bars + 8 * stage, slab_bytes);
      bulk_g2s(smem_u32(ring + (size_t)stage * slab_bytes), p.w + ((int64_t)next * KB + (int64_t)ks * p.KBc) * 256,
               slab_bytes, bars + 8 * stage);
    }
  }
}

}  // namespace ergm

using namespace ergm;

// LN-mode launch geometry: the largest cluster size whose grid can be (almost) fully co-resident
// (cudaOccupancyMaxActiveClusters), probed once per kernel instance.
template <int NV, int KS>
static int ln_cluster_and_ctas(int smem, int want_ctas, int* cluster_out) {
  static int cached_cluster = 0, cached_max = 0;
  if (cached_cluster == 0) {
    int best_c = 1, best_n = num_sms();
    const char* e = getenv("ERGM_DEC_CLUSTER");
    const int cmax = e ? atoi(e) : 2;  // measured: 2 beats 1 and 4 (profiles/r1_decode.md)
    for (int c = cmax; c >= 2; c >>= 1) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(num_sms() / c * c)); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = (size_t)smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (unsigned)c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int ncl = 0;
      if (cudaOccupancyMaxActiveClusters(&ncl, dec_gemm_kernel<NV, KS>, &cfg) == cudaSuccess && ncl * c >= 96) {
        best_c = c; best_n = ncl * c;
        break;
      }
      cudaGetLastError();
    }
    cached_cluster = best_c; cached_max = best_n;
  }
  *cluster_out = cached_cluster;
  int n = want_ctas < cached_max ? want_ctas : cached_max;
  n = n / cached_cluster * cached_cluster;
  return n < cached_cluster ? cached_cluster : n;
}

template <int NV, int KS>
static int launch_dec_gemm(DecGemmParams& p, dim3 grid, cudaStream_t stream) {
  ERGM_SET_SMEM_ATTR((dec_gemm_kernel<NV, KS>), 232448);
  const int threads = NV > 0 ? 512 : 256;
  const int a_bytes = DG_M * (p.KBc * 32 + 16);
  const int red_bytes = (threads / 32) * 2048;
  const int ring_off = (DG_A_OFF + (a_bytes > red_bytes ? a_bytes : red_bytes) + 127) & ~127;
  const int slab_bytes = p.KBc * 512;
  p.cluster = 1;
  if (NV > 0) {
    int cl = 1;
    const int worst = ring_off + DG_MAX_STAGES * slab_bytes;
    const int n = ln_cluster_and_ctas<NV, KS>(worst > 232448 ? 232448 : worst, p.NB < num_sms() ? p.NB : num_sms(), &cl);
    p.cluster = cl;
    grid = dim3((unsigned)n, 1);
  }
  const int per_cta = (p.NB + (int)grid.x - 1) / (int)grid.x;
  int stages = (232448 - ring_off) / slab_bytes;
  if (stages > DG_MAX_STAGES) stages = DG_MAX_STAGES;
  if (stages > per_cta) stages = per_cta;
  if (stages < 1) return ERGM_ERR_UNSUPPORTED;
  p.stages = stages;
  const int smem = ring_off + stages * slab_bytes;
  ERGM_CUDA_TRY(launch_pdl(dec_gemm_kernel<NV, KS>, grid, dim3((unsigned)threads), (size_t)smem, stream, p.cluster, p));
  return ERGM_OK;
}

extern "C" int ergm_dec_pack_weight(const float* w_f32, int64_t ld, int K, int N, int w_is_nk, const float* gamma,
                                    const float* beta, const float* bias_in, void* packed, float* bias_out,
                                    void* stream) {
  if (!w_f32 || !packed || K <= 0 || N <= 0 || K % 16) return ERGM_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(packed) & 15) return ERGM_ERR_ARG;
  if ((beta || bias_in) && !bias_out) return ERGM_ERR_ARG;
  if (w_is_nk && beta && (ld % 4 || K % 4 || (reinterpret_cast<uintptr_t>(w_f32) & 15) ||
                          (reinterpret_cast<uintptr_t>(beta) & 15))) return ERGM_ERR_ARG;
  const int64_t n = (int64_t)((N + 15) / 16) * (K / 16) * 32;
  dec_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      w_f32, ld, K, N, w_is_nk, gamma, reinterpret_cast<__nv_bfloat16*>(packed));
  if (bias_out)
    dec_fold_bias_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(w_f32, ld, K, N, w_is_nk, beta, bias_in,
                                                                            bias_out);
  return (int)cudaGetLastError();
}

extern "C" int ergm_dec_gemm(const float* x_f32, const void* a_bf16, int64_t lda, float eps, const void* w_packed,
                             int K, int N, const float* bias, void* out, int64_t ldo, int out_mode, int gelu,
                             int M, void* stream) {
  if ((!x_f32) == (!a_bf16) || !w_packed || !out || K <= 0 || N <= 0 || M <= 0) return ERGM_ERR_ARG;
  if (M > DG_M) return ERGM_ERR_UNSUPPORTED;
  if (out_mode < 0 || out_mode > 2 || (gelu && out_mode == 2)) return ERGM_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(w_packed) & 15) return ERGM_ERR_ARG;
  DecGemmParams p{};
  p.x = x_f32; p.a = reinterpret_cast<const __nv_bfloat16*>(a_bf16); p.lda = lda; p.eps = eps;
  p.w = reinterpret_cast<const __nv_bfloat16*>(w_packed);
  p.K = K; p.N = N; p.NB = (N + 15) / 16;
  p.bias = bias; p.out = out; p.ldo = ldo; p.out_mode = out_mode; p.gelu = gelu; p.M = M;
  cudaStream_t st = (cudaStream_t)stream;
  if (x_f32) {
    // LayerNorm-fused: whole-K slabs, 16 warps = 2 row halves x 8 k-ways, KS = K / 128 k16-steps per warp
    if (lda % 4 || (reinterpret_cast<uintptr_t>(x_f32) & 15) || out_mode == 2) return ERGM_ERR_ARG;
    p.KBc = K / 16;
    const dim3 grid(1, 1);
    switch (K) {
      case 128: return launch_dec_gemm<1, 1>(p, grid, st);
      case 256: return launch_dec_gemm<2, 2>(p, grid, st);
      case 512: return launch_dec_gemm<4, 4>(p, grid, st);
      case 768: return launch_dec_gemm<6, 6>(p, grid, st);
      case 1024: return launch_dec_gemm<8, 8>(p, grid, st);
    }
    return ERGM_ERR_UNSUPPORTED;
  }
  // residual "+=" problems: K split into ranges of 256 / 128 / 64 so that small-N problems still fill
  // the GPU; 8 warps = 2 row halves x 4 k-ways
  if (out_mode != 2) return ERGM_ERR_UNSUPPORTED;
  if (lda % 8 || (reinterpret_cast<uintptr_t>(a_bf16) & 15) || K % 64) return ERGM_ERR_ARG;
  static int kc_pref = 0;
  if (!kc_pref) { const char* e = getenv("ERGM_DEC_KC"); kc_pref = e ? atoi(e) : 512; }  // measured: MLP proj 8.4 -> 4.95 us with 512-wide K ranges (6 instead of 12 splits)
  const int Kc = (kc_pref == 512 && K % 512 == 0 && K >= 2048) ? 512 : (K % 256 == 0 ? 256 : (K % 128 == 0 ? 128 : 64));
  p.KBc = Kc / 16;
  const dim3 grid((unsigned)p.NB, (unsigned)(K / Kc));
  if (grid.y > 65535) return ERGM_ERR_UNSUPPORTED;
  switch (Kc) {
    case 512: return launch_dec_gemm<0, 8>(p, grid, st);
    case 256: return launch_dec_gemm<0, 4>(p, grid, st);
    case 128: return launch_dec_gemm<0, 2>(p, grid, st);
    default: return launch_dec_gemm<0, 1>(p, grid, st);
  }
}

// ergm_gemm_bf16 — persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = epilogue( A[M,K] * B[K,N] ),  bf16 operands, fp32 accumulation in TMEM.
//
// Replaces every dense contraction of the reference forward (transformers Conv1D =
// addmm(bias, x, W[K,N]) at /root/reference/src/model.py:218,219,222,244,263,265 and the
// tied nn.Linear lm_head at model.py:698) plus the dgrad / wgrad products autograd derives
// from them.  Operand "majorness" is a run-time flag so one kernel serves
//   forward Conv1D        A = x[M,K]   (K-major)   B = W[K,N]   (MN-major)
//   lm_head / dgrad       A = .[M,K]   (K-major)   B = W[N,K]   (K-major)
//   wgrad                 A = x[Mr,K]^T (MN-major) B = dY[Mr,N] (MN-major)
//
// Roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one lane), warp 2 = TMEM
// allocator, warps 4..11 = epilogue (two warps per TMEM lane quarter, each half of the
// columns).  Three pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty double
// buffer (MMA <-> epilogue), static persistent tile schedule.
#include "../../include/ergm_b200.h"
#include "common.cuh"
#include "dropout.cuh"

namespace ergm {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_EPI_WARPS = 8;
constexpr int SMEM_BUDGET = 192 * 1024;
constexpr int EPI_STAGING = NUM_EPI_WARPS * 4096;  // per-warp 32x32 fp32 transpose tiles

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (SMEM_BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;  // power of 2
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGING + 1024 /*align*/ + 256 /*barriers*/;
};

struct GemmParams {
  void* d;
  const float* bias;
  const float* residual;
  void* preact;
  float* colsum;
  int64_t ldd, ldr;
  int M, N, K;
  int a_mn, b_mn;
  int d_f32;
  int epi;
  int split_k;
  int m_tiles, n_tiles, kb_total, kb_per_split;
  float dropout_p;
  DropoutSite site;
  const int* dyn;  // nullable device int: the run-time extent of M (dyn_dim 1) or K (dyn_dim 2), <= the static one
  int dyn_dim;
  int pdl_trigger;  // release the stream's next kernel early (training-size problems; the M <= 128 decode-step
                    // GEMMs keep the implicit trigger at exit: an early-resident top-k sampler cost 110 us per step)
};

// Run-time problem extent (label-sparse LM head, packed batches): the host sizes the launch, the tensor maps and
// the workspaces for the static capacity; the kernel re-derives its tile schedule from a device-side count, so a
// CUDA graph captured once serves every batch.  dyn_dim 1: rows of A / D beyond the count are never stored (the
// last tile may LOAD stale rows: they only feed accumulator rows that are discarded).  dyn_dim 2: the reduction
// stops at the count rounded up to BK: the caller keeps the operand rows in [count, roundup(count, 128)) zero.
// With ERGM_EPI_ATOMIC and dyn_dim 1 the K split is chosen here so that the (few) row tiles still fill the GPU.
ERGM_DEVINL void apply_dyn(GemmParams& p, int tile_m, int slots) {
  if (!p.dyn) return;
  int v = *p.dyn;
  v = v < 0 ? 0 : v;
  if (p.dyn_dim == 1) {
    p.M = v < p.M ? v : p.M;
    p.m_tiles = (p.M + tile_m - 1) / tile_m;
    if ((p.epi & ERGM_EPI_ATOMIC) && p.m_tiles > 0) {
      int sk = slots / (p.m_tiles * p.n_tiles);
      const int max_sk = p.kb_total / 8 > 0 ? p.kb_total / 8 : 1;  // at least 8 k-blocks per split
      sk = sk < 1 ? 1 : (sk > max_sk ? max_sk : sk);
      p.kb_per_split = (p.kb_total + sk - 1) / sk;
      p.split_k = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
    }
  } else if (p.dyn_dim == 2) {
    p.K = v < p.K ? v : p.K;
    p.kb_total = (p.K + BK - 1) / BK;
    int sk = p.split_k < p.kb_total ? p.split_k : p.kb_total;
    sk = sk < 1 ? 1 : sk;
    p.kb_per_split = p.kb_total > 0 ? (p.kb_total + sk - 1) / sk : 1;
    p.split_k = p.kb_total > 0 ? (p.kb_total + p.kb_per_split - 1) / p.kb_per_split : 0;
  }
}


struct EpiFlags {
  bool has_bias, do_gelu, do_res, do_atomic, do_drop, do_pre, exact, do_gelu_grad;
  float keep_scale;
  DropoutSite site;
};

ERGM_DEVINL EpiFlags make_epi_flags(const GemmParams& p) {
  EpiFlags e;
  e.has_bias = (p.epi & ERGM_EPI_BIAS) != 0;
  e.do_gelu = (p.epi & ERGM_EPI_GELU) != 0;
  e.do_res = (p.epi & ERGM_EPI_RESIDUAL) != 0;
  e.do_atomic = (p.epi & ERGM_EPI_ATOMIC) != 0;
  e.do_drop = (p.epi & ERGM_EPI_DROPOUT) != 0;
  e.do_pre = (p.epi & ERGM_EPI_PREACT) != 0;
  e.exact = (p.epi & ERGM_EPI_EXACT) != 0;
  e.do_gelu_grad = (p.epi & ERGM_EPI_GELU_GRAD) != 0;
  e.site = p.site.resolved();
  e.keep_scale = e.do_drop ? e.site.keep_scale() : 1.0f;
  return e;
}

// Scalar tail path of the epilogue: the lane's 4 columns straddle N (only the last column chunk of a
// matrix whose N is not a multiple of 4, e.g. V = 50260 ... never on the hot path).  Kept out of line
// so the unrolled fast path stays small enough for the instruction cache.
__device__ __noinline__ void epilogue_tail(void* d, void* preact, const float* residual, int64_t ldd, int64_t ldr,
                                           int N, int d_f32, int epi, uint32_t keep, float keep_scale, int row,
                                           int col, float4 x) {
  const float xv[4] = {x.x, x.y, x.z, x.w};
  for (int j = 0; j < 4; ++j) {
    if (col + j >= N) break;
    const int64_t off = (int64_t)row * ldd + col + j;
    float y = xv[j];
    if (epi & ERGM_EPI_PREACT) reinterpret_cast<__nv_bfloat16*>(preact)[off] = __float2bfloat16_rn(y);
    if (epi & ERGM_EPI_GELU_GRAD) y *= gelu_new_grad<true>(__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(preact)[off]));
    if (epi & ERGM_EPI_GELU) y = gelu_new<true>(y);
    if (epi & ERGM_EPI_DROPOUT) y = ((keep >> j) & 1u) ? y * keep_scale : 0.f;
    if (epi & ERGM_EPI_RESIDUAL) y += residual[(int64_t)row * ldr + col + j];
    if (d_f32) {
      float* dp = reinterpret_cast<float*>(d) + off;
      if (epi & ERGM_EPI_ATOMIC) atomicAdd(dp, y); else *dp = y;
    } else {
      reinterpret_cast<__nv_bfloat16*>(d)[off] = __float2bfloat16_rn(y);
    }
  }
}

// Epilogue classes: which fused variants a kernel instantiation carries (keeps each instantiation's
// unrolled epilogue small enough for the instruction cache).
enum { EC_LINEAR = 0, EC_GELU = 1, EC_GELU_GRAD = 2, EC_EXACT = 3 };

// Fused epilogue of one 32-row x 32-column chunk of an accumulator tile, per warp.
//
// tcgen05.ld hands every thread one accumulator ROW (32 consecutive columns).  Storing from that
// layout makes each warp-wide 16-byte access touch 32 different 128-byte lines.  The chunk is
// therefore transposed through a per-warp, XOR-swizzled 4 KB shared-memory tile: afterwards lane l
// owns 4 consecutive columns (float4) of rows {4*it + l/8}, so one warp instruction reads / writes
// four fully coalesced 128-byte (fp32) or 64-byte (bf16) row segments.  The residual / saved
// pre-activation operands of all 8 row groups are fetched up front (before the transposition) so
// their latency overlaps instead of serialising.  Fused order:
// bias -> (store pre-activation) -> GELU' / GELU -> dropout -> + residual -> store / red.add.
template <int EC>
ERGM_DEVINL void epilogue_chunk(const GemmParams& p, const EpiFlags& ep, int row0, int col0, bool first_split,
                                const float (&v)[32], uint32_t stage_smem, int lane) {
  const int c4 = lane & 7;             // float4 column index inside the chunk
  const int rsub = lane >> 3;          // row inside each group of 4
  const int col = col0 + 4 * c4;       // first of this lane's 4 columns
  const bool col_full = col + 4 <= p.N;
  const bool col_any = col < p.N;
  // ---- operand prefetch ----
  float4 res[8];
  uint2 pre[8];
  if (ep.do_res && col_full) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = row0 + 4 * it + rsub;
      res[it] = row < p.M ? *reinterpret_cast<const float4*>(p.residual + (int64_t)row * p.ldr + col)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if ((EC == EC_GELU_GRAD || EC == EC_EXACT) && ep.do_gelu_grad && col_full) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = row0 + 4 * it + rsub;
      pre[it] = row < p.M ? *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.preact) +
                                                            (int64_t)row * p.ldd + col)
                          : make_uint2(0u, 0u);
    }
  }
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ep.has_bias && first_split && col_any) {
    if (col_full) {
      b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    } else {
      b4.x = __ldg(p.bias + col);
      if (col + 1 < p.N) b4.y = __ldg(p.bias + col + 1);
      if (col + 2 < p.N) b4.z = __ldg(p.bias + col + 2);
    }
  }
  // ---- transpose: row-per-thread -> (row group, float4 column) per lane ----
  __syncwarp();  // previous chunk's reads of the staging tile are done
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t addr = stage_smem + lane * 128 + ((uint32_t)(c ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * c]), "f"(v[4 * c + 1]),
                 "f"(v[4 * c + 2]), "f"(v[4 * c + 3])
                 : "memory");
  }
  __syncwarp();
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);  // GELU' mode with p.colsum: column sums of the stored (bf16) values
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = 4 * it + rsub;
    const int row = row0 + r;
    float4 x;
    {
      const uint32_t addr = stage_smem + r * 128 + ((uint32_t)(c4 ^ (r & 7)) << 4);
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(addr));
    }
    if (row >= p.M || !col_any) continue;
    if (!col_full) {
      const uint32_t keep = ep.do_drop ? ep.site.keep4((uint32_t)row, (uint32_t)col >> 2) : 0xfu;
      epilogue_tail(p.d, p.preact, p.residual, p.ldd, p.ldr, p.N, p.d_f32, p.epi, keep, ep.keep_scale, row, col,
                    make_float4(x.x + b4.x, x.y + b4.y, x.z + b4.z, x.w + b4.w));
      continue;
    }
    x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
    const int64_t off = (int64_t)row * p.ldd + col;
    if ((EC == EC_GELU || EC == EC_EXACT) && ep.do_pre)
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.preact) + off) =
          make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
    if ((EC == EC_GELU_GRAD || EC == EC_EXACT) && ep.do_gelu_grad) {  // x *= gelu_new'(u) (model.py:264 backward)
      const float2 f0 = unpack_bf16x2(pre[it].x), f1 = unpack_bf16x2(pre[it].y);
      constexpr bool kExact = (EC == EC_EXACT);
      x.x *= gelu_new_grad<kExact>(f0.x); x.y *= gelu_new_grad<kExact>(f0.y);
      x.z *= gelu_new_grad<kExact>(f1.x); x.w *= gelu_new_grad<kExact>(f1.y);
    }
    if ((EC == EC_GELU || EC == EC_EXACT) && ep.do_gelu) {
      constexpr bool kExact = (EC == EC_EXACT);
      x.x = gelu_new<kExact>(x.x); x.y = gelu_new<kExact>(x.y); x.z = gelu_new<kExact>(x.z); x.w = gelu_new<kExact>(x.w);
    }
    if (ep.do_drop) {
      const uint32_t keep = ep.site.keep4((uint32_t)row, (uint32_t)col >> 2);
      x.x = (keep & 1u) ? x.x * ep.keep_scale : 0.f;
      x.y = (keep & 2u) ? x.y * ep.keep_scale : 0.f;
      x.z = (keep & 4u) ? x.z * ep.keep_scale : 0.f;
      x.w = (keep & 8u) ? x.w * ep.keep_scale : 0.f;
    }
    if (ep.do_res) { x.x += res[it].x; x.y += res[it].y; x.z += res[it].z; x.w += res[it].w; }
    if (p.d_f32) {
      float* dp = reinterpret_cast<float*>(p.d) + off;
      if (ep.do_atomic)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dp), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w)
                     : "memory");
      else
        *reinterpret_cast<float4*>(dp) = x;
    } else {
      const uint2 pk = make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.d) + off) = pk;
      if ((EC == EC_GELU_GRAD) && p.colsum) {
        const float2 r0 = unpack_bf16x2(pk.x), r1 = unpack_bf16x2(pk.y);
        cs.x += r0.x; cs.y += r0.y; cs.z += r1.x; cs.w += r1.y;
      }
    }
  }
  if ((EC == EC_GELU_GRAD) && p.colsum && ep.do_gelu_grad) {
    // boundary slab of a run-time-M problem (packed batch): the lean epilogue only runs on whole 32-row slabs, the
    // rows below M of this one still belong to the bias gradient.  Warp-uniform branch (p.colsum, flags).
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
      cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
    }
    if (lane < 8 && col_full)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.colsum + col), "f"(cs.x), "f"(cs.y), "f"(cs.z),
                   "f"(cs.w)
                   : "memory");
  }
}


// ------------------------------------------------------------------------------------------
// Lean epilogues for the hot configurations.  ncu of the generic epilogue above (profiles/r1_gemm_epilogue.md):
// ~670 warp instructions per 32x32 chunk, 60 of them branches and ~150 index / predicate arithmetic for
// flags that are uniform per launch, and one exposed DRAM round trip per chunk for the residual operand;
// with two epilogue warps per scheduler that made a 128x256 tile's epilogue take 8 us against a 4.5 us
// mainloop.  The fast path is compiled per mode (no run-time flag tests), runs only on interior tiles
// (no bounds tests; edge tiles take the generic path), strength-reduces the addressing and prefetches
// the residual operand of chunk c+1 (and of the next tile's first chunk) while chunk c is processed.
// ------------------------------------------------------------------------------------------
enum { FM_NONE = 0, FM_BF16 = 1, FM_BF16_BIAS = 2, FM_F32 = 3, FM_F32_RES = 4, FM_GELU = 5, FM_ATOMIC = 6, FM_GELU_GRAD = 7 };

template <int FM>
ERGM_DEVINL void fast_res_prefetch(const GemmParams& p, int row0, int col0, int lane, float4 (&res)[8]) {
  if constexpr (FM == FM_F32_RES) {
    const float* rp = p.residual + (int64_t)(row0 + (lane >> 3)) * p.ldr + col0 + 4 * (lane & 7);
    const int64_t step = 4 * p.ldr;
#pragma unroll
    for (int it = 0; it < 8; ++it) res[it] = *reinterpret_cast<const float4*>(rp + it * step);
  }
  if constexpr (FM == FM_GELU_GRAD) {  // saved pre-activation u (bf16, same layout as D): 4 values = 8 bytes per lane
    const __nv_bfloat16* up = reinterpret_cast<const __nv_bfloat16*>(p.preact) + (int64_t)(row0 + (lane >> 3)) * p.ldd +
                              col0 + 4 * (lane & 7);
    const int64_t step = 4 * p.ldd;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const uint2 u = *reinterpret_cast<const uint2*>(up + it * step);
      res[it] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), 0.f, 0.f);
    }
  }
}

template <int FM>
ERGM_DEVINL void epilogue_chunk_fast(const GemmParams& p, const EpiFlags& ep, int row0, int col0,
                                     const float (&v)[32], const float4 (&res)[8], uint32_t stage_smem, int lane) {
  const int c4 = lane & 7, rsub = lane >> 3;
  const int col = col0 + 4 * c4;
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (FM == FM_BF16_BIAS || FM == FM_GELU) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
  if constexpr (FM == FM_F32_RES) {
    if (ep.has_bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
  }
  __syncwarp();  // previous chunk's reads of the staging tile are done
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t addr = stage_smem + lane * 128 + ((uint32_t)(c ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * c]), "f"(v[4 * c + 1]),
                 "f"(v[4 * c + 2]), "f"(v[4 * c + 3])
                 : "memory");
  }
  __syncwarp();
  const int64_t off0 = (int64_t)(row0 + rsub) * p.ldd + col;
  const int64_t dstep = 4 * p.ldd;
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);  // FM_GELU_GRAD: column sums of the stored values (bias gradient)
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = 4 * it + rsub;
    float4 x;
    {
      const uint32_t addr = stage_smem + r * 128 + ((uint32_t)(c4 ^ (r & 7)) << 4);
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(addr));
    }
    const int64_t off = off0 + it * dstep;
    if constexpr (FM == FM_BF16_BIAS || FM == FM_GELU || FM == FM_F32_RES) {
      x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
    }
    if constexpr (FM == FM_GELU) {
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.preact) + off) =
          make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
      x.x = gelu_new<false>(x.x); x.y = gelu_new<false>(x.y); x.z = gelu_new<false>(x.z); x.w = gelu_new<false>(x.w);
    }
    if constexpr (FM == FM_F32_RES) {
      if (ep.do_drop) {
        const uint32_t keep = ep.site.keep4((uint32_t)(row0 + r), (uint32_t)col >> 2);
        x.x = (keep & 1u) ? x.x * ep.keep_scale : 0.f;
        x.y = (keep & 2u) ? x.y * ep.keep_scale : 0.f;
        x.z = (keep & 4u) ? x.z * ep.keep_scale : 0.f;
        x.w = (keep & 8u) ? x.w * ep.keep_scale : 0.f;
      }
      x.x += res[it].x; x.y += res[it].y; x.z += res[it].z; x.w += res[it].w;
    }
    if constexpr (FM == FM_GELU_GRAD) {  // x *= gelu_new'(u)  (model.py:264 backward), then the bias-gradient column sum
      const float2 f0 = unpack_bf16x2(__float_as_uint(res[it].x)), f1 = unpack_bf16x2(__float_as_uint(res[it].y));
      x.x *= gelu_new_grad<false>(f0.x); x.y *= gelu_new_grad<false>(f0.y);
      x.z *= gelu_new_grad<false>(f1.x); x.w *= gelu_new_grad<false>(f1.y);
      const uint2 pk = make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.d) + off) = pk;
      const float2 r0 = unpack_bf16x2(pk.x), r1 = unpack_bf16x2(pk.y);
      cs.x += r0.x; cs.y += r0.y; cs.z += r1.x; cs.w += r1.y;
      continue;
    }
    if constexpr (FM == FM_F32 || FM == FM_F32_RES) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.d) + off) = x;
    } else if constexpr (FM == FM_ATOMIC) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<float*>(p.d) + off), "f"(x.x),
                   "f"(x.y), "f"(x.z), "f"(x.w)
                   : "memory");
    } else {
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.d) + off) =
          make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
    }
  }
  if constexpr (FM == FM_GELU_GRAD) {
    if (p.colsum) {  // lanes c4, c4+8, c4+16, c4+24 hold the same 4 columns of different rows
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
        cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
      }
      if (lane < 8)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.colsum + col), "f"(cs.x), "f"(cs.y), "f"(cs.z),
                     "f"(cs.w)
                     : "memory");
    }
  }
}

// Epilogue of one warp's share (32 rows x BN/2 columns) of an accumulator tile.  `pre_valid`: res_pre
// already holds the residual operand of the first chunk (prefetched before the accumulator was ready).
template <int BN, int EC, int FM>
ERGM_DEVINL void epilogue_tile(const GemmParams& p, const EpiFlags& ep, uint32_t tmem_tile, int row0, int n0, int half,
                               bool first_split, uint32_t stage_smem, int lane, float4 (&res_pre)[8], bool fast) {
  const int c_lo = half * (BN / 2), c_hi = (half + 1) * (BN / 2);
  if (FM != FM_NONE && fast) {
#pragma unroll 1
    for (int c = c_lo; c < c_hi; c += 32) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_tile + c, r);
      float4 res_cur[8];
      if constexpr (FM == FM_F32_RES || FM == FM_GELU_GRAD) {
#pragma unroll
        for (int it = 0; it < 8; ++it) res_cur[it] = res_pre[it];
        if (c + 32 < c_hi) fast_res_prefetch<FM>(p, row0, n0 + c + 32, lane, res_pre);
      }
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
      epilogue_chunk_fast<FM>(p, ep, row0, n0 + c, v, res_cur, stage_smem, lane);
    }
    return;
  }
#pragma unroll 1
  for (int c = c_lo; c < c_hi; c += 32) {
    const int col0 = n0 + c;
    if (col0 >= p.N) break;  // warp-uniform
    uint32_t r[32];
    tmem_ld_32x32b_x32(tmem_tile + c, r);
    tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    epilogue_chunk<EC>(p, ep, row0, col0, first_split, v, stage_smem, lane);
  }
}

template <int BN, int EC, int FM>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a,
                 const __grid_constant__ CUtensorMap tmap_b, const GemmParams p_in) {
  using Cfg = GemmCfg<BN>;
  GemmParams p = p_in;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned tiles
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_all = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = stage_all + EPI_STAGING;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), NUM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) touches no
  // global memory and may overlap the tail of the previous kernel of the stream; nothing below runs before that
  // kernel has completed.  Our own dependents may be scheduled as soon as every CTA of this grid got here.
  pdl_wait();
  if (p.pdl_trigger) pdl_launch_dependents();
  apply_dyn(p, BM, (int)gridDim.x);
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int tiles_mn = p.m_tiles * p.n_tiles;
  const int total_tiles = tiles_mn * p.split_k;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / tiles_mn;
        const int mn = tile - split * tiles_mn;
        const int m0 = (mn % p.m_tiles) * BM;
        const int n0 = (mn / p.m_tiles) * BN;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          const int k0 = kb * BK;
          if (!p.a_mn) {
            tma_load_2d(sa, &tmap_a, full_bar(stage), k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(sa + j * 8192, &tmap_a, full_bar(stage), m0 + 64 * j, k0);
          }
          if (!p.b_mn) {
            tma_load_2d(sb, &tmap_b, full_bar(stage), k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * 8192, &tmap_b, full_bar(stage), n0 + 64 * j, k0);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
      // K-major: 8-row groups 1024 B apart, K step of 16 elements = +32 B inside the atom.
      // MN-major: 64-wide MN chunks 8192 B apart (LBO), 8-deep K groups 1024 B apart (SBO),
      //           K step of 16 = +2048 B.
      const uint32_t a_lbo = p.a_mn ? 8192u : 16u, a_kstep = p.a_mn ? 2048u : 32u;
      const uint32_t b_lbo = p.b_mn ? 8192u : 16u, b_kstep = p.b_mn ? 2048u : 32u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / tiles_mn;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            const uint64_t adesc = make_smem_desc_sw128(sa + ks * a_kstep, a_lbo, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(sb + ks * b_kstep, b_lbo, 1024);
            umma_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue =====================
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = (warp - EPI_WARP0) >> 2;  // which half of the columns
    const EpiFlags ep = make_epi_flags(p);
    const uint32_t stage_smem = stage_all + (warp - EPI_WARP0) * 4096;
    int acc = 0;
    uint32_t acc_phase = 0;
    float4 res_pre[8];
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = tile / tiles_mn;
      const int mn = tile - split * tiles_mn;
      const int m0 = (mn % p.m_tiles) * BM;
      const int n0 = (mn / p.m_tiles) * BN;
      const int row0 = m0 + q * 32;
      const bool fast = FM != FM_NONE && row0 + 32 <= p.M && n0 + (half + 1) * (BN / 2) <= p.N;
      if (fast) fast_res_prefetch<FM>(p, row0, n0 + half * (BN / 2), lane, res_pre);  // independent of the MMAs
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      epilogue_tile<BN, EC, FM>(p, ep, tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN, row0, n0, half, split == 0,
                                stage_smem, lane, res_pre, fast);
      // hand the accumulator buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}


// ------------------------------------------------------------------------------------------
// CTA-pair variant: tcgen05.mma.cta_group::2, one 256 x BN output tile per 2-CTA cluster.
// Each CTA stages its own 128 x 64 slice of A and HALF of the B tile (BN/2 x 64), the pair's
// leader issues the MMAs for both tensor cores, accumulators land in each CTA's own TMEM
// (128 lanes x BN columns).  Per SM and k-block this moves 16 KB (A) + BN/2*128 B (B) instead of
// 16 KB + BN*128 B, which is what lifts the kernel off the L2->smem bandwidth / latency bound the
// single-CTA kernel sits on (ncu: tensor pipe ~40 % active at 128x256, profiles/r1_gemm_attn.md).
// Barriers: full[s] lives in the leader and counts the TMA bytes of BOTH CTAs; empty[s] and
// tmem_full[a] are signalled in both CTAs by a multicast tcgen05.commit; tmem_empty[a] lives in
// the leader and is arrived on remotely by the epilogue warps of both CTAs.
// ------------------------------------------------------------------------------------------
template <int BN>
struct Gemm2Cfg {
  static constexpr int A_BYTES = 128 * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (SMEM_BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGING + 1024 + 256;
};

template <int BN, int EC, int FM>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a,
                  const __grid_constant__ CUtensorMap tmap_b, const GemmParams p_in) {
  using Cfg = Gemm2Cfg<BN>;
  GemmParams p = p_in;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_all = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = stage_all + EPI_STAGING;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 2 * NUM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  pdl_wait();               // (see gemm_bf16_kernel) no global memory access above this line
  if (p.pdl_trigger) pdl_launch_dependents();
  apply_dyn(p, 256, (int)(gridDim.x >> 1));
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int tiles_mn = p.m_tiles * p.n_tiles;  // m_tiles counts 256-row tiles here
  const int total_tiles = tiles_mn * p.split_k;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int split = tile / tiles_mn;
        const int mn = tile - split * tiles_mn;
        const int m0 = (mn % p.m_tiles) * 256 + (int)rank * 128;
        const int n0 = (mn / p.m_tiles) * BN + (int)rank * (BN / 2);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint32_t fb = mapa_cluster(full_bar(stage), 0);  // the leader's barrier
          if (leader) mbar_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
          const int k0 = kb * BK;
          if (!p.a_mn) {
            tma_load_2d_2sm(sa, &tmap_a, fb, k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d_2sm(sa + j * 8192, &tmap_a, fb, m0 + 64 * j, k0);
          }
          if (!p.b_mn) {
            tma_load_2d_2sm(sb, &tmap_b, fb, k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 128; ++j) tma_load_2d_2sm(sb + j * 8192, &tmap_b, fb, n0 + 64 * j, k0);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && leader) {
      const uint32_t idesc = make_idesc_bf16(256, BN, p.a_mn, p.b_mn);
      const uint32_t a_lbo = p.a_mn ? 8192u : 16u, a_kstep = p.a_mn ? 2048u : 32u;
      const uint32_t b_lbo = p.b_mn ? 8192u : 16u, b_kstep = p.b_mn ? 2048u : 32u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int split = tile / tiles_mn;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            const uint64_t adesc = make_smem_desc_sw128(sa + ks * a_kstep, a_lbo, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(sb + ks * b_kstep, b_lbo, 1024);
            umma_ss_2sm(d_tmem, adesc, bdesc, idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
          }
          umma_commit_2sm(empty_bar(stage), 3);  // frees the slot in BOTH CTAs
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(tfull_bar(acc), 3);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int q = warp & 3;
    const int half = (warp - EPI_WARP0) >> 2;
    const EpiFlags ep = make_epi_flags(p);
    const uint32_t stage_smem = stage_all + (warp - EPI_WARP0) * 4096;
    int acc = 0;
    uint32_t acc_phase = 0;
    float4 res_pre[8];
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const int split = tile / tiles_mn;
      const int mn = tile - split * tiles_mn;
      const int m0 = (mn % p.m_tiles) * 256 + (int)rank * 128;
      const int n0 = (mn / p.m_tiles) * BN;
      const int row0 = m0 + q * 32;
      const bool fast = FM != FM_NONE && row0 + 32 <= p.M && n0 + (half + 1) * (BN / 2) <= p.N;
      if (fast) fast_res_prefetch<FM>(p, row0, n0 + half * (BN / 2), lane, res_pre);  // independent of the MMAs
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      epilogue_tile<BN, EC, FM>(p, ep, tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN, row0, n0, half, split == 0,
                                stage_smem, lane, res_pre, fast);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_cluster(tempty_bar(acc), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync();
  if (warp == 2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
}

template <int BN, int EC, int FM>
static int launch_gemm2(const ergm_gemm_args* a, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN>;
  CUtensorMap ta, tb;
  int rc;
  if (a->a_major == ERGM_MAJOR_K)
    rc = encode_tmap_2d(&ta, a->a, 2, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda * 2, BK, 128);
  else
    rc = encode_tmap_2d(&ta, a->a, 2, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda * 2, 64, BK);
  if (rc) return rc;
  if (a->b_major == ERGM_MAJOR_K)
    rc = encode_tmap_2d(&tb, a->b, 2, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb * 2, BK, BN / 2);
  else
    rc = encode_tmap_2d(&tb, a->b, 2, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb * 2, 64, BK);
  if (rc) return rc;
  GemmParams p;
  p.d = a->d; p.bias = a->bias; p.residual = a->residual; p.preact = a->preact; p.colsum = a->colsum;
  p.ldd = a->ldd; p.ldr = a->ldr;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.a_mn = a->a_major == ERGM_MAJOR_MN; p.b_mn = a->b_major == ERGM_MAJOR_MN;
  p.d_f32 = a->d_dtype == ERGM_DT_F32;
  p.epi = a->epilogue;
  p.split_k = a->split_k < 1 ? 1 : a->split_k;
  p.m_tiles = (a->M + 255) / 256;
  p.n_tiles = (a->N + BN - 1) / BN;
  p.kb_total = (a->K + BK - 1) / BK;
  if (p.split_k > p.kb_total) p.split_k = p.kb_total;
  p.kb_per_split = (p.kb_total + p.split_k - 1) / p.split_k;
  p.split_k = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.dropout_p = a->dropout_p;
  p.site = make_site(a->seed, a->offset, a->dropout_p, (uint32_t)a->N);
  p.dyn = a->dyn_count; p.dyn_dim = a->dyn_count ? a->dyn_dim : 0;
  p.pdl_trigger = a->M > 128;
  ERGM_SET_SMEM_ATTR((gemm2_bf16_kernel<BN, EC, FM>), Cfg::SMEM_BYTES);
  const int total = p.m_tiles * p.n_tiles * p.split_k;
  const int max_clusters = num_sms() / 2;
  const int clusters = total < max_clusters ? total : max_clusters;
  return (int)launch_pdl(gemm2_bf16_kernel<BN, EC, FM>, dim3(2 * clusters), dim3(GEMM_THREADS), (size_t)Cfg::SMEM_BYTES,
                         stream, 2, ta, tb, p);
}

template <int BN, int EC, int FM>
static int launch_gemm(const ergm_gemm_args* a, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap ta, tb;
  int rc;
  // A: K-major -> tensor [M rows][K] (dim0 = K); MN-major -> tensor [K rows][M] (dim0 = M)
  if (a->a_major == ERGM_MAJOR_K)
    rc = encode_tmap_2d(&ta, a->a, 2, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda * 2, BK, BM);
  else
    rc = encode_tmap_2d(&ta, a->a, 2, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda * 2, 64, BK);
  if (rc) return rc;
  if (a->b_major == ERGM_MAJOR_K)
    rc = encode_tmap_2d(&tb, a->b, 2, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb * 2, BK, BN);
  else
    rc = encode_tmap_2d(&tb, a->b, 2, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb * 2, 64, BK);
  if (rc) return rc;

  GemmParams p;
  p.d = a->d; p.bias = a->bias; p.residual = a->residual; p.preact = a->preact; p.colsum = a->colsum;
  p.ldd = a->ldd; p.ldr = a->ldr;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.a_mn = a->a_major == ERGM_MAJOR_MN; p.b_mn = a->b_major == ERGM_MAJOR_MN;
  p.d_f32 = a->d_dtype == ERGM_DT_F32;
  p.epi = a->epilogue;
  p.split_k = a->split_k < 1 ? 1 : a->split_k;
  p.m_tiles = (a->M + BM - 1) / BM;
  p.n_tiles = (a->N + BN - 1) / BN;
  p.kb_total = (a->K + BK - 1) / BK;
  if (p.split_k > p.kb_total) p.split_k = p.kb_total;
  p.kb_per_split = (p.kb_total + p.split_k - 1) / p.split_k;
  p.split_k = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.dropout_p = a->dropout_p;
  p.site = make_site(a->seed, a->offset, a->dropout_p, (uint32_t)a->N);
  p.dyn = a->dyn_count; p.dyn_dim = a->dyn_count ? a->dyn_dim : 0;
  p.pdl_trigger = a->M > 128;

  ERGM_SET_SMEM_ATTR((gemm_bf16_kernel<BN, EC, FM>), Cfg::SMEM_BYTES);
  const int total = p.m_tiles * p.n_tiles * p.split_k;
  const int grid = total < num_sms() ? total : num_sms();
  return (int)launch_pdl(gemm_bf16_kernel<BN, EC, FM>, dim3(grid), dim3(GEMM_THREADS), (size_t)Cfg::SMEM_BYTES, stream,
                         1, ta, tb, p);
}

}  // namespace ergm

extern "C" int ergm_gemm_bf16(const ergm_gemm_args* a, void* stream) {
  using namespace ergm;
  if (!a || !a->a || !a->b || !a->d) return ERGM_ERR_ARG;
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return ERGM_ERR_ARG;
  if (a->epilogue & ~0xff) return ERGM_ERR_ARG;  // only the ERGM_EPI_* bits of the header exist
  if ((a->epilogue & ERGM_EPI_BIAS) && !a->bias) return ERGM_ERR_ARG;
  if ((a->epilogue & ERGM_EPI_RESIDUAL) && !a->residual) return ERGM_ERR_ARG;
  if ((a->epilogue & (ERGM_EPI_PREACT | ERGM_EPI_GELU_GRAD)) && !a->preact) return ERGM_ERR_ARG;
  if ((a->epilogue & ERGM_EPI_PREACT) && (a->epilogue & ERGM_EPI_GELU_GRAD)) return ERGM_ERR_ARG;
  if ((a->epilogue & ERGM_EPI_GELU_GRAD) && a->d_dtype != ERGM_DT_BF16) return ERGM_ERR_ARG;
  if ((a->epilogue & ERGM_EPI_ATOMIC) && a->d_dtype != ERGM_DT_F32) return ERGM_ERR_ARG;
  if (a->split_k > 1 && !(a->epilogue & ERGM_EPI_ATOMIC)) return ERGM_ERR_ARG;
  if ((a->epilogue & ERGM_EPI_DROPOUT) && !(a->dropout_p >= 0.f && a->dropout_p < 1.f))
    return ERGM_ERR_ARG;
  // vector paths need 16-byte aligned rows
  const int dal = a->d_dtype == ERGM_DT_F32 ? 4 : 8;
  if (a->ldd % dal || (reinterpret_cast<uintptr_t>(a->d) & 15)) return ERGM_ERR_ARG;
  if ((a->epilogue & ERGM_EPI_RESIDUAL) &&
      (a->ldr % 4 || (reinterpret_cast<uintptr_t>(a->residual) & 15)))
    return ERGM_ERR_ARG;
  if ((a->epilogue & ERGM_EPI_BIAS) && (reinterpret_cast<uintptr_t>(a->bias) & 15))
    return ERGM_ERR_ARG;
  if (a->lda % 8 || a->ldb % 8) return ERGM_ERR_ARG;
  if (a->dyn_count && a->dyn_dim != 1 && a->dyn_dim != 2) return ERGM_ERR_ARG;
  if (a->dyn_count && a->dyn_dim == 2 && !(a->epilogue & ERGM_EPI_ATOMIC)) return ERGM_ERR_ARG;  // K = 0 must write nothing
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int bn = a->block_n;
  if (bn == 0 && a->dyn_count && a->dyn_dim == 1 && a->dyn_hint > 0 && a->dyn_hint < a->M && a->N >= 256) {
    // Run-time M (packed batch) with a known typical row count: choose the tile by a wave-quantisation cost model on
    // the rows that will really exist, cost = waves x (per-SM tile area) x (K + epilogue-equivalent) / speed.  The
    // static-M rules below were measured at M = 8192 and e.g. send K >= 1536 to 256x256 pair tiles whatever the tile
    // count - at 6156 rows that is 75 tiles on 74 clusters: two waves.
    const long Mh = a->dyn_hint;
    const int sk = a->split_k < 1 ? 1 : a->split_k;
    const int nc = num_sms() / 2, ns = num_sms();
    const double kw = (double)a->K + 384.0;            // per-element epilogue cost ~ a 384-deep mainloop (r1_gemm_epilogue.md)
    auto cost = [&](long tiles, int slots, double area_per_sm, double speed) {
      const long waves = (tiles + slots - 1) / slots;
      return (double)waves * area_per_sm * kw / speed;
    };
    double best = cost(((Mh + 255) / 256) * ((a->N + 255) / 256) * sk, nc, 128.0 * 256.0, a->K >= 1536 ? 1.12 : 1.0);
    bn = 2256;
    const long mt = (Mh + BM - 1) / BM;
    const double c256 = cost(mt * ((a->N + 255) / 256) * sk, ns, 128.0 * 256.0, 1.0);
    if (c256 < best) { best = c256; bn = 256; }
    if (a->N % 192 == 0) {
      const double c192 = cost(mt * (a->N / 192) * sk, ns, 128.0 * 192.0, 0.97);
      if (c192 < best) { best = c192; bn = 192; }
    }
    const double c128 = cost(mt * ((a->N + 127) / 128) * sk, ns, 128.0 * 128.0, 0.92);
    if (c128 < best) { best = c128; bn = 128; }
  }
  if (bn == 0 && a->M >= 512 && a->N >= 256) {
    // Measured on the model's shapes (profiles/r1_gemm_epilogue.md): the CTA-pair kernel with 256x256 tiles
    // wins whenever its tiles fill the 74 clusters reasonably or the mainloop is long (K >= 1536: the
    // wave-quantisation loss is outweighed by the halved B traffic per SM); short-K problems that would
    // leave a third of the clusters idle (N = 768, K = 768: 96 tiles) run on single-CTA 128x128 / 128x256
    // tiles; the 256x128 pair tiles never won.
    const int sk = a->split_k < 1 ? 1 : a->split_k;
    const long t2 = (long)((a->M + 255) / 256) * ((a->N + 255) / 256) * sk;
    const int nc = num_sms() / 2, ns = num_sms();
    auto eff = [](long tiles, int slots) { return (double)tiles / (double)(((tiles + slots - 1) / slots) * slots); };
    if (a->K >= 1536 || eff(t2, nc) >= 0.8) {
      bn = 2256;
    } else {
      const long mt = (a->M + BM - 1) / BM;
      const long t256 = mt * ((a->N + 255) / 256) * sk, t128 = mt * ((a->N + 127) / 128) * sk;
      bn = eff(t256, ns) >= eff(t128, ns) - 0.05 ? 256 : 128;
      // 128x192 tiles: N = 768 splits into 4 (256 tiles = 1.73 waves, as wave-efficient as 128x128 with a third
      // fewer accumulator hand-offs): 8192x768x768 dgrad 16.7 -> 14.8 us, fwd + residual + dropout 29.9 -> 26.1 us
      if (bn == 128 && a->N % 192 == 0 && eff(mt * (a->N / 192) * sk, ns) >= eff(t128, ns) - 0.02) bn = 192;
    }
  }
  if (bn == 0) {
    // auto: widest tile that still yields at least ~1 wave of CTAs
    const long mt = (a->M + BM - 1) / BM;
    const int sk = a->split_k < 1 ? 1 : a->split_k;
    bn = 256;
    while (bn > 64 && mt * ((a->N + bn - 1) / bn) * sk < num_sms()) bn >>= 1;
  }
  int ec = EC_LINEAR;
  if (a->epilogue & ERGM_EPI_EXACT) ec = EC_EXACT;
  else if (a->epilogue & ERGM_EPI_GELU) ec = EC_GELU;
  else if (a->epilogue & ERGM_EPI_GELU_GRAD) ec = EC_GELU_GRAD;
  else if (a->epilogue & ERGM_EPI_PREACT) ec = EC_GELU;
  // lean per-mode epilogue (interior tiles) for the hot launch configurations
  int fm = FM_NONE;
  {
    const int e = a->epilogue;
    const bool f32 = a->d_dtype == ERGM_DT_F32;
    const bool aligned = (reinterpret_cast<uintptr_t>(a->d) & 15) == 0 && a->ldd % (f32 ? 4 : 8) == 0;
    if (aligned && bn != 64 && !(bn == 192 && a->N % 192)) {
      if (!f32 && e == 0) fm = FM_BF16;
      else if (!f32 && e == ERGM_EPI_BIAS) fm = FM_BF16_BIAS;
      else if (!f32 && e == (ERGM_EPI_BIAS | ERGM_EPI_GELU | ERGM_EPI_PREACT) && a->preact) fm = FM_GELU;
      else if (f32 && e == 0) fm = FM_F32;
      else if (f32 && (e & ERGM_EPI_RESIDUAL) &&
               (e & ~(ERGM_EPI_BIAS | ERGM_EPI_RESIDUAL | ERGM_EPI_DROPOUT)) == 0) fm = FM_F32_RES;
      else if (f32 && e == ERGM_EPI_ATOMIC) fm = FM_ATOMIC;
      else if (!f32 && e == ERGM_EPI_GELU_GRAD && a->preact) fm = FM_GELU_GRAD;
    }
  }
  if (a->colsum) {
    // fused bias-gradient column sums exist only in the lean GELU' epilogue, which runs on interior tiles
    const int tm = bn > 2000 ? 256 : 128, tn = bn % 1000;
    if (fm != FM_GELU_GRAD || a->M % tm || a->N % tn) return ERGM_ERR_UNSUPPORTED;
  }
#define ERGM_DISPATCH_EC(FN, BNV)                                   \
  switch (fm) {                                                     \
    case FM_BF16: return FN<BNV, EC_LINEAR, FM_BF16>(a, s);         \
    case FM_BF16_BIAS: return FN<BNV, EC_LINEAR, FM_BF16_BIAS>(a, s); \
    case FM_F32: return FN<BNV, EC_LINEAR, FM_F32>(a, s);           \
    case FM_F32_RES: return FN<BNV, EC_LINEAR, FM_F32_RES>(a, s);   \
    case FM_GELU: return FN<BNV, EC_GELU, FM_GELU>(a, s);           \
    case FM_ATOMIC: return FN<BNV, EC_LINEAR, FM_ATOMIC>(a, s);     \
    case FM_GELU_GRAD: return FN<BNV, EC_GELU_GRAD, FM_GELU_GRAD>(a, s); \
    default: break;                                                 \
  }                                                                 \
  switch (ec) {                                                     \
    case EC_LINEAR: return FN<BNV, EC_LINEAR, FM_NONE>(a, s);       \
    case EC_GELU: return FN<BNV, EC_GELU, FM_NONE>(a, s);           \
    case EC_GELU_GRAD: return FN<BNV, EC_GELU_GRAD, FM_NONE>(a, s); \
    default: return FN<BNV, EC_EXACT, FM_NONE>(a, s);               \
  }
#define ERGM_DISPATCH_EC_SLOW(FN, BNV)                              \
  switch (ec) {                                                     \
    case EC_LINEAR: return FN<BNV, EC_LINEAR, FM_NONE>(a, s);       \
    case EC_GELU: return FN<BNV, EC_GELU, FM_NONE>(a, s);           \
    case EC_GELU_GRAD: return FN<BNV, EC_GELU_GRAD, FM_NONE>(a, s); \
    default: return FN<BNV, EC_EXACT, FM_NONE>(a, s);               \
  }
  switch (bn) {
    case 2256: ERGM_DISPATCH_EC(launch_gemm2, 256)
    case 2128: ERGM_DISPATCH_EC(launch_gemm2, 128)
    case 256: ERGM_DISPATCH_EC(launch_gemm, 256)
    case 192: ERGM_DISPATCH_EC(launch_gemm, 192)
    case 128: ERGM_DISPATCH_EC(launch_gemm, 128)
    case 64: ERGM_DISPATCH_EC_SLOW(launch_gemm, 64)
    default: return ERGM_ERR_ARG;
  }
#undef ERGM_DISPATCH_EC_SLOW
#undef ERGM_DISPATCH_EC
}

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
constexpr int SMEM_BUDGET = 200 * 1024;

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (SMEM_BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

struct GemmParams {
  void* d;
  const float* bias;
  const float* residual;
  void* preact;
  int64_t ldd, ldr;
  int M, N, K;
  int a_mn, b_mn;
  int d_f32;
  int epi;
  int split_k;
  int m_tiles, n_tiles, kb_total, kb_per_split;
  float dropout_p;
  DropoutSite site;
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a,
                 const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned tiles
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
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
    const bool has_bias = (p.epi & ERGM_EPI_BIAS) != 0;
    const bool do_gelu = (p.epi & ERGM_EPI_GELU) != 0;
    const bool do_res = (p.epi & ERGM_EPI_RESIDUAL) != 0;
    const bool do_atomic = (p.epi & ERGM_EPI_ATOMIC) != 0;
    const bool do_drop = (p.epi & ERGM_EPI_DROPOUT) != 0;
    const bool do_pre = (p.epi & ERGM_EPI_PREACT) != 0;
    const bool exact = (p.epi & ERGM_EPI_EXACT) != 0;
    const bool do_gelu_grad = (p.epi & ERGM_EPI_GELU_GRAD) != 0;
    const float keep_scale = do_drop ? 1.0f / (1.0f - p.dropout_p) : 1.0f;
    const DropoutSite site = p.site.resolved();
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = tile / tiles_mn;
      const int mn = tile - split * tiles_mn;
      const int m0 = (mn % p.m_tiles) * BM;
      const int n0 = (mn / p.m_tiles) * BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const bool first_split = (split == 0);
#pragma unroll 1
      for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 32) {
        const int col0 = n0 + c;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        const bool full = (col0 + 32 <= p.N);
        if (has_bias && first_split) {
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
              v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i < p.N) v[i] += __ldg(p.bias + col0 + i);
          }
        }
        if (row_ok) {
          if (do_pre) {
            __nv_bfloat16* pp = reinterpret_cast<__nv_bfloat16*>(p.preact) + (int64_t)row * p.ldd + col0;
            if (full) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint4 u = make_uint4(pack_bf16x2(v[i], v[i + 1]), pack_bf16x2(v[i + 2], v[i + 3]),
                                     pack_bf16x2(v[i + 4], v[i + 5]), pack_bf16x2(v[i + 6], v[i + 7]));
                *reinterpret_cast<uint4*>(pp + i) = u;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) pp[i] = __float2bfloat16_rn(v[i]);
            }
          }
          if (do_gelu_grad) {
            // v *= gelu_new'(u), u = saved pre-activation (bf16 [M, ldd]) -> dU of model.py:264
            const __nv_bfloat16* up =
                reinterpret_cast<const __nv_bfloat16*>(p.preact) + (int64_t)row * p.ldd + col0;
            if (full) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                const uint4 u4 = *reinterpret_cast<const uint4*>(up + i);
                const uint32_t uu[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = unpack_bf16x2(uu[j]);
                  v[i + 2 * j] *= exact ? gelu_new_grad<true>(f.x) : gelu_new_grad<false>(f.x);
                  v[i + 2 * j + 1] *= exact ? gelu_new_grad<true>(f.y) : gelu_new_grad<false>(f.y);
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) v[i] *= gelu_new_grad<true>(__bfloat162float(up[i]));
            }
          }
          if (do_gelu) {
            if (exact) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = gelu_new<true>(v[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = gelu_new<false>(v[i]);
            }
          }
          if (do_drop) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const uint32_t keep = site.keep4((uint32_t)row, (uint32_t)(col0 + i) >> 2);
              v[i] = (keep & 1u) ? v[i] * keep_scale : 0.f;
              v[i + 1] = (keep & 2u) ? v[i + 1] * keep_scale : 0.f;
              v[i + 2] = (keep & 4u) ? v[i + 2] * keep_scale : 0.f;
              v[i + 3] = (keep & 8u) ? v[i + 3] * keep_scale : 0.f;
            }
          }
          if (do_res) {
            const float* rp = p.residual + (int64_t)row * p.ldr + col0;
            if (full) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 r4 = *reinterpret_cast<const float4*>(rp + i);
                v[i] += r4.x; v[i + 1] += r4.y; v[i + 2] += r4.z; v[i + 3] += r4.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) v[i] += rp[i];
            }
          }
          if (p.d_f32) {
            float* dp = reinterpret_cast<float*>(p.d) + (int64_t)row * p.ldd + col0;
            if (do_atomic) {
              if (full) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dp + i),
                               "f"(v[i]), "f"(v[i + 1]), "f"(v[i + 2]), "f"(v[i + 3])
                               : "memory");
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (col0 + i < p.N) atomicAdd(dp + i, v[i]);
              }
            } else if (full) {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(dp + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) dp[i] = v[i];
            }
          } else {
            __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.d) + (int64_t)row * p.ldd + col0;
            if (full) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint4 u = make_uint4(pack_bf16x2(v[i], v[i + 1]), pack_bf16x2(v[i + 2], v[i + 3]),
                                     pack_bf16x2(v[i + 4], v[i + 5]), pack_bf16x2(v[i + 6], v[i + 7]));
                *reinterpret_cast<uint4*>(dp + i) = u;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) dp[i] = __float2bfloat16_rn(v[i]);
            }
          }
        }
      }
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

template <int BN>
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
  p.d = a->d; p.bias = a->bias; p.residual = a->residual; p.preact = a->preact;
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

  static bool attr_set = false;
  if (!attr_set) {
    ERGM_CUDA_TRY(cudaFuncSetAttribute(gemm_bf16_kernel<BN>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int total = p.m_tiles * p.n_tiles * p.split_k;
  const int grid = total < num_sms() ? total : num_sms();
  gemm_bf16_kernel<BN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
  return (int)cudaGetLastError();
}

}  // namespace ergm

extern "C" int ergm_gemm_bf16(const ergm_gemm_args* a, void* stream) {
  using namespace ergm;
  if (!a || !a->a || !a->b || !a->d) return ERGM_ERR_ARG;
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) return ERGM_ERR_ARG;
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
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int bn = a->block_n;
  if (bn == 0) {
    // auto: widest tile that still yields at least ~1 wave of CTAs
    const long mt = (a->M + BM - 1) / BM;
    const int sk = a->split_k < 1 ? 1 : a->split_k;
    bn = 256;
    while (bn > 64 && mt * ((a->N + bn - 1) / bn) * sk < num_sms()) bn >>= 1;
  }
  switch (bn) {
    case 256: return launch_gemm<256>(a, s);
    case 128: return launch_gemm<128>(a, s);
    case 64: return launch_gemm<64>(a, s);
    default: return ERGM_ERR_ARG;
  }
}

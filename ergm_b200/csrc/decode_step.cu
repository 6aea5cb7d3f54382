// ergm_decode_stack - all transformer blocks of one decode step (one new token for each of <= 64 sequences)
// as ONE persistent kernel of head-clusters.  Replaces, for generation, L x GPT2Block.forward
// (/root/reference/src/model.py:286-341) driven token by token (main.py:253-282, model.py:228-236).
//
// Why (profiles/r1_decode.md, r2_decode.md): a decode step of GPT-2 small moves 545 MB (84 us of HBM time) but is
// a chain of 5 dependent stages per block; as separate launches (or as a persistent kernel with one grid barrier
// per stage) every stage costs 4-9 us of pure dependency latency: 492 us per step = 17 % of the HBM roofline.
// Here the dependencies are cut to TWO grid-wide synchronisations per block:
//
//   phase 1  (cluster c = head c, 8 CTAs):  LN1 -> q/k/v columns of head c -> [cluster barrier] -> paged attention
//            of head c (CTA r owns sequences 8r .. 8r+7, appends the new K/V) -> [cluster barrier] -> out-proj
//            partial  x_next += ctx_c @ W_o[64c:64c+64, :]   (split-K over heads: 12 reductions per element)
//   phase 2  (cluster c = 256 columns of the MLP's inner dimension):  LN2 -> fc + gelu_new -> [cluster barriers]
//            -> MLP-proj partial  x_next += g[:, 256c:256c+256] @ W_p[256c:256c+256, :]
//
// Everything a stage needs from its predecessor INSIDE a phase travels through distributed shared memory behind
// a hardware cluster barrier (~0.3 us) instead of L2 + a grid barrier (~2 us).  The residual stream rotates
// through three fp32 buffers (read / accumulate / being zeroed), so a fast cluster's reductions can never race
// with a slow cluster's LayerNorm reads.  Weights are re-packed once per weight version into one contiguous
// blob per (layer, phase, CTA) in mma.m16n8k16 B-fragment order (LayerNorm gamma folded in, beta folded into
// the bias): ONE cp.async.bulk per CTA per phase, issued one to two phases ahead of its use.
#include "../../include/ergm_b200.h"
#include "common.cuh"
#include <cstdio>
#include <cstdlib>

namespace ergm {

constexpr int DS_THREADS = 512;
constexpr int DS_CS = 8;      // CTAs per cluster
constexpr int DS_M = 64;      // sequences (rows) per step
constexpr int DS_PAGE = 16;   // tokens per KV page (decode.cu)
constexpr int DS_UNR = 8;     // tokens in flight per 8-lane group
constexpr int DS_ACC_LD = 36; // fp32 words per row of the k-split accumulation tile

struct DsLayer {   // device resident, one per block
  const __nv_bfloat16* w1;   // phase-1 blobs  [nh * 8][p1_bytes]   (q/k/v part, then out-proj part)
  const __nv_bfloat16* wfc;  // fc blobs       [nh * 8][fc_bytes]
  const __nv_bfloat16* wp2;  // MLP-proj blobs [nh * 8][pj_bytes]
  const float* b_qkv;        // [3H]  (ln_1 beta folded in)
  const float* b_o;          // [H]
  const float* b_fc;         // [I]   (ln_2 beta folded in)
  const float* b_p2;         // [H]
  __nv_bfloat16* pool;       // paged K/V of this layer [pages][2][nh][16][64]
};

struct DsParams {
  const DsLayer* layers;
  int L, H, I, nh, B;
  float* x0; float* x1; float* x2;   // residual stream ring (x0: embeddings in; result in ring[(2L) % 3])
  const int* block_table;            // [B, max_pages]
  const int* seq_lens;               // [B] cached tokens (the new token goes to this slot)
  int max_pages;
  float eps, scale;
  unsigned int* sync_ctr;            // zeroed by the host before every launch
  long long* trace;                  // nullable profiling aid: %globaltimer stamps of CTA 0 (16 per block, see DS_STAMP)
  int a_stride;                      // bytes per row of the normalised A operand
  int off_a, off_w0, off_w1, off_qkv, off_ctx, off_acc, off_merge, off_tab, off_bias;
  int p1_bytes, qkv_part_bytes, fc_bytes, pj_bytes;
};

ERGM_DEVINL void ds_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
ERGM_DEVINL void ds_ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
ERGM_DEVINL void ds_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
ERGM_DEVINL uint2 ds_lds_v2(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
ERGM_DEVINL void ds_red_shared(uint32_t addr, float v) {
  asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
ERGM_DEVINL void ds_st_cluster_b32(uint32_t caddr, uint32_t v) {
  asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(caddr), "r"(v) : "memory");
}
ERGM_DEVINL void ds_st_cluster_v2(uint32_t caddr, uint32_t a, uint32_t b) {
  asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(caddr), "r"(a), "r"(b) : "memory");
}
ERGM_DEVINL void ds_st_cluster_v4(uint32_t caddr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(caddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
ERGM_DEVINL void ds_red_global_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
ERGM_DEVINL unsigned int ds_ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// grid-wide barrier: every CTA arrives once per call; `epoch` counts the calls (identical in all threads).
// The release is cumulative over the CTA's earlier writes / reductions (ordered before it by the bar.sync).
ERGM_DEVINL void ds_grid_sync(unsigned int* ctr, unsigned int& epoch) {
  __syncthreads();
  ++epoch;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    const unsigned int target = epoch * gridDim.x;
    while (ds_ld_acquire(ctr) < target) {}
  }
  __syncthreads();
}

ERGM_DEVINL float ds_dot8(const uint4 a, const uint4 b) {
  const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
  const float2 b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y), b2 = unpack_bf16x2(b.z), b3 = unpack_bf16x2(b.w);
  return a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
}

// LayerNorm of the fp32 residual rows (gamma / beta live in the packed weights / folded biases) -> bf16 A operand
// in shared memory; zeroes the k-split accumulation tile on the way.  16 warps x 4 rows.
template <int NV>
ERGM_DEVINL void ds_layernorm(const DsParams& p, unsigned char* smem, const float* x) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* a_sm = smem + p.off_a;
  float* acc_sm = reinterpret_cast<float*>(smem + p.off_acc);
  for (int i = threadIdx.x; i < DS_M * DS_ACC_LD; i += DS_THREADS) acc_sm[i] = 0.f;
  constexpr int H = NV * 128;
  const float invH = 1.f / (float)H;
#pragma unroll
  for (int rr = 0; rr < 4; rr += 2) {
    const int row0 = warp * 4 + rr;
    float4 v[2][NV];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int i = 0; i < NV; ++i)
        v[q][i] = (row0 + q) < p.B ? __ldcg(reinterpret_cast<const float4*>(x + (int64_t)(row0 + q) * H) + lane + 32 * i)
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
        if (row < p.B)
          o = make_uint2(pack_bf16x2((v[q][i].x - mean) * rstd, (v[q][i].y - mean) * rstd),
                         pack_bf16x2((v[q][i].z - mean) * rstd, (v[q][i].w - mean) * rstd));
        *reinterpret_cast<uint2*>(a_sm + (size_t)row * p.a_stride + (lane + 32 * i) * 8) = o;
      }
    }
  }
}

// [64 x (8 * NT)] += A[64 x K] @ W (K split over the four warp quarters, reduced with shared-memory atomics).
// A: row-major bf16 in smem (stride a_stride bytes); W: fragment-packed [n8][kb][lane][4 bf16] at `w`.
template <int NT>
ERGM_DEVINL void ds_mma_ksplit(uint32_t a_sm, int a_stride, uint32_t w, int KB, uint32_t acc_sm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = warp & 3, kq = warp >> 2;
  const int kper = KB >> 2;
  float acc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[n][i] = 0.f;
  const uint32_t a_base = a_sm + (uint32_t)(mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * a_stride + (uint32_t)(lane >> 4) * 16;
#pragma unroll 4
  for (int kk = 0; kk < kper; ++kk) {
    const int k = kq * kper + kk;
    uint32_t a0, a1, a2, a3;
    ds_ldmatrix_x4(a_base + k * 32, a0, a1, a2, a3);
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const uint2 b = ds_lds_v2(w + (uint32_t)((n * KB + k) * 32 + lane) * 8);
      ds_mma(acc[n], a0, a1, a2, a3, b.x, b.y);
    }
  }
  const int row = mt * 16 + (lane >> 2), col = (lane & 3) * 2;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
    const uint32_t base = acc_sm + (uint32_t)(row * DS_ACC_LD + n * 8 + col) * 4;
    ds_red_shared(base, acc[n][0]);
    ds_red_shared(base + 4, acc[n][1]);
    ds_red_shared(base + 8 * DS_ACC_LD * 4, acc[n][2]);
    ds_red_shared(base + 8 * DS_ACC_LD * 4 + 4, acc[n][3]);
  }
}

// x_next[64, cols of this CTA] += A[64 x (16 * KB)] @ W (+ bias + x_cur for the designated cluster): the residual
// projections.  W fragment-packed [n8][kb][lane][4]; this CTA owns columns [col0, col0 + 8 * NTILES).
ERGM_DEVINL void ds_mma_residual(const DsParams& p, uint32_t a_sm, int a_stride, uint32_t w, int KB, int ntiles, int col0,
                                 const float* bias_sm, const float* x_cur, float* x_next, bool lead) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = warp & 3, ng = warp >> 2;
  float acc[3][4];
  // The designated cluster adds the bias and carries the residual stream over: its accumulators START from
  // x_cur + bias (loads issued before the k loop, so their L2 latency hides behind the MMAs)
#pragma unroll
  for (int n = 0; n < 3; ++n) {
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[n][i] = 0.f;
    const int n8 = ng + 4 * n;
    if (lead && n8 < ntiles) {
      const int cl = n8 * 8 + (lane & 3) * 2;
      const float b0 = bias_sm[cl], b1 = bias_sm[cl + 1];
#pragma unroll
      for (int hrow = 0; hrow < 2; ++hrow) {
        const int rr = mt * 16 + (lane >> 2) + 8 * hrow;
        if (rr < p.B) {
          const float2 xc = __ldcg(reinterpret_cast<const float2*>(x_cur + (int64_t)rr * p.H + col0 + cl));
          acc[n][2 * hrow] = xc.x + b0;
          acc[n][2 * hrow + 1] = xc.y + b1;
        }
      }
    }
  }
  const uint32_t a_base = a_sm + (uint32_t)(mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * a_stride + (uint32_t)(lane >> 4) * 16;
#pragma unroll 4
  for (int k = 0; k < KB; ++k) {
    uint32_t a0, a1, a2, a3;
    ds_ldmatrix_x4(a_base + k * 32, a0, a1, a2, a3);
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      const int n8 = ng + 4 * n;
      if (n8 < ntiles) {
        const uint2 b = ds_lds_v2(w + (uint32_t)((n8 * KB + k) * 32 + lane) * 8);
        ds_mma(acc[n], a0, a1, a2, a3, b.x, b.y);
      }
    }
  }
  const int row = mt * 16 + (lane >> 2);
#pragma unroll
  for (int n = 0; n < 3; ++n) {
    const int n8 = ng + 4 * n;
    if (n8 >= ntiles) continue;
    const int col = col0 + n8 * 8 + (lane & 3) * 2;
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int r = row + 8 * hrow;
      if (r >= p.B) continue;
      ds_red_global_v2(x_next + (int64_t)r * p.H + col, acc[n][2 * hrow], acc[n][2 * hrow + 1]);
    }
  }
}

// profiling aid (ergm_decode_stack's `trace` argument): thread 0 of CTA 0 records where the block's time goes
#define DS_STAMP(slot)                                                                   \
  do {                                                                                   \
    if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) {                                \
      long long t_;                                                                      \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                             \
      p.trace[l * 16 + (slot)] = t_;                                                     \
    }                                                                                    \
  } while (0)

template <int NV>
__global__ void __launch_bounds__(DS_THREADS, 1) decode_stack_kernel(const DsParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = (int)cluster_ctarank();          // rank in the cluster
  const int c = (int)blockIdx.x / DS_CS;         // cluster = head (phase 1) / inner-dimension range (phase 2)
  const int H = p.H, nh = p.nh;
  const int KB = H / 16;                         // k16 steps of an H-long reduction
  const int IC = p.I / nh;                       // inner-dimension columns per cluster (256 for I = 4H)
  const int HC = H / DS_CS;                      // output columns per CTA of the residual projections
  const uint32_t sm0 = smem_u32(smem);
  const uint32_t wbar[2] = {sm0, sm0 + 8};
  const uint32_t wreg[2] = {sm0 + (uint32_t)p.off_w0, sm0 + (uint32_t)p.off_w1};
  const uint32_t a_sm = sm0 + (uint32_t)p.off_a, acc_sm = sm0 + (uint32_t)p.off_acc;
  const uint32_t qkv_sm = sm0 + (uint32_t)p.off_qkv, ctx_sm = sm0 + (uint32_t)p.off_ctx;
  float* merge_sm = reinterpret_cast<float*>(smem + p.off_merge);
  float* bias_sm = reinterpret_cast<float*>(smem + p.off_bias);   // [0, 32): q|k|v or fc bias of my columns; [32, 32 + HC): residual bias
  int* tab_sm = reinterpret_cast<int*>(smem + p.off_tab);   // [8][max_pages] block-table rows, then [8] lengths
  uint32_t wpar = 0;                                         // parities of the two weight-region barriers
  unsigned int epoch = 0;
  float* ring[3] = {p.x0, p.x1, p.x2};
  const int blob = c * DS_CS + r;

  if (tid == 0) {
    mbar_init(wbar[0], 1);
    mbar_init(wbar[1], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < DS_CS * p.max_pages; i += DS_THREADS) {
    const int b = r * DS_CS + i / p.max_pages;
    tab_sm[i] = b < p.B ? p.block_table[b * p.max_pages + i % p.max_pages] : 0;
  }
  if (tid < DS_CS) tab_sm[DS_CS * p.max_pages + tid] = (r * DS_CS + tid) < p.B ? p.seq_lens[r * DS_CS + tid] : 0;
  __syncthreads();
  auto load_w = [&](int region, const __nv_bfloat16* src, int bytes) {
    if (tid == 0) {
      mbar_expect_tx(wbar[region], (uint32_t)bytes);
      ds_bulk_g2s(wreg[region], src, (uint32_t)bytes, wbar[region]);
    }
  };
  auto wait_w = [&](int region) {
    mbar_wait(wbar[region], (wpar >> region) & 1u);
    wpar ^= 1u << region;
  };
  // weights of block 0: phase 1 -> region 0, fc -> region 1
  load_w(0, p.layers[0].w1 + (size_t)blob * (p.p1_bytes / 2), p.p1_bytes);
  load_w(1, p.layers[0].wfc + (size_t)blob * (p.fc_bytes / 2), p.fc_bytes);
  cluster_sync();  // every CTA of the cluster is running: its shared memory may be written remotely from now on

  for (int l = 0; l < p.L; ++l) {
    const DsLayer ly = p.layers[l];
    const int r1 = l & 1, r2 = r1 ^ 1;           // weight regions: phase 1 / MLP-proj in r1, fc in r2
    // =========================== phase 1: attention block ===========================
    {
      const int ph = 2 * l;
      const float* x_cur = ring[ph % 3];
      float* x_next = ring[(ph + 1) % 3];
      float* x_zero = ring[(ph + 2) % 3];
      DS_STAMP(0);
      if (l > 0) ds_grid_sync(p.sync_ctr, epoch);   // x_cur complete (block l-1's MLP reductions have landed)
      DS_STAMP(1);
      for (int i = (int)blockIdx.x * DS_THREADS + tid; i < p.B * H / 4; i += (int)gridDim.x * DS_THREADS)
        reinterpret_cast<float4*>(x_zero)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      // this phase's biases -> smem now, so that no epilogue waits for an L2 round trip
      if (tid < 24) bias_sm[tid] = __ldg(ly.b_qkv + (tid >> 3) * H + c * 64 + r * 8 + (tid & 7));
      else if (tid >= 32 && tid < 32 + HC) bias_sm[tid] = __ldg(ly.b_o + HC * r + tid - 32);
      ds_layernorm<NV>(p, smem, x_cur);
      DS_STAMP(2);
      wait_w(r1);
      __syncthreads();
      ds_mma_ksplit<3>(a_sm, p.a_stride, wreg[r1], KB, acc_sm);   // q | k | v columns 8r .. 8r+7 of head c
      __syncthreads();
      DS_STAMP(3);
      // bias, bf16, and hand every sequence's 24 values to the CTA that owns the sequence (rank = row / 8)
      for (int i = tid; i < DS_M * 12; i += DS_THREADS) {
        const int row = i / 12, rem = i - row * 12, which = rem >> 2, pr = rem & 3;
        const float2 a = *reinterpret_cast<const float2*>(smem + p.off_acc + (size_t)(row * DS_ACC_LD + which * 8 + 2 * pr) * 4);
        const uint32_t v = pack_bf16x2(a.x + bias_sm[which * 8 + 2 * pr], a.y + bias_sm[which * 8 + 2 * pr + 1]);
        const uint32_t dst = qkv_sm + (uint32_t)(((row & 7) * 192 + which * 64 + r * 8 + 2 * pr) * 2);
        ds_st_cluster_b32(mapa_cluster(dst, (uint32_t)(row >> 3)), v);
      }
      cluster_arrive();   // my share of q / k / v has been handed out ...
      // ---- paged one-query attention: 2 warps per sequence, 8 eight-lane groups with private softmax states ----
      {
        const int ls = warp >> 1, half = warp & 1;
        const int b = r * DS_CS + ls;
        const int g8 = half * 4 + (lane >> 3), gl = lane & 7;
        const int n_old = b < p.B ? tab_sm[DS_CS * p.max_pages + ls] : 0;
        const int* bt = tab_sm + ls * p.max_pages;
        const unsigned char* qrow = smem + p.off_qkv + (size_t)ls * 384;
        auto kv_row = [&](int tok, int which) -> const uint4* {
          const int page = bt[tok / DS_PAGE];
          return reinterpret_cast<const uint4*>(ly.pool + ((((int64_t)page * 2 + which) * nh + c) * DS_PAGE + tok % DS_PAGE) * 64) + gl;
        };
        uint4 kk[DS_UNR], vv[DS_UNR];
        auto load_pass = [&](int tb) {
#pragma unroll
          for (int u = 0; u < DS_UNR; ++u) {
            const int tok = tb + u * 8 + g8;
            const bool ok = tok < n_old;
            kk[u] = ok ? __ldcg(kv_row(tok, 0)) : make_uint4(0u, 0u, 0u, 0u);
            vv[u] = ok ? __ldcg(kv_row(tok, 1)) : make_uint4(0u, 0u, 0u, 0u);
          }
        };
        load_pass(0);     // the cached K / V do not depend on this step's q: in flight across the cluster barrier
        cluster_wait();   // ... and everybody else's share for my 8 sequences has arrived
        DS_STAMP(4);
        const uint4 qv = *reinterpret_cast<const uint4*>(qrow + gl * 16);
        float m_run = -INFINITY, l_run = 0.f;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int tb = 0; tb < n_old; tb += 8 * DS_UNR) {
          if (tb > 0) load_pass(tb);
          float sc[DS_UNR];
          float m_new = m_run;
#pragma unroll
          for (int u = 0; u < DS_UNR; ++u) {
            float s = ds_dot8(qv, kk[u]);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            sc[u] = (tb + u * 8 + g8 < n_old) ? s * p.scale : -INFINITY;
            m_new = fmaxf(m_new, sc[u]);
          }
          if (m_new > -INFINITY) {
            const float corr = __expf(m_run - m_new);
            l_run *= corr;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] *= corr;
#pragma unroll
            for (int u = 0; u < DS_UNR; ++u) {
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
        DS_STAMP(5);
        if (l + 1 < p.L && b < p.B) {
          // next block's K / V pages of this (sequence, head): pull them into L2 now (2 KB per page and K / V),
          // so that the next block's attention reads hit L2 instead of paying the DRAM latency in the chain
          const __nv_bfloat16* npool = p.layers[l + 1].pool;
          const int npg = (n_old + DS_PAGE - 1) / DS_PAGE;
          for (int i = half * 32 + lane; i < 2 * npg; i += 64) {
            const int page = bt[i >> 1], which = i & 1;
            const __nv_bfloat16* src = npool + (((int64_t)page * 2 + which) * nh + c) * DS_PAGE * 64;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(DS_PAGE * 64 * 2) : "memory");
          }
        }
        if (half == 0) {
          // the new token (from the projection just computed): scored by all four groups of this warp (uniform
          // shuffles), folded into group 0's state; lanes 0-7 / 8-15 append its K / V row to the page pool
          const uint4 kn = *reinterpret_cast<const uint4*>(qrow + 128 + gl * 16);
          const uint4 vn = *reinterpret_cast<const uint4*>(qrow + 256 + gl * 16);
          float s = ds_dot8(qv, kn);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s *= p.scale;
          if (lane < 8) {
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
          if (lane < 16 && b < p.B) {
            const int pos = n_old, which = lane >> 3;
            const int page = bt[pos / DS_PAGE];
            __nv_bfloat16* dst = ly.pool + ((((int64_t)page * 2 + which) * nh + c) * DS_PAGE + pos % DS_PAGE) * 64 + gl * 8;
            *reinterpret_cast<uint4*>(dst) = which == 0 ? kn : vn;
          }
        }
        // merge the four groups of the warp (lanes gl, gl+8, gl+16, gl+24 hold the same 8 output dims)
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
          const float m_o = __shfl_xor_sync(0xffffffffu, m_run, off);
          const float l_o = __shfl_xor_sync(0xffffffffu, l_run, off);
          const float m_new = fmaxf(m_run, m_o);
          const float wa = m_run > -INFINITY ? __expf(m_run - m_new) : 0.f;
          const float wb = m_o > -INFINITY ? __expf(m_o - m_new) : 0.f;
          l_run = l_run * wa + l_o * wb;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float a_o = __shfl_xor_sync(0xffffffffu, acc[i], off);
            acc[i] = acc[i] * wa + a_o * wb;
          }
          m_run = m_new;
        }
        float* ms = merge_sm + (ls * 2 + half) * 68;
        if (lane < 8) {
#pragma unroll
          for (int i = 0; i < 8; ++i) ms[gl * 8 + i] = acc[i];
          if (lane == 0) { ms[64] = m_run; ms[65] = l_run; }
        }
        __syncthreads();
        if (half == 0 && lane < 8) {
          const float* m0 = merge_sm + (ls * 2) * 68;
          const float* m1 = m0 + 68;
          const float M = fmaxf(m0[64], m1[64]);
          const float w0 = m0[64] > -INFINITY ? __expf(m0[64] - M) : 0.f;
          const float w1 = m1[64] > -INFINITY ? __expf(m1[64] - M) : 0.f;
          const float Ls = m0[65] * w0 + m1[65] * w1;
          const float inv = Ls > 0.f ? 1.f / Ls : 0.f;
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = (m0[gl * 8 + i] * w0 + m1[gl * 8 + i] * w1) * inv;
          const uint4 pk = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                                      pack_bf16x2(o[6], o[7]));
          // the attention output of sequence b, head c -> row b of every CTA's ctx tile (A operand of the out-proj)
          const uint32_t dst = ctx_sm + (uint32_t)(b * 144 + gl * 16);
#pragma unroll
          for (int peer = 0; peer < DS_CS; ++peer) ds_st_cluster_v4(mapa_cluster(dst, (uint32_t)peer), pk);
        }
      }
      cluster_sync();   // ctx tile [64 x 64] of head c complete everywhere
      DS_STAMP(6);
      // ---- out-proj partial of head c: x_next[:, HC*r .. ] += ctx_c @ W_o[64c:64c+64, HC*r ..] ----
      ds_mma_residual(p, ctx_sm, 144, wreg[r1] + (uint32_t)p.qkv_part_bytes, 4, HC / 8, HC * r, bias_sm + 32, x_cur, x_next, c == 0);
      __syncthreads();  // region r1 is free: the MLP-proj weights of this block move in
      DS_STAMP(7);
      load_w(r1, ly.wp2 + (size_t)blob * (p.pj_bytes / 2), p.pj_bytes);
    }
    // =========================== phase 2: MLP ===========================
    {
      const int ph = 2 * l + 1;
      const float* x_cur = ring[ph % 3];
      float* x_next = ring[(ph + 1) % 3];
      float* x_zero = ring[(ph + 2) % 3];
      ds_grid_sync(p.sync_ctr, epoch);              // x_cur complete (all heads' out-proj reductions have landed)
      DS_STAMP(8);
      for (int i = (int)blockIdx.x * DS_THREADS + tid; i < p.B * H / 4; i += (int)gridDim.x * DS_THREADS)
        reinterpret_cast<float4*>(x_zero)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tid < 32) bias_sm[tid] = __ldg(ly.b_fc + c * IC + r * 32 + tid);
      else if (tid < 32 + HC) bias_sm[tid] = __ldg(ly.b_p2 + HC * r + tid - 32);
      ds_layernorm<NV>(p, smem, x_cur);
      DS_STAMP(9);
      wait_w(r2);
      __syncthreads();
      ds_mma_ksplit<4>(a_sm, p.a_stride, wreg[r2], KB, acc_sm);   // fc columns IC*c + 32r .. +32
      __syncthreads();  // region r2 is free: the phase-1 weights of the next block move in
      DS_STAMP(10);
      if (l + 1 < p.L) load_w(r2, p.layers[l + 1].w1 + (size_t)blob * (p.p1_bytes / 2), p.p1_bytes);
      cluster_arrive();   // I am done reading my A operand ...
      {
        // bias + gelu_new, bf16, broadcast my 32 columns of g to all 8 CTAs (g tile [64 x IC] aliases the A region)
        const int row = tid >> 3, c4 = (tid & 7) * 4;
        const float4 a = *reinterpret_cast<const float4*>(smem + p.off_acc + (size_t)(row * DS_ACC_LD + c4) * 4);
        const float4 bb = *reinterpret_cast<const float4*>(bias_sm + c4);
        const uint32_t lo = pack_bf16x2(gelu_new<false>(a.x + bb.x), gelu_new<false>(a.y + bb.y));
        const uint32_t hi = pack_bf16x2(gelu_new<false>(a.z + bb.z), gelu_new<false>(a.w + bb.w));
        const uint32_t dst = a_sm + (uint32_t)(row * (2 * IC + 16) + (r * 32 + c4) * 2);
        cluster_wait();   // ... and so is every other CTA of the cluster: the g tile may overwrite the A operands
#pragma unroll
        for (int peer = 0; peer < DS_CS; ++peer) ds_st_cluster_v2(mapa_cluster(dst, (uint32_t)peer), lo, hi);
      }
      cluster_arrive();
      wait_w(r1);         // MLP-proj weights (in flight since the end of phase 1)
      cluster_wait();     // g tile complete everywhere
      DS_STAMP(11);
      ds_mma_residual(p, a_sm, 2 * IC + 16, wreg[r1], IC / 16, HC / 8, HC * r, bias_sm + 32, x_cur, x_next, c == 0);
      __syncthreads();  // region r1 is free: the fc weights of the next block move in
      DS_STAMP(12);
      if (l + 1 < p.L) load_w(r1, p.layers[l + 1].wfc + (size_t)blob * (p.fc_bytes / 2), p.fc_bytes);
    }
  }
  cluster_sync();  // nobody may exit while a peer can still write into its shared memory
}

// ------------------------------------------------------------------------------------------
// weight packing for decode_stack_kernel: W (fp32, logical [K, N], row stride ld) -> per-(cluster, rank) blobs of
// mma.m16n8k16 B fragments [n8][kb][lane][4 bf16]:
//   lane l, element j:  k = k0 + kb*16 + (l%4)*2 + (j&1) + ((j&2) ? 8 : 0),   n = n0 + n8*8 + l/4
// kind 0: q|k|v part of the phase-1 blob   (n8 = q/k/v, columns which*H + 64c + 8r ..,  K = H, gamma = ln_1)
// kind 1: out-proj part of the phase-1 blob (rows 64c .. 64c+64, columns (H/8) r ..)
// kind 2: fc blob                           (columns (I/nh) c + 32 r .., K = H, gamma = ln_2)
// kind 3: MLP-proj blob                     (rows (I/nh) c .., columns (H/8) r ..)
// ------------------------------------------------------------------------------------------
__global__ void ds_pack_kernel(const float* __restrict__ w, int64_t ld, const float* __restrict__ gamma, int kind, int H,
                               int I, int nh, __nv_bfloat16* __restrict__ dst, int blob_elems, int part_off_elems) {
  const int KB = kind == 1 ? 4 : (kind == 3 ? (I / nh) / 16 : H / 16);
  const int NT = kind == 0 ? 3 : (kind == 2 ? 4 : H / 64);
  const int64_t per_blob = (int64_t)NT * KB * 32;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per_blob * nh * DS_CS) return;
  const int blob = (int)(idx / per_blob);
  const int64_t rem = idx - (int64_t)blob * per_blob;
  const int lane = (int)(rem & 31);
  const int kb = (int)((rem >> 5) % KB), n8 = (int)((rem >> 5) / KB);
  const int c = blob / DS_CS, r = blob % DS_CS;
  int k0, n0;
  if (kind == 0) { k0 = 0; n0 = n8 * H + c * 64 + r * 8; }
  else if (kind == 1) { k0 = c * 64; n0 = r * (H / 8) + n8 * 8; }
  else if (kind == 2) { k0 = 0; n0 = c * (I / nh) + r * 32 + n8 * 8; }
  else { k0 = c * (I / nh); n0 = r * (H / 8) + n8 * 8; }
  __align__(8) __nv_bfloat16 v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = k0 + kb * 16 + (lane & 3) * 2 + (j & 1) + ((j & 2) ? 8 : 0);
    const int n = n0 + (lane >> 2);
    float x = w[(int64_t)k * ld + n];
    if (gamma) x *= gamma[k];
    v[j] = __float2bfloat16_rn(x);
  }
  *reinterpret_cast<uint2*>(dst + (int64_t)blob * blob_elems + part_off_elems + rem * 4) = *reinterpret_cast<const uint2*>(v);
}

struct DsGeom {
  int qkv_part_bytes, p1_bytes, fc_bytes, pj_bytes;
};
static DsGeom ds_geom(int H, int I, int nh) {
  DsGeom g;
  g.qkv_part_bytes = 3 * (H / 16) * 256;
  g.p1_bytes = g.qkv_part_bytes + (H / 64) * 4 * 256;
  g.fc_bytes = 4 * (H / 16) * 256;
  g.pj_bytes = (H / 64) * ((I / nh) / 16) * 256;
  return g;
}

}  // namespace ergm

using namespace ergm;

static int ds_supported(int H, int I, int nh, int B) {
  if (B <= 0 || B > DS_M || nh <= 0 || H != nh * 64) return 0;
  if (H != 128 && H != 256 && H != 512 && H != 768) return 0;  // H = 1024: the operand tiles exceed 227 KB of smem
  if (I != 4 * H) return 0;
  return 1;
}

extern "C" int ergm_decode_stack_blob_bytes(int H, int I, int nh, int64_t* p1, int64_t* fc, int64_t* pj) {
  if (!ds_supported(H, I, nh, 1) || !p1 || !fc || !pj) return ERGM_ERR_UNSUPPORTED;
  const DsGeom g = ds_geom(H, I, nh);
  *p1 = (int64_t)g.p1_bytes * nh * DS_CS;
  *fc = (int64_t)g.fc_bytes * nh * DS_CS;
  *pj = (int64_t)g.pj_bytes * nh * DS_CS;
  return ERGM_OK;
}

extern "C" int ergm_decode_stack_pack(const float* w_qkv, const float* gamma1, const float* w_o, const float* w_fc,
                                      const float* gamma2, const float* w_p2, int H, int I, int nh, void* p1_blobs,
                                      void* fc_blobs, void* pj_blobs, void* stream) {
  if (!w_qkv || !gamma1 || !w_o || !w_fc || !gamma2 || !w_p2 || !p1_blobs || !fc_blobs || !pj_blobs) return ERGM_ERR_ARG;
  if (!ds_supported(H, I, nh, 1)) return ERGM_ERR_UNSUPPORTED;
  const DsGeom g = ds_geom(H, I, nh);
  cudaStream_t st = (cudaStream_t)stream;
  auto launch = [&](const float* w, int64_t ld, const float* gamma, int kind, void* dst, int blob_bytes, int part_off) {
    const int KB = kind == 1 ? 4 : (kind == 3 ? (I / nh) / 16 : H / 16);
    const int NT = kind == 0 ? 3 : (kind == 2 ? 4 : H / 64);
    const int64_t n = (int64_t)NT * KB * 32 * nh * DS_CS;
    ds_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, ld, gamma, kind, H, I, nh,
                                                                reinterpret_cast<__nv_bfloat16*>(dst), blob_bytes / 2, part_off / 2);
  };
  launch(w_qkv, 3 * H, gamma1, 0, p1_blobs, g.p1_bytes, 0);
  launch(w_o, H, nullptr, 1, p1_blobs, g.p1_bytes, g.qkv_part_bytes);
  launch(w_fc, I, gamma2, 2, fc_blobs, g.fc_bytes, 0);
  launch(w_p2, H, nullptr, 3, pj_blobs, g.pj_bytes, 0);
  return (int)cudaGetLastError();
}

extern "C" int ergm_decode_stack(const void* layer_table, int L, int H, int I, int nh, int B, float* x0, float* x1,
                                 float* x2, const int* block_table, const int* seq_lens, int max_pages, float eps,
                                 unsigned int* sync_ctr, int64_t* trace, void* stream) {
  if (!layer_table || !x0 || !x1 || !x2 || !block_table || !seq_lens || !sync_ctr || L <= 0 || max_pages <= 0)
    return ERGM_ERR_ARG;
  if (!ds_supported(H, I, nh, B)) return ERGM_ERR_UNSUPPORTED;
  const DsGeom g = ds_geom(H, I, nh);
  DsParams p{};
  p.layers = reinterpret_cast<const DsLayer*>(layer_table);
  p.L = L; p.H = H; p.I = I; p.nh = nh; p.B = B;
  p.x0 = x0; p.x1 = x1; p.x2 = x2;
  p.block_table = block_table; p.seq_lens = seq_lens; p.max_pages = max_pages;
  p.eps = eps; p.scale = 0.125f;
  p.sync_ctr = sync_ctr;
  p.trace = reinterpret_cast<long long*>(trace);
  p.qkv_part_bytes = g.qkv_part_bytes; p.p1_bytes = g.p1_bytes; p.fc_bytes = g.fc_bytes; p.pj_bytes = g.pj_bytes;
  p.a_stride = 2 * H + 16;
  int a_bytes = DS_M * p.a_stride;
  const int g_bytes = DS_M * (2 * (I / nh) + 16);
  if (a_bytes < g_bytes) a_bytes = g_bytes;
  int wmax = g.p1_bytes > g.fc_bytes ? g.p1_bytes : g.fc_bytes;
  if (wmax < g.pj_bytes) wmax = g.pj_bytes;
  auto al = [](int v) { return (v + 127) & ~127; };
  p.off_a = 128;
  p.off_w0 = al(p.off_a + a_bytes);
  p.off_w1 = al(p.off_w0 + wmax);
  p.off_qkv = al(p.off_w1 + wmax);
  p.off_ctx = al(p.off_qkv + DS_CS * 192 * 2);
  p.off_acc = al(p.off_ctx + DS_M * 144);
  p.off_merge = al(p.off_acc + DS_M * DS_ACC_LD * 4);
  p.off_tab = al(p.off_merge + DS_CS * 2 * 68 * 4);
  p.off_bias = al(p.off_tab + (DS_CS * max_pages + DS_CS) * 4);
  const int smem = al(p.off_bias + (32 + H / DS_CS) * 4);
  if (smem > 232448) return ERGM_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  ERGM_CUDA_TRY(cudaMemsetAsync(sync_ctr, 0, sizeof(unsigned int), st));
  ERGM_CUDA_TRY(cudaMemsetAsync(x1, 0, (size_t)B * H * sizeof(float), st));  // phase 0 accumulates into ring[1]
  auto launch = [&](auto kern) -> int {
    ERGM_SET_SMEM_ATTR(kern, 232448);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(nh * DS_CS));
    cfg.blockDim = dim3(DS_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = DS_CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;   // all nh clusters must be co-resident: the grid barrier spins
    at[1].val.cooperative = 1;
    cfg.attrs = at;
    // the grid barrier needs every cluster resident at once: refuse geometries the device cannot hold
    cfg.numAttrs = 1;
    int max_clusters = 0;
    const cudaError_t oe = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
    static const bool dbg = getenv("ERGM_DS_DEBUG") != nullptr;   // read once per process
    if (dbg) fprintf(stderr, "[ergm_decode_stack] smem %d B, occupancy query rc %d, max active clusters %d, need %d\n",
                     smem, (int)oe, max_clusters, nh);
    if (oe != cudaSuccess) { cudaGetLastError(); return ERGM_ERR_UNSUPPORTED; }
    if (max_clusters < nh) return ERGM_ERR_UNSUPPORTED;
    static std::atomic<int> coop_ok{-1};   // does this driver accept cooperative + cluster launches?
    if (coop_ok.load() != 0) {
      cfg.numAttrs = 2;
      const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
      if (dbg) fprintf(stderr, "[ergm_decode_stack] cooperative cluster launch rc %d\n", (int)e);
      if (e == cudaSuccess) { coop_ok.store(1); return 0; }
      if (coop_ok.load() == 1) return (int)e;
      cudaGetLastError();
      coop_ok.store(0);   // fall through: plain cluster launch (occupancy checked above; one CTA per SM)
    }
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kern, p);
  };
  switch (H / 128) {
    case 1: return launch(decode_stack_kernel<1>);
    case 2: return launch(decode_stack_kernel<2>);
    case 4: return launch(decode_stack_kernel<4>);
    case 6: return launch(decode_stack_kernel<6>);
  }
  return ERGM_ERR_UNSUPPORTED;
}

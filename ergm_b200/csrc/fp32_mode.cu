// "fp32 mode" support kernels (BASELINE north_star: logits within 1e-4 / greedy ids bit-exact).
//
// The tensor cores only take 16-bit operands, so an fp32-accurate product is computed as a
// split-operand bf16 GEMM on the same tcgen05 kernel: every fp32 value x is decomposed exactly
// into three bf16 pieces x = h + m + l (24 significant bits), and
//     a*b ~= ah*bh + ah*bm + am*bh + ah*bl + al*bh + am*bm          (error ~2^-24 |a b|)
// which is ONE bf16 GEMM over a 6x longer reduction dimension:
//     A' = [ah | ah | am | ah | al | am]   (ergm_split3_expand, side 0, along K)
//     B' = [bh | bm | bh | bl | bh | bm]   (side 1)
// with fp32 accumulation in TMEM.  Attention in this mode is a plain fp32 CUDA-core kernel (one
// warp per query row, online softmax) — it is a verification path, not the throughput path.
#include "../../include/ergm_b200.h"
#include "common.cuh"

namespace ergm {

__device__ __forceinline__ void split3(float x, __nv_bfloat16& h, __nv_bfloat16& m, __nv_bfloat16& l) {
  h = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(h);
  m = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(m);
  l = __float2bfloat16_rn(r2);
}

// src fp32 [rows, cols] (ld_src).  kdim == 1: the reduction dimension is `cols` (K-major operand):
// dst is [rows, 6*cols] with the six segments side by side.  kdim == 0: the reduction dimension is
// `rows` (MN-major operand, e.g. Conv1D weight [K, N]): dst is [6*rows, cols], segments stacked.
__global__ void __launch_bounds__(256)
split3_expand_kernel(const float* __restrict__ src, int64_t ld_src, __nv_bfloat16* __restrict__ dst,
                     int64_t ld_dst, int rows, int cols, int side, int kdim) {
  // segment -> which piece (0=h,1=m,2=l)
  const int pieceA[6] = {0, 0, 1, 0, 2, 1};
  const int pieceB[6] = {0, 1, 0, 2, 0, 1};
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
    __nv_bfloat16 p[3];
    split3(src[(int64_t)r * ld_src + c], p[0], p[1], p[2]);
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const __nv_bfloat16 v = p[side == 0 ? pieceA[s] : pieceB[s]];
      if (kdim == 1) dst[(int64_t)r * ld_dst + (int64_t)s * cols + c] = v;
      else dst[((int64_t)s * rows + r) * ld_dst + c] = v;
    }
  }
}

// fp32 attention: one warp per (b, h, query); q/k/v fp32 [B*T, ld]; out fp32 merged heads.
__global__ void __launch_bounds__(256)
attn_fwd_f32_kernel(const float* __restrict__ q, int64_t ld_q, int q_col0, const float* __restrict__ k,
                    int64_t ld_k, int k_col0, const float* __restrict__ v, int64_t ld_v, int v_col0,
                    float* __restrict__ out, int64_t ld_out, const int* __restrict__ kv_lens, int B, int nh,
                    int Tq, int Tk, int causal, int causal_off) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * 8 + warp;
  if (item >= (int64_t)B * nh * Tq) return;
  const int qi = (int)(item % Tq);
  const int h = (int)((item / Tq) % nh);
  const int b = (int)(item / ((int64_t)Tq * nh));
  int kv_len = Tk;
  if (kv_lens) kv_len = min(kv_len, kv_lens[b]);
  const int last = causal ? min(kv_len - 1, qi + causal_off) : kv_len - 1;
  const float* qp = q + ((int64_t)b * Tq + qi) * ld_q + q_col0 + h * 64;
  const float q0 = qp[lane], q1 = qp[lane + 32];
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int j = 0; j <= last; ++j) {
    const float* kp = k + ((int64_t)b * Tk + j) * ld_k + k_col0 + h * 64;
    float s = q0 * kp[lane] + q1 * kp[lane + 32];
    s = warp_sum(s) / 8.0f;  // model.py:122-125: / sqrt(head_dim)
    const float mn = fmaxf(m, s);
    const float alpha = expf(m - mn), p = expf(s - mn);
    const float* vp = v + ((int64_t)b * Tk + j) * ld_v + v_col0 + h * 64;
    o0 = o0 * alpha + p * vp[lane];
    o1 = o1 * alpha + p * vp[lane + 32];
    l = l * alpha + p;
    m = mn;
  }
  float* op = out + ((int64_t)b * Tq + qi) * ld_out + h * 64;
  const float inv = l > 0.f ? 1.f / l : 0.f;
  op[lane] = o0 * inv;
  op[lane + 32] = o1 * inv;
}

}  // namespace ergm

using namespace ergm;

extern "C" int ergm_split3_expand(const float* src, int64_t ld_src, void* dst_bf16, int64_t ld_dst, int rows,
                                  int cols, int side, int kdim, void* stream) {
  if (!src || !dst_bf16 || rows <= 0 || cols <= 0 || (side != 0 && side != 1) || (kdim != 0 && kdim != 1))
    return ERGM_ERR_ARG;
  const int64_t n = (int64_t)rows * cols;
  int64_t blocks = (n + 255) / 256;
  if (blocks > num_sms() * 32) blocks = num_sms() * 32;
  split3_expand_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      src, ld_src, reinterpret_cast<__nv_bfloat16*>(dst_bf16), ld_dst, rows, cols, side, kdim);
  return (int)cudaGetLastError();
}

extern "C" int ergm_attn_fwd_f32(const float* q, int64_t ld_q, int q_col0, const float* k, int64_t ld_k,
                                 int k_col0, const float* v, int64_t ld_v, int v_col0, float* out,
                                 int64_t ld_out, const int* kv_lens, int B, int nh, int Tq, int Tk,
                                 int head_dim, int causal, int causal_off, void* stream) {
  if (!q || !k || !v || !out || B <= 0 || nh <= 0 || Tq <= 0 || Tk <= 0) return ERGM_ERR_ARG;
  if (head_dim != 64) return ERGM_ERR_UNSUPPORTED;
  const int64_t items = (int64_t)B * nh * Tq;
  attn_fwd_f32_kernel<<<(unsigned)((items + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      q, ld_q, q_col0, k, ld_k, k_col0, v, ld_v, v_col0, out, ld_out, kv_lens, B, nh, Tq, Tk, causal, causal_off);
  return (int)cudaGetLastError();
}

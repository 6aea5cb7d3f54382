// HBM-bound row kernels of the ERGM path: embedding + multimodal fusion (model.py:458-507),
// LayerNorm forward / backward fused with the residual-gradient add (model.py:298,318,332,578),
// column sums for bias gradients, dtype casts.  One warp owns one row; every global access is a
// 16-byte vector, rows are contiguous so each warp request covers whole 512-byte spans.
#include "../../include/ergm_b200.h"
#include "common.cuh"
#include "dropout.cuh"

namespace ergm {

constexpr int ROW_WARPS = 8;  // warps (rows in flight) per CTA

// ------------------------------------------------------------------------------------------
// embed_fuse_fwd:  h[b,t] = ((wte[id] (+img_b if t==0) (+aud_b if t==1)) + wpe[pos]) + wte[tt]
// ------------------------------------------------------------------------------------------
struct EmbedFwdParams {
  const int64_t* ids;
  const int64_t* tts;       // nullable
  const int64_t* pos_ids;   // nullable; pos = pos_ids[b * pos_stride_b + t] (stride 0: shared [T])
  int64_t pos_stride_b;
  const int* past_lens;     // nullable: per-sequence past length (decode with ragged prompts)
  const float* wte;
  const float* wpe;
  const float* imgs;        // nullable, [B, ld_img]
  const float* auds;        // nullable, [B, ld_aud]
  float* out;               // [B*T, H]
  int64_t ld_img, ld_aud;
  int B, T, H, past_len, vocab, n_pos;
  DropoutSite drop;
  int do_drop;
  int* err_flag;            // set to 1 on an out-of-range id (device-side assert replacement)
  PackView pk;              // packed batch: output row r <- (sample row_b[r], position row_t[r])
};

__global__ void __launch_bounds__(ROW_WARPS * 32) embed_fuse_fwd_kernel(const EmbedFwdParams p_in) {
  EmbedFwdParams p = p_in;
  p.drop = p_in.drop.resolved();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = p.B * p.T;
  const float keep_scale = p.do_drop ? p.drop.keep_scale() : 1.f;
  const int n_rows = p.pk.on() ? min(rows, *p.pk.n_rows) : rows;
  for (int row = blockIdx.x * ROW_WARPS + warp; row < n_rows; row += gridDim.x * ROW_WARPS) {
    int b = row / p.T, t = row - b * p.T;
    if (p.pk.on()) { b = p.pk.row_b[row]; t = p.pk.row_t[row]; }
    const int src = b * p.T + t;   // index into the padded [B, T] id / type arrays
    const int64_t id = p.ids[src];
    const int64_t tt = p.tts ? p.tts[src] : -1;
    const int64_t pos = p.pos_ids ? p.pos_ids[b * p.pos_stride_b + t]
                                  : (int64_t)((p.past_lens ? p.past_lens[b] : p.past_len) + t);
    if (id < 0 || id >= p.vocab || pos < 0 || pos >= p.n_pos || (p.tts && (tt < 0 || tt >= p.vocab))) {
      // the reference raises IndexError / a device assert here; we flag (checked by the host after the step)
      // and leave a defined (zero) row behind instead of stale workspace contents
      if (lane == 0) *p.err_flag = 1;
      float4* oz = reinterpret_cast<float4*>(p.out + (int64_t)row * p.H);
      for (int c = lane; c < p.H / 4; c += 32) oz[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const float4* we = reinterpret_cast<const float4*>(p.wte + id * p.H);
    const float4* wp = reinterpret_cast<const float4*>(p.wpe + pos * p.H);
    const float4* wt = p.tts ? reinterpret_cast<const float4*>(p.wte + tt * p.H) : nullptr;
    const float4* fz = nullptr;
    if (p.imgs && t == 0) fz = reinterpret_cast<const float4*>(p.imgs + (int64_t)b * p.ld_img);
    if (p.auds && t == 1) fz = reinterpret_cast<const float4*>(p.auds + (int64_t)b * p.ld_aud);
    float4* o = reinterpret_cast<float4*>(p.out + (int64_t)row * p.H);
    for (int c = lane; c < p.H / 4; c += 32) {
      float4 e = __ldg(we + c);
      if (fz) { const float4 f = __ldg(fz + c); e.x += f.x; e.y += f.y; e.z += f.z; e.w += f.w; }
      const float4 q = __ldg(wp + c);
      e.x += q.x; e.y += q.y; e.z += q.z; e.w += q.w;
      if (wt) { const float4 s = __ldg(wt + c); e.x += s.x; e.y += s.y; e.z += s.z; e.w += s.w; }
      if (p.do_drop) {
        const uint32_t k = p.drop.keep4((uint32_t)row, (uint32_t)c);
        e.x = (k & 1u) ? e.x * keep_scale : 0.f; e.y = (k & 2u) ? e.y * keep_scale : 0.f;
        e.z = (k & 4u) ? e.z * keep_scale : 0.f; e.w = (k & 8u) ? e.w * keep_scale : 0.f;
      }
      o[c] = e;
    }
  }
}

// gather rows of an fp32 table into a bf16 matrix (caption embeddings, model.py:460-463)
__global__ void __launch_bounds__(ROW_WARPS * 32)
gather_rows_bf16_kernel(const int64_t* ids, const float* table, __nv_bfloat16* out, int rows, int H,
                        int vocab, int* err_flag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = blockIdx.x * ROW_WARPS + warp; row < rows; row += gridDim.x * ROW_WARPS) {
    const int64_t id = ids[row];
    uint2* dst = reinterpret_cast<uint2*>(out + (int64_t)row * H);
    if (id < 0 || id >= vocab) {
      if (lane == 0) *err_flag = 1;
      for (int c = lane; c < H / 4; c += 32) dst[c] = make_uint2(0u, 0u);
      continue;
    }
    const float4* src = reinterpret_cast<const float4*>(table + id * H);
    for (int c = lane; c < H / 4; c += 32) {
      const float4 e = __ldg(src + c);
      dst[c] = make_uint2(pack_bf16x2(e.x, e.y), pack_bf16x2(e.z, e.w));
    }
  }
}

// ------------------------------------------------------------------------------------------
// embed_bwd: scatter-add of dh rows into dwte (by input id and by token-type id) and dwpe.
// A CTA walks `rows_per_cta` consecutive rows; each thread owns 4 columns and keeps a running
// sum per index stream, flushing with one red.add.v4 whenever the index changes (token-type
// ids come in long runs, so their heavy contention collapses to one atomic per run).
// ------------------------------------------------------------------------------------------
struct EmbedBwdParams {
  const float* dh;          // [rows, H]
  const int64_t* ids;       // nullable
  const int64_t* tts;       // nullable
  const int64_t* pos_ids;   // nullable; positions used only when dwpe != null
  int64_t pos_stride_b;
  float* dwte;
  float* dwpe;              // nullable
  float* dimgs;             // nullable [B, H]  (+= dh[b,0])
  float* dauds;             // nullable [B, H]  (+= dh[b,1])
  int rows, T, H, past_len, rows_per_cta;
  DropoutSite drop;
  int do_drop;
  int vocab, n_pos;         // table heights: an index outside [0, height) is skipped and flagged
  int* err_flag;            // nullable
  PackView pk;              // packed batch: dh row r belongs to (sample row_b[r], position row_t[r])
};

ERGM_DEVINL void red_add4(float* addr, const float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

__global__ void embed_bwd_kernel(const EmbedBwdParams p_in) {
  EmbedBwdParams p = p_in;
  p.drop = p_in.drop.resolved();
  const int c = threadIdx.x;  // float4 column
  if (c >= p.H / 4) return;
  const int r0 = blockIdx.x * p.rows_per_cta;
  const int r1 = min(r0 + p.rows_per_cta, p.pk.on() ? min(p.rows, *p.pk.n_rows) : p.rows);
  const float keep_scale = p.do_drop ? p.drop.keep_scale() : 1.f;
  int64_t cur[3] = {-1, -1, -1};
  float4 acc[3];
#pragma unroll
  for (int s = 0; s < 3; ++s) acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
  float* tables[3] = {p.ids ? p.dwte : nullptr, p.tts ? p.dwte : nullptr, p.dwpe};
  for (int row = r0; row < r1; ++row) {
    float4 g = __ldg(reinterpret_cast<const float4*>(p.dh + (int64_t)row * p.H) + c);
    if (p.do_drop) {
      const uint32_t k = p.drop.keep4((uint32_t)row, (uint32_t)c);
      g.x = (k & 1u) ? g.x * keep_scale : 0.f; g.y = (k & 2u) ? g.y * keep_scale : 0.f;
      g.z = (k & 4u) ? g.z * keep_scale : 0.f; g.w = (k & 8u) ? g.w * keep_scale : 0.f;
    }
    int bq = row / p.T, t = row % p.T;
    if (p.pk.on()) { bq = p.pk.row_b[row]; t = p.pk.row_t[row]; }
    const int src = bq * p.T + t;
    int64_t idx[3];
    idx[0] = p.ids ? p.ids[src] : -1;
    idx[1] = p.tts ? p.tts[src] : -1;
    idx[2] = p.dwpe ? (p.pos_ids ? p.pos_ids[bq * p.pos_stride_b + t] : (int64_t)(p.past_len + t)) : -1;
    // never scatter outside the gradient tables (the forward flagged the same row): skip + flag
    const bool bad0 = p.ids && (idx[0] < 0 || idx[0] >= p.vocab), bad1 = p.tts && (idx[1] < 0 || idx[1] >= p.vocab),
               bad2 = p.dwpe && (idx[2] < 0 || idx[2] >= p.n_pos);
    if (bad0 || bad1 || bad2) {
      if (c == 0 && p.err_flag) *p.err_flag = 1;
      if (bad0) idx[0] = -1;
      if (bad1) idx[1] = -1;
      if (bad2) idx[2] = -1;
    }
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      if (!tables[s]) continue;
      if (idx[s] != cur[s]) {
        if (cur[s] >= 0) red_add4(tables[s] + cur[s] * p.H + 4 * c, acc[s]);
        cur[s] = idx[s];
        acc[s] = g;
      } else {
        acc[s].x += g.x; acc[s].y += g.y; acc[s].z += g.z; acc[s].w += g.w;
      }
    }
    if (p.dimgs && t == 0) red_add4(p.dimgs + (int64_t)bq * p.H + 4 * c, g);
    if (p.dauds && t == 1) red_add4(p.dauds + (int64_t)bq * p.H + 4 * c, g);
  }
#pragma unroll
  for (int s = 0; s < 3; ++s)
    if (tables[s] && cur[s] >= 0) red_add4(tables[s] + cur[s] * p.H + 4 * c, acc[s]);
}

// ------------------------------------------------------------------------------------------
// LayerNorm forward: y = (x - mean) * rstd * gamma + beta, biased variance, eps inside sqrt
// ------------------------------------------------------------------------------------------
template <int NV>  // NV = H / 128 float4 per lane
__global__ void __launch_bounds__(ROW_WARPS * 32)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ beta, __nv_bfloat16* __restrict__ y_bf16,
              float* __restrict__ y_f32, float* __restrict__ mean_out, float* __restrict__ rstd_out,
              int rows, float eps, const int* __restrict__ row_idx, const int* __restrict__ rows_dyn) {
  constexpr int H = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();  // the next kernel's blocks may be scheduled as ours drain ...
  pdl_wait();               // ... and we touch nothing before everything upstream has completed
  if (rows_dyn) rows = min(rows, *rows_dyn);   // packed batch: run-time row count
  float4 g[NV], bt[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    g[i] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
    bt[i] = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
  }
  for (int row = blockIdx.x * ROW_WARPS + warp; row < rows; row += gridDim.x * ROW_WARPS) {
    const int src_row = row_idx ? row_idx[row] : row;  // optional gather (last prompt token per sequence)
    const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)src_row * H);
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i] = xr[lane + 32 * i];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.f / H);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      ss += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(ss) * (1.f / H) + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o;
      o.x = (v[i].x - mean) * rstd * g[i].x + bt[i].x;
      o.y = (v[i].y - mean) * rstd * g[i].y + bt[i].y;
      o.z = (v[i].z - mean) * rstd * g[i].z + bt[i].z;
      o.w = (v[i].w - mean) * rstd * g[i].w + bt[i].w;
      if (y_bf16)
        reinterpret_cast<uint2*>(y_bf16 + (int64_t)row * H)[lane + 32 * i] =
            make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      if (y_f32) reinterpret_cast<float4*>(y_f32 + (int64_t)row * H)[lane + 32 * i] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm backward fused with the residual-stream gradient:
//   dx_out = dres_in + LN'(dy)           (fp32)
//   dx_bf16 = bf16(dropmask(dx_out))     (operand of the next dgrad / wgrad GEMMs)
//   dgamma += sum_rows dy * xhat, dbeta += sum_rows dy, dbias_next += sum_rows dx_bf16
// ------------------------------------------------------------------------------------------
struct LnBwdParams {
  const void* dy;        // bf16 or fp32 [rows, H]
  const float* x;
  const float* mean;
  const float* rstd;
  const float* gamma;
  const float* dres_in;  // nullable
  float* dx_out;         // nullable (may alias dres_in)
  __nv_bfloat16* dx_bf16;  // nullable
  float* dgamma;
  float* dbeta;
  float* dbias_next;     // nullable, colsum of dx_bf16 (as rounded)
  int rows;
  int dy_f32;
  DropoutSite drop;      // mask applied to the bf16 copy only
  int do_drop;
  const int* rows_dyn;   // nullable: run-time row count (packed batch); rows beyond it are not touched
};

// HBM-bound (dy + x + dres read, dx fp32 + dx bf16 written: 16 B / element at fp32 dy).  v1 (one warp
// per row, demand loads) was register-bound to 8 warps / SM: 2.3 TB/s; v2 (thread-per-column, a block
// barrier per 4 rows) 2.1 TB/s: both leave the memory pipe idle while they reduce.  v3: a producer warp
// streams tiles of 8 consecutive rows (three contiguous cp.async.bulk copies per tile: rows are
// contiguous) through a 2-3 stage shared-memory ring, so loads are in flight no matter what the 8
// consumer warps are doing; a consumer warp owns one row of the tile (warp-shuffle statistics, no block
// barrier in the loop) and keeps its share of the three column sums in registers.
constexpr int LNB_ROWS = 8;            // rows per tile = consumer warps
constexpr int LNB_THREADS = LNB_ROWS * 32;  // 256: registers are allocated per 4 warps, a 9th (producer) warp capped us at 168

ERGM_DEVINL void lnb_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int NV>
__global__ void __launch_bounds__(LNB_THREADS, 1) ln_bwd_kernel(const LnBwdParams p_in, int stages) {
  constexpr int H = NV * 128;
  extern __shared__ __align__(128) unsigned char lnb_smem[];
  LnBwdParams p = p_in;
  pdl_wait();               // programmatic dependent launch: nothing is read before the previous kernel completed
  pdl_launch_dependents();
  p.drop = p_in.drop.resolved();
  if (p.rows_dyn) p.rows = min(p.rows, *p.rows_dyn);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t dy_row = (uint32_t)H * (p.dy_f32 ? 4u : 2u), f_row = (uint32_t)H * 4u;
  const uint32_t dy_bytes = LNB_ROWS * dy_row, f_bytes = LNB_ROWS * f_row;
  const uint32_t stage_bytes = dy_bytes + f_bytes + (p.dres_in ? f_bytes : 0u);
  const uint32_t bars = smem_u32(lnb_smem);            // full[s] at 8s, empty[s] at 64 + 8s
  float4* s_gamma = reinterpret_cast<float4*>(lnb_smem + 128);  // gamma lives in smem, not in 24 registers
  unsigned char* ring = lnb_smem + 128 + H * 4;
  const int n_tiles = (p.rows + LNB_ROWS - 1) / LNB_ROWS;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 64 + 8 * s, LNB_ROWS); }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < H / 4; i += blockDim.x) s_gamma[i] = __ldg(reinterpret_cast<const float4*>(p.gamma) + i);
  __syncthreads();
  float4 ag[NV], ab[NV], an[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) ag[i] = ab[i] = an[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  // lane 0 of warp 0 doubles as the producer: it keeps `stages` tiles in flight, refilling the slot of tile
  // it-1 at the top of iteration it (by then the other warps have normally released it)
  auto issue_tile = [&](int tile, int s) {
    const int row0 = tile * LNB_ROWS;
    const uint32_t nr = (uint32_t)min(LNB_ROWS, p.rows - row0);
    const uint32_t dst = smem_u32(ring + (size_t)s * stage_bytes);
    mbar_expect_tx(bars + 8 * s, nr * (dy_row + f_row + (p.dres_in ? f_row : 0u)));
    lnb_bulk_g2s(dst, reinterpret_cast<const unsigned char*>(p.dy) + (size_t)row0 * dy_row, nr * dy_row, bars + 8 * s);
    lnb_bulk_g2s(dst + dy_bytes, p.x + (size_t)row0 * H, nr * f_row, bars + 8 * s);
    if (p.dres_in) lnb_bulk_g2s(dst + dy_bytes + f_bytes, p.dres_in + (size_t)row0 * H, nr * f_row, bars + 8 * s);
  };
  const bool producer = threadIdx.x == 0;
  if (producer) {
    int tile = blockIdx.x;
    for (int s = 0; s < stages && tile < n_tiles; ++s, tile += gridDim.x) issue_tile(tile, s);
  }
  {
    // ===================== consumers: warp w owns row w of every tile =====================
    const float keep_scale = p.do_drop ? p.drop.keep_scale() : 1.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int s = it % stages;
      if (producer && it >= 1) {
        const int prev = it - 1;
        const int nxt_tile = tile + (stages - 1) * (int)gridDim.x;  // tile index of iteration prev + stages
        if (nxt_tile < n_tiles) {
          mbar_wait(bars + 64 + 8 * (prev % stages), (uint32_t)((prev / stages) & 1));
          issue_tile(nxt_tile, prev % stages);
        }
      }
      const int row = tile * LNB_ROWS + warp;
      const bool row_ok = row < p.rows;
      const float mean = row_ok ? __ldg(p.mean + row) : 0.f, rstd = row_ok ? __ldg(p.rstd + row) : 0.f;
      mbar_wait(bars + 8 * s, (uint32_t)((it / stages) & 1));
      const unsigned char* st = ring + (size_t)s * stage_bytes;
      float4 dy[NV], xh[NV];
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if (p.dy_f32) {
            dy[i] = reinterpret_cast<const float4*>(st + (size_t)warp * dy_row)[lane + 32 * i];
          } else {
            const uint2 u = reinterpret_cast<const uint2*>(st + (size_t)warp * dy_row)[lane + 32 * i];
            const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
            dy[i] = make_float4(a.x, a.y, b.x, b.y);
          }
          xh[i] = reinterpret_cast<const float4*>(st + dy_bytes + (size_t)warp * f_row)[lane + 32 * i];
        }
      } else {  // row past the end (last tile): nothing to do, hand the slot back
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 64 + 8 * s);
        continue;
      }
      float c1 = 0.f, c2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        xh[i].x = (xh[i].x - mean) * rstd; xh[i].y = (xh[i].y - mean) * rstd;
        xh[i].z = (xh[i].z - mean) * rstd; xh[i].w = (xh[i].w - mean) * rstd;
        const float4 gm = s_gamma[lane + 32 * i];
        const float gx = dy[i].x * gm.x, gy = dy[i].y * gm.y, gz = dy[i].z * gm.z, gw = dy[i].w * gm.w;
        c1 += (gx + gy) + (gz + gw);
        c2 += (gx * xh[i].x + gy * xh[i].y) + (gz * xh[i].z + gw * xh[i].w);
      }
      c1 = warp_sum(c1) * (1.f / H);
      c2 = warp_sum(c2) * (1.f / H);
      const int64_t off = (int64_t)row * H;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c = lane + 32 * i;
        const float4 gm = s_gamma[c];
        // residual-stream gradient: read from the stage right before use (keeps 24 registers free)
        const float4 rs = p.dres_in
                              ? reinterpret_cast<const float4*>(st + dy_bytes + f_bytes + (size_t)warp * f_row)[c]
                              : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 dx;
        dx.x = rstd * (dy[i].x * gm.x - c1 - xh[i].x * c2) + rs.x;
        dx.y = rstd * (dy[i].y * gm.y - c1 - xh[i].y * c2) + rs.y;
        dx.z = rstd * (dy[i].z * gm.z - c1 - xh[i].z * c2) + rs.z;
        dx.w = rstd * (dy[i].w * gm.w - c1 - xh[i].w * c2) + rs.w;
        ag[i].x += dy[i].x * xh[i].x; ag[i].y += dy[i].y * xh[i].y; ag[i].z += dy[i].z * xh[i].z; ag[i].w += dy[i].w * xh[i].w;
        ab[i].x += dy[i].x; ab[i].y += dy[i].y; ab[i].z += dy[i].z; ab[i].w += dy[i].w;
        if (p.dx_out) reinterpret_cast<float4*>(p.dx_out + off)[c] = dx;
        if (p.dx_bf16) {
          if (p.do_drop) {
            const uint32_t k = p.drop.keep4((uint32_t)row, (uint32_t)c);
            dx.x = (k & 1u) ? dx.x * keep_scale : 0.f; dx.y = (k & 2u) ? dx.y * keep_scale : 0.f;
            dx.z = (k & 4u) ? dx.z * keep_scale : 0.f; dx.w = (k & 8u) ? dx.w * keep_scale : 0.f;
          }
          const uint2 u = make_uint2(pack_bf16x2(dx.x, dx.y), pack_bf16x2(dx.z, dx.w));
          reinterpret_cast<uint2*>(p.dx_bf16 + off)[c] = u;
          if (p.dbias_next) {
            const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
            an[i].x += a.x; an[i].y += a.y; an[i].z += b.x; an[i].w += b.y;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 64 + 8 * s);  // every read of this stage is done: slot may be refilled
    }
  }
  // ---- column sums: 8 warps -> smem (the ring is drained) -> one atomic per column per CTA ----
  __syncthreads();
  float* red = reinterpret_cast<float*>(ring);  // [3][LNB_ROWS][H]
  {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      reinterpret_cast<float4*>(red + ((size_t)0 * LNB_ROWS + warp) * H)[lane + 32 * i] = ag[i];
      reinterpret_cast<float4*>(red + ((size_t)1 * LNB_ROWS + warp) * H)[lane + 32 * i] = ab[i];
      reinterpret_cast<float4*>(red + ((size_t)2 * LNB_ROWS + warp) * H)[lane + 32 * i] = an[i];
    }
  }
  __syncthreads();
  const bool has_n = p.dbias_next && p.dx_bf16;
  for (int idx = threadIdx.x; idx < 3 * H; idx += blockDim.x) {
    const int q = idx / H, c = idx - q * H;
    if (q == 2 && !has_n) break;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < LNB_ROWS; ++w) sum += red[((size_t)q * LNB_ROWS + w) * H + c];
    atomicAdd((q == 0 ? p.dgamma : q == 1 ? p.dbeta : p.dbias_next) + c, sum);
  }
}

// ------------------------------------------------------------------------------------------
// colsum: out[N] += sum_rows src[rows, N]  (bf16 source; bias gradients)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ src, int64_t ld, int rows, int N,
                   float* __restrict__ out, int rows_per_cta) {
  __shared__ float red[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(r0 + rows_per_cta, rows);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col0 < N) {
    for (int r = r0 + warp; r < r1; r += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(src + (int64_t)r * ld + col0);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
      acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = threadIdx.x;
  if (blockIdx.x * 256 + c < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][c];
    atomicAdd(out + blockIdx.x * 256 + c, s);
  }
}

// GELU backward fused with the c_fc bias gradient (model.py:263-264 backward):
//   dU = dG * gelu_new'(U)  (in place over dG, bf16), colsum[n] += sum_rows dU[., n]
// Kept out of the dgrad GEMM epilogue on purpose: there the U tile would have to be fetched by
// demand loads that queue behind the mainloop's saturated TMA stream (measured +80 us per GEMM).
__global__ void __launch_bounds__(256)
gelu_bwd_colsum_kernel(__nv_bfloat16* __restrict__ dg, const __nv_bfloat16* __restrict__ u, int64_t ld,
                       int rows, int N, float* __restrict__ colsum, int rows_per_cta, int exact,
                       const int* __restrict__ rows_dyn) {
  __shared__ float red[8][256];
  if (rows_dyn) rows = min(rows, *rows_dyn);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(r0 + rows_per_cta, rows);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col0 < N) {
    for (int r = r0 + warp; r < r1; r += 8) {
      uint4 g4 = *reinterpret_cast<const uint4*>(dg + (int64_t)r * ld + col0);
      const uint4 u4 = *reinterpret_cast<const uint4*>(u + (int64_t)r * ld + col0);
      uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w};
      const uint32_t uw[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 g = unpack_bf16x2(gw[j]), uu = unpack_bf16x2(uw[j]);
        const float d0 = g.x * (exact ? gelu_new_grad<true>(uu.x) : gelu_new_grad<false>(uu.x));
        const float d1 = g.y * (exact ? gelu_new_grad<true>(uu.y) : gelu_new_grad<false>(uu.y));
        gw[j] = pack_bf16x2(d0, d1);
        const float2 rr = unpack_bf16x2(gw[j]);  // sum what the GEMMs will actually read
        acc[2 * j] += rr.x;
        acc[2 * j + 1] += rr.y;
      }
      *reinterpret_cast<uint4*>(dg + (int64_t)r * ld + col0) = make_uint4(gw[0], gw[1], gw[2], gw[3]);
    }
  }
  if (!colsum) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = threadIdx.x;
  if (blockIdx.x * 256 + c < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][c];
    atomicAdd(colsum + blockIdx.x * 256 + c, s);
  }
}

// cast fp32 [rows, N] (ld_src) -> bf16 [rows, N] (ld_dst), optional fused column sum of the
// rounded values (used for dQ: fp32 atomic accumulator -> bf16 operand + bias gradient)
__global__ void __launch_bounds__(256)
cast_f32_bf16_2d_kernel(const float* __restrict__ src, int64_t ld_src, __nv_bfloat16* __restrict__ dst,
                        int64_t ld_dst, int rows, int N, float* __restrict__ colsum, int rows_per_cta,
                        const int* __restrict__ rows_dyn) {
  __shared__ float red[8][128];
  if (rows_dyn) rows = min(rows, *rows_dyn);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col0 = blockIdx.x * 128 + lane * 4;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(r0 + rows_per_cta, rows);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col0 < N) {
    for (int r = r0 + warp; r < r1; r += 8) {
      const float4 v = *reinterpret_cast<const float4*>(src + (int64_t)r * ld_src + col0);
      const uint2 u = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      *reinterpret_cast<uint2*>(dst + (int64_t)r * ld_dst + col0) = u;
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
      acc.x += a.x; acc.y += a.y; acc.z += b.x; acc.w += b.y;
    }
  }
  if (!colsum) return;
  reinterpret_cast<float4*>(&red[warp][0])[lane] = acc;
  __syncthreads();
  const int c = threadIdx.x;
  if (c < 128 && blockIdx.x * 128 + c < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][c];
    atomicAdd(colsum + blockIdx.x * 128 + c, s);
  }
}

// flat fp32 -> bf16 cast (weight shadow refresh)
__global__ void __launch_bounds__(256)
cast_f32_bf16_flat_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

template <typename F>
static int dispatch_nv(int H, F&& f) {
  switch (H / 128) {
    case 1: return f(std::integral_constant<int, 1>{});
    case 2: return f(std::integral_constant<int, 2>{});
    case 3: return f(std::integral_constant<int, 3>{});
    case 4: return f(std::integral_constant<int, 4>{});
    case 6: return f(std::integral_constant<int, 6>{});
    case 8: return f(std::integral_constant<int, 8>{});
    case 10: return f(std::integral_constant<int, 10>{});
    default: return ERGM_ERR_UNSUPPORTED;
  }
}

static int row_grid(int rows) {
  const int need = (rows + ROW_WARPS - 1) / ROW_WARPS;
  const int cap = num_sms() * 8;
  return need < cap ? (need > 0 ? need : 1) : cap;
}

// ------------------------------------------------------------------------------------------
// Modality pooling (SURVEY §8 A3 extension): per-utterance feature SEQUENCES -> time mean.
// feature_extraction.py:63,69 pools audio [1,113,768] / key-frame visual [1,197,768] features to
// one vector offline; here the raw sequences [B, T, D] stay on the device and the mean is taken
// in-kernel, before the D -> H projection GEMM (mean and Linear commute, so pooling first makes
// the projection an M = B GEMM instead of an M = B*T one).  HBM-bound: B*T*D*4 bytes read once.
// grid (B, ceil(D/128)); 256 threads = 32 float4 column lanes x 8 time lanes; fixed-order
// reduction across the time lanes -> bitwise deterministic.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mm_pool_kernel(const float* __restrict__ seq, int64_t ld_b, int64_t ld_t, const int* __restrict__ lens,
               int T, int D, float* __restrict__ pooled_f32, int64_t ld_f32,
               __nv_bfloat16* __restrict__ pooled_bf16, int64_t ld_bf16) {
  __shared__ float4 red[8][32];
  const int b = blockIdx.x;
  const int cl = threadIdx.x & 31, tl = threadIdx.x >> 5;
  const int c4 = blockIdx.y * 32 + cl;  // float4 column
  const bool ok = c4 * 4 < D;
  int n = lens ? min(T, max(lens[b], 0)) : T;
  const float* base = seq + (int64_t)b * ld_b + 4 * (int64_t)c4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) {
    int t = tl;
    for (; t + 24 < n; t += 32) {  // 4 independent 16-byte loads in flight per thread
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(base + (int64_t)t * ld_t));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(base + (int64_t)(t + 8) * ld_t));
      const float4 v2 = __ldg(reinterpret_cast<const float4*>(base + (int64_t)(t + 16) * ld_t));
      const float4 v3 = __ldg(reinterpret_cast<const float4*>(base + (int64_t)(t + 24) * ld_t));
      acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
      acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; t < n; t += 8) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(base + (int64_t)t * ld_t));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  red[tl][cl] = acc;
  __syncthreads();
  if (tl == 0 && ok) {
#pragma unroll
    for (int i = 1; i < 8; ++i) {
      const float4 o = red[i][cl];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    const float inv = n > 0 ? 1.f / (float)n : 0.f;
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    if (pooled_f32) *reinterpret_cast<float4*>(pooled_f32 + (int64_t)b * ld_f32 + 4 * c4) = acc;
    if (pooled_bf16)
      *reinterpret_cast<uint2*>(pooled_bf16 + (int64_t)b * ld_bf16 + 4 * c4) =
          make_uint2(pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
  }
}

}  // namespace ergm

using namespace ergm;

extern "C" int ergm_embed_fuse_fwd(const int64_t* ids, const int64_t* token_type_ids,
                                   const int64_t* position_ids, int64_t pos_stride_b,
                                   const int* past_lens, const float* wte, const float* wpe,
                                   const float* imgs, int64_t ld_img, const float* auds,
                                   int64_t ld_aud, float* out, int B, int T, int H, int past_len,
                                   int vocab, int n_pos, float dropout_p, uint64_t seed,
                                   uint64_t offset, int* err_flag, const ergm_pack* pack, void* stream) {
  if (!ids || !wte || !wpe || !out || !err_flag || B <= 0 || T <= 0 || H % 128) return ERGM_ERR_ARG;
  if ((imgs && ld_img % 4) || (auds && ld_aud % 4)) return ERGM_ERR_ARG;
  if (pack && (!pack->row_b || !pack->row_t || !pack->n_rows)) return ERGM_ERR_ARG;
  EmbedFwdParams p{ids, token_type_ids, position_ids, pos_stride_b, past_lens, wte, wpe, imgs, auds, out, ld_img, ld_aud,
                   B, T, H, past_len, vocab, n_pos,
                   make_site(seed, offset, dropout_p, (uint32_t)H), dropout_p > 0.f, err_flag, ERGM_PACK_VIEW(pack)};
  embed_fuse_fwd_kernel<<<row_grid(B * T), ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}

extern "C" int ergm_gather_rows_bf16(const int64_t* ids, const float* table, void* out_bf16,
                                     int rows, int H, int vocab, int* err_flag, void* stream) {
  if (!ids || !table || !out_bf16 || !err_flag || rows <= 0 || H % 128) return ERGM_ERR_ARG;
  gather_rows_bf16_kernel<<<row_grid(rows), ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(
      ids, table, reinterpret_cast<__nv_bfloat16*>(out_bf16), rows, H, vocab, err_flag);
  return (int)cudaGetLastError();
}

extern "C" int ergm_embed_bwd(const float* dh, const int64_t* ids, const int64_t* token_type_ids,
                              const int64_t* position_ids, int64_t pos_stride_b, float* dwte, float* dwpe, float* dimgs,
                              float* dauds, int rows, int T, int H, int past_len, int vocab, int n_pos,
                              float dropout_p, uint64_t seed, uint64_t offset, int* err_flag, const ergm_pack* pack,
                              void* stream) {
  if (!dh || rows <= 0 || T <= 0 || H % 128 || H / 4 > 1024) return ERGM_ERR_ARG;
  if (pack && (!pack->row_b || !pack->row_t || !pack->n_rows)) return ERGM_ERR_ARG;
  if ((ids || token_type_ids) && (!dwte || vocab <= 0)) return ERGM_ERR_ARG;
  if (dwpe && n_pos <= 0) return ERGM_ERR_ARG;
  EmbedBwdParams p{dh, ids, token_type_ids, position_ids, pos_stride_b, dwte, dwpe, dimgs, dauds, rows, T, H,
                   past_len, 32, make_site(seed, offset, dropout_p, (uint32_t)H),
                   dropout_p > 0.f, vocab, n_pos, err_flag, ERGM_PACK_VIEW(pack)};
  const int threads = ((H / 4 + 31) / 32) * 32;
  embed_bwd_kernel<<<(rows + 31) / 32, threads, 0, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}

extern "C" int ergm_ln_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16,
                           float* y_f32, float* mean, float* rstd, int rows, int H, float eps,
                           const int* row_idx, const int* rows_dyn, void* stream) {
  if (!x || !gamma || !beta || rows <= 0 || H % 128) return ERGM_ERR_ARG;
  return dispatch_nv(H, [&](auto nv) {
    return (int)launch_pdl(ln_fwd_kernel<decltype(nv)::value>, dim3((unsigned)row_grid(rows)), dim3(ROW_WARPS * 32), 0,
                           (cudaStream_t)stream, 1, x, gamma, beta, reinterpret_cast<__nv_bfloat16*>(y_bf16), y_f32, mean,
                           rstd, rows, eps, row_idx, rows_dyn);
  });
}

extern "C" int ergm_ln_bwd(const void* dy, int dy_is_f32, const float* x, const float* mean,
                           const float* rstd, const float* gamma, const float* dres_in,
                           float* dx_out, void* dx_bf16, float* dgamma, float* dbeta,
                           float* dbias_next, int rows, int H, float dropout_p, uint64_t seed,
                           uint64_t offset, const int* rows_dyn, void* stream) {
  if (!dy || !x || !mean || !rstd || !gamma || !dgamma || !dbeta || rows <= 0 || H % 128)
    return ERGM_ERR_ARG;
  LnBwdParams p{dy, x, mean, rstd, gamma, dres_in, dx_out, reinterpret_cast<__nv_bfloat16*>(dx_bf16),
                dgamma, dbeta, dbias_next, rows, dy_is_f32,
                make_site(seed, offset, dropout_p, (uint32_t)H), dropout_p > 0.f, rows_dyn};
  if (reinterpret_cast<uintptr_t>(dy) & 15 || reinterpret_cast<uintptr_t>(x) & 15 ||
      (dres_in && (reinterpret_cast<uintptr_t>(dres_in) & 15)))
    return ERGM_ERR_ARG;
  const int stage_bytes = LNB_ROWS * H * ((dy_is_f32 ? 4 : 2) + 4 + (dres_in ? 4 : 0));
  int stages = (232448 - 128 - H * 4) / stage_bytes;
  if (stages > 4) stages = 4;
  const int n_tiles = (rows + LNB_ROWS - 1) / LNB_ROWS;
  int grid = n_tiles < num_sms() ? n_tiles : num_sms();
  const int per_cta = (n_tiles + grid - 1) / grid;
  if (stages > per_cta) stages = per_cta;
  if (stages < 1) return ERGM_ERR_UNSUPPORTED;
  int smem = 128 + H * 4 + stages * stage_bytes;
  const int red_bytes = 128 + H * 4 + 3 * LNB_ROWS * H * 4;  // column-sum reduction reuses the ring
  if (smem < red_bytes) smem = red_bytes;
  if (smem > 232448) return ERGM_ERR_UNSUPPORTED;
  auto launch = [&](auto kern) -> int {
    ERGM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    return (int)launch_pdl(kern, dim3((unsigned)grid), dim3(LNB_THREADS), (size_t)smem, (cudaStream_t)stream, 1, p, stages);
  };
  switch (H / 128) {
    case 1: return launch(ln_bwd_kernel<1>);
    case 2: return launch(ln_bwd_kernel<2>);
    case 3: return launch(ln_bwd_kernel<3>);
    case 4: return launch(ln_bwd_kernel<4>);
    case 6: return launch(ln_bwd_kernel<6>);
    case 8: return launch(ln_bwd_kernel<8>);
  }
  return ERGM_ERR_UNSUPPORTED;
}

extern "C" int ergm_colsum_bf16(const void* src, int64_t ld, int rows, int N, float* out,
                                void* stream) {
  if (!src || !out || rows <= 0 || N <= 0 || N % 8 || ld % 8) return ERGM_ERR_ARG;
  const int col_ctas = (N + 255) / 256;
  int row_splits = (2 * num_sms() + col_ctas - 1) / col_ctas;
  if (row_splits > (rows + 63) / 64) row_splits = (rows + 63) / 64;
  const int rpc = (rows + row_splits - 1) / row_splits;
  colsum_bf16_kernel<<<dim3(col_ctas, (rows + rpc - 1) / rpc), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), ld, rows, N, out, rpc);
  return (int)cudaGetLastError();
}

extern "C" int ergm_gelu_bwd_colsum(void* dg_bf16, const void* u_bf16, int64_t ld, int rows, int N,
                                    float* colsum, int exact, const int* rows_dyn, void* stream) {
  if (!dg_bf16 || !u_bf16 || rows <= 0 || N <= 0 || N % 8 || ld % 8) return ERGM_ERR_ARG;
  const int col_ctas = (N + 255) / 256;
  int row_splits = (4 * num_sms() + col_ctas - 1) / col_ctas;
  if (row_splits > (rows + 63) / 64) row_splits = (rows + 63) / 64;
  const int rpc = (rows + row_splits - 1) / row_splits;
  gelu_bwd_colsum_kernel<<<dim3(col_ctas, (rows + rpc - 1) / rpc), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<__nv_bfloat16*>(dg_bf16), reinterpret_cast<const __nv_bfloat16*>(u_bf16), ld, rows, N,
      colsum, rpc, exact, rows_dyn);
  return (int)cudaGetLastError();
}

extern "C" int ergm_cast_f32_bf16_2d(const float* src, int64_t ld_src, void* dst, int64_t ld_dst,
                                     int rows, int N, float* colsum, const int* rows_dyn, void* stream) {
  if (!src || !dst || rows <= 0 || N <= 0 || N % 4 || ld_src % 4 || ld_dst % 4) return ERGM_ERR_ARG;
  const int col_ctas = (N + 127) / 128;
  int row_splits = (2 * num_sms() + col_ctas - 1) / col_ctas;
  if (row_splits > (rows + 63) / 64) row_splits = (rows + 63) / 64;
  const int rpc = (rows + row_splits - 1) / row_splits;
  cast_f32_bf16_2d_kernel<<<dim3(col_ctas, (rows + rpc - 1) / rpc), 256, 0, (cudaStream_t)stream>>>(
      src, ld_src, reinterpret_cast<__nv_bfloat16*>(dst), ld_dst, rows, N, colsum, rpc, rows_dyn);
  return (int)cudaGetLastError();
}

extern "C" int ergm_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (!src || !dst || n <= 0 || n % 4) return ERGM_ERR_ARG;
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  cast_f32_bf16_flat_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      src, reinterpret_cast<__nv_bfloat16*>(dst), n4);
  return (int)cudaGetLastError();
}

extern "C" int ergm_mm_pool_fwd(const float* seq, int64_t ld_b, int64_t ld_t, const int* lens, int B,
                                int T, int D, float* pooled_f32, int64_t ld_f32, void* pooled_bf16,
                                int64_t ld_bf16, void* stream) {
  if (!seq || B <= 0 || T <= 0 || D <= 0 || D % 4 || (!pooled_f32 && !pooled_bf16)) return ERGM_ERR_ARG;
  if (ld_b % 4 || ld_t % 4 || (reinterpret_cast<uintptr_t>(seq) & 15)) return ERGM_ERR_ARG;
  if (pooled_f32 && (ld_f32 % 4 || (reinterpret_cast<uintptr_t>(pooled_f32) & 15))) return ERGM_ERR_ARG;
  if (pooled_bf16 && (ld_bf16 % 4 || (reinterpret_cast<uintptr_t>(pooled_bf16) & 7))) return ERGM_ERR_ARG;
  mm_pool_kernel<<<dim3(B, (D + 127) / 128), 256, 0, (cudaStream_t)stream>>>(
      seq, ld_b, ld_t, lens, T, D, pooled_f32, ld_f32, reinterpret_cast<__nv_bfloat16*>(pooled_bf16), ld_bf16);
  return (int)cudaGetLastError();
}

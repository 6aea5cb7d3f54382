// Counter-based dropout masks.  A dropout *site* (one nn.Dropout application in the
// reference: model.py:142 attn_dropout, :245 / :266 resid dropout, :506 embd dropout) is
// identified by (seed, offset); the keep decision of element (row, col) of that site's
// [rows, ncols] tensor is a pure function of (seed, offset, row, col), so forward and
// backward kernels regenerate the same mask without storing it.
#pragma once
#include "common.cuh"

namespace ergm {

struct DropoutSite {
  uint64_t seed, offset;
  float p;
  uint32_t ncol4;  // ceil(ncols / 4)
  uint32_t key;    // (seed, offset [, step]) pre-mixed once per kernel: see resolved()
  // Optional device-resident step counter added to the seed, so that a CUDA-graph replay
  // (whose kernel arguments are frozen) still draws fresh masks every step.
  const uint64_t* step;

  // call once at kernel entry: folds *step into the seed
  ERGM_DEVINL DropoutSite resolved() const {
    DropoutSite r = *this;
    if (step) r.seed += __ldg(reinterpret_cast<const unsigned long long*>(step));
    r.step = nullptr;
    r.key = mix_key(r.seed, r.offset);
    return r;
  }

  // Philox4x32-10 variant of keep4 (reference-quality generator; kept for tests / comparison)
  ERGM_DEVINL uint32_t keep4_philox(uint32_t row, uint32_t col4) const {
    Philox ph(seed, offset);
    const uint4 r = ph((uint64_t)row * ncol4 + col4);
    uint32_t m = 0;
    m |= (u01(r.x) >= p) ? 1u : 0u;
    m |= (u01(r.y) >= p) ? 2u : 0u;
    m |= (u01(r.z) >= p) ? 4u : 0u;
    m |= (u01(r.w) >= p) ? 8u : 0u;
    return m;
  }
  // keep bits for elements (row, 4*col4 .. 4*col4+3); bit i set = keep.  Two avalanche hashes give
  // four 16-bit uniforms (see hash2 below): ~7 integer instructions per element instead of ~25 for
  // Philox-10, which matters because the masks are regenerated inside GEMM / LayerNorm epilogues.
  ERGM_DEVINL uint32_t keep4(uint32_t row, uint32_t col4) const {
    const uint32_t t = thr16();
    const uint32_t h0 = hash2(row, 2u * col4), h1 = hash2(row, 2u * col4 + 1u);
    uint32_t m = 0;
    m |= ((h0 & 0xffffu) >= t) ? 1u : 0u;
    m |= ((h0 >> 16) >= t) ? 2u : 0u;
    m |= ((h1 & 0xffffu) >= t) ? 4u : 0u;
    m |= ((h1 >> 16) >= t) ? 8u : 0u;
    return m;
  }
  ERGM_DEVINL bool keep(uint32_t row, uint32_t col) const { return keep_hash(row, col); }
  // realised keep probability of the 16-bit threshold (use this, not 1-p, to rescale kept values)
  ERGM_DEVINL float keep_scale() const { return 65536.f / (65536.f - (float)thr16()); }

  // ---- cheap per-element generator for the attention-probability dropout (model.py:142) ----
  // The [B*nh*Tq, Tk] probability tile is two orders of magnitude larger than any other dropout
  // site and is visited once per element in BOTH forward and backward, so Philox-10 would cost
  // more than the softmax itself.  One 32-bit avalanche hash (two multiply-xorshift rounds on a
  // counter keyed by seed/offset, "lowbias32" constants) yields two independent 16-bit uniforms:
  // element (row, col) is kept iff its 16-bit lane >= thr16 = round(p * 65536).
  ERGM_DEVINL uint32_t thr16() const { return (uint32_t)(p * 65536.f + 0.5f); }
  // key: two avalanche rounds over (seed, offset), computed ONCE per kernel by resolved(); the
  // per-element work is then a single round on a Weyl-spread counter (8 integer instructions per
  // two elements - the first version re-mixed seed and offset for every pair: 18).
  static __host__ __device__ __forceinline__ uint32_t avalanche(uint32_t h) {
    h ^= h >> 16; h *= 0x7feb352du;
    h ^= h >> 15; h *= 0x846ca68bu;
    h ^= h >> 16;
    return h;
  }
  static __host__ __device__ __forceinline__ uint32_t mix_key(uint64_t seed, uint64_t offset) {
    uint32_t k = avalanche((uint32_t)seed ^ 0x9E3779B9u);
    k = avalanche(k + (uint32_t)(seed >> 32) * 0x85EBCA6Bu);
    k = avalanche(k ^ ((uint32_t)offset * 0x9E3779B9u + (uint32_t)(offset >> 32)));
    return k;
  }
  ERGM_DEVINL uint32_t hash2(uint32_t row, uint32_t col2) const {
    return avalanche((row * ncol4 * 2u + col2) * 0x9E3779B1u + key);
  }
  ERGM_DEVINL bool keep_hash(uint32_t row, uint32_t col) const {
    const uint32_t h = hash2(row, col >> 1);
    return ((col & 1u) ? (h >> 16) : (h & 0xffffu)) >= thr16();
  }
};

// host side: library-global step pointer (set by ergm_set_rng_step_ptr, see api.cu)
extern const uint64_t* g_rng_step_ptr;
inline DropoutSite make_site(uint64_t seed, uint64_t offset, float p, uint32_t ncols) {
  return DropoutSite{seed, offset, p, (ncols + 3) / 4, DropoutSite::mix_key(seed, offset),
                     p > 0.f ? g_rng_step_ptr : nullptr};
}

}  // namespace ergm

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

  // keep bits for elements (row, 4*col4 .. 4*col4+3); bit i set = keep
  ERGM_DEVINL uint32_t keep4(uint32_t row, uint32_t col4) const {
    Philox ph(seed, offset);
    const uint4 r = ph((uint64_t)row * ncol4 + col4);
    uint32_t m = 0;
    m |= (u01(r.x) >= p) ? 1u : 0u;
    m |= (u01(r.y) >= p) ? 2u : 0u;
    m |= (u01(r.z) >= p) ? 4u : 0u;
    m |= (u01(r.w) >= p) ? 8u : 0u;
    return m;
  }
  ERGM_DEVINL bool keep(uint32_t row, uint32_t col) const {
    return (keep4(row, col >> 2) >> (col & 3)) & 1u;
  }
};

}  // namespace ergm

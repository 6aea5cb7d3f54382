// Shared device helpers for the ergm_b200 sm_100a kernels: raw PTX wrappers for
// mbarrier / TMA / tcgen05 / TMEM, a counter-based Philox RNG for dropout, and
// small numeric utilities.  Everything here is hand-written inline PTX; nothing
// depends on CUTLASS.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

#define ERGM_DEVINL __device__ __forceinline__

namespace ergm {

// ----------------------------------------------------------------------------
// shared-memory address helpers
// ----------------------------------------------------------------------------
ERGM_DEVINL uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

ERGM_DEVINL uint32_t lane_id() { return threadIdx.x & 31; }

ERGM_DEVINL bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------
ERGM_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
ERGM_DEVINL void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
ERGM_DEVINL void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
ERGM_DEVINL void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
ERGM_DEVINL void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
ERGM_DEVINL bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
ERGM_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ----------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) loads, 2-D and 3-D, completing on an mbarrier
// ----------------------------------------------------------------------------
ERGM_DEVINL void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
ERGM_DEVINL void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
ERGM_DEVINL void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1,
                             int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// TMA store of one smem tile (bulk async-group completion): smem must have been written + fence.proxy.async'ed
ERGM_DEVINL void tma_store_3d(const void* tmap, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// TMA reduction: global[box] += smem tile (element type of the tensor map; fp32 here), same completion mechanism
ERGM_DEVINL void tma_reduce_add_3d(const void* tmap, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
ERGM_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING their smem source (the global writes may still be in flight)
ERGM_DEVINL void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
ERGM_DEVINL void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ----------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------
ERGM_DEVINL void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
}
ERGM_DEVINL void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
ERGM_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
ERGM_DEVINL void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
ERGM_DEVINL void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16/fp16 inputs, fp32 accumulate.
ERGM_DEVINL void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
ERGM_DEVINL void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma has completed.
ERGM_DEVINL void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
ERGM_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
ERGM_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 columns of fp32: thread t of the warp receives lane (base_lane + t),
// columns [col, col+32).
ERGM_DEVINL void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
ERGM_DEVINL void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 columns store (used to place bf16-packed P tiles into TMEM).
ERGM_DEVINL void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
      "r"(r[15])
      : "memory");
}

ERGM_DEVINL void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
      "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}

// ----------------------------------------------------------------------------
// thread-block clusters and CTA pairs (tcgen05 cta_group::2)
// ----------------------------------------------------------------------------
ERGM_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
ERGM_DEVINL void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
ERGM_DEVINL void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
ERGM_DEVINL void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
ERGM_DEVINL void st_cluster_v2(uint32_t cluster_addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(cluster_addr), "r"(a), "r"(b) : "memory");
}
// Programmatic dependent launch (PDL): launch_dependents lets the NEXT kernel of the stream start its
// independent prologue (barrier init, weight / KV-cache prefetch) while this one is still running;
// pdl_wait blocks until every kernel this one depends on has completed and flushed.  Both are no-ops
// for kernels launched without the programmatic-serialization attribute.
ERGM_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
ERGM_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank`
ERGM_DEVINL uint32_t mapa_cluster(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
ERGM_DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair; the transaction bytes are counted on `bar`, a
// shared::cluster address that may live in the peer (leader) CTA.
ERGM_DEVINL void tma_load_2d_2sm(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
ERGM_DEVINL void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
}
ERGM_DEVINL void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
ERGM_DEVINL void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem, both CTAs] (+)= A * B with M = 256 split over the CTA pair; issued by the leader CTA only.
ERGM_DEVINL void umma_ss_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on the mbarrier at this shared offset in every CTA of `mask` once all prior MMAs retire.
ERGM_DEVINL void umma_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}

// ----------------------------------------------------------------------------
// UMMA descriptors (bit layout: see DESIGN.md "tcgen05 descriptors")
// ----------------------------------------------------------------------------
// Shared-memory matrix descriptor, 128-byte swizzle, version 1 (sm_100).
//   bits [ 0,14) start address >> 4      bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4 bits [46,48) version = 1
//   bits [61,64) layout type (2 = SWIZZLE_128B)
ERGM_DEVINL uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulate.
//   [4,6) c_format=1 (f32)  [7,10) a_format=1 (bf16)  [10,13) b_format=1 (bf16)
//   [15] a_major (0=K,1=MN) [16] b_major  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ----------------------------------------------------------------------------
// Philox4x32-10, counter based: (seed, subsequence=offset, counter=index/4)
// ----------------------------------------------------------------------------
struct Philox {
  uint32_t k0, k1;
  uint64_t sub;
  ERGM_DEVINL Philox(uint64_t seed, uint64_t subsequence)
      : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), sub(subsequence) {}
  ERGM_DEVINL uint4 operator()(uint64_t ctr) const {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = (uint32_t)sub,
             c3 = (uint32_t)(sub >> 32);
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
// keep-decision for element `idx` of a dropout site: uniform in [0,1) >= p
ERGM_DEVINL float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// ----------------------------------------------------------------------------
// numerics
// ----------------------------------------------------------------------------
// 2^x for x <= 0 (softmax exponents): one MUFU.EX2, denormal results flush to zero
ERGM_DEVINL float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
ERGM_DEVINL float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// gelu_new (tanh form), transformers NewGELUActivation; reference: model.py:259,264
template <bool kExact>
ERGM_DEVINL float gelu_new(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * x * fmaf(k1, x * x, 1.0f);
  float t = kExact ? tanhf(inner) : tanh_fast(inner);
  return 0.5f * x * (1.0f + t);
}
template <bool kExact>
ERGM_DEVINL float gelu_new_grad(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float x2 = x * x;
  float inner = k0 * x * fmaf(k1, x2, 1.0f);
  float t = kExact ? tanhf(inner) : tanh_fast(inner);
  float dinner = k0 * fmaf(3.0f * k1, x2, 1.0f);
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * dinner;
}

ERGM_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
ERGM_DEVINL float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

ERGM_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
ERGM_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace ergm

// ----------------------------------------------------------------------------
// Packed variable-length batches (collate of custom_dataset.py:102-132 right-pads every sample to
// the batch maximum and model.py then computes every pad position): with an ergm_pack the row kernels and GEMMs
// work on the concatenation of the samples' real rows.  Device view of include/ergm_b200.h's ergm_pack.
// ----------------------------------------------------------------------------
namespace ergm {
struct PackView {
  const int* cu;       // [B + 1] first packed row of every sample
  const int* row_b;    // [capacity] sample of a packed row
  const int* row_t;    // [capacity] position (column of the padded [B, T] layout) of a packed row
  const int* n_rows;   // [1] packed rows of this batch (run-time value)
  const int* kv_lens;  // [B] attendable keys per sample (its real tokens)
  ERGM_DEVINL bool on() const { return row_b != nullptr; }
};
}  // namespace ergm
#define ERGM_PACK_VIEW(pack) \
  ((pack) ? ergm::PackView{(pack)->cu_rows, (pack)->row_b, (pack)->row_t, (pack)->n_rows, (pack)->kv_lens} \
          : ergm::PackView{nullptr, nullptr, nullptr, nullptr, nullptr})

// ----------------------------------------------------------------------------
// host-side helpers shared by the C-ABI translation units
// ----------------------------------------------------------------------------
#define ERGM_OK 0
#define ERGM_ERR_ARG (-1)
#define ERGM_ERR_UNSUPPORTED (-2)
#define ERGM_ERR_DRIVER (-3)

#define ERGM_CUDA_TRY(expr)                   \
  do {                                        \
    cudaError_t _e = (expr);                  \
    if (_e != cudaSuccess) return (int)_e;    \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is per function AND per device: remember which devices
// have it in an atomic bit mask (two host threads racing here both set the same value: harmless).
#define ERGM_SET_SMEM_ATTR(kernel, bytes)                                                              \
  do {                                                                                                 \
    static std::atomic<uint64_t> _done_mask{0};                                                        \
    int _dev = 0;                                                                                      \
    ERGM_CUDA_TRY(cudaGetDevice(&_dev));                                                               \
    const uint64_t _bit = 1ull << (_dev & 63);                                                         \
    if (!(_done_mask.load(std::memory_order_acquire) & _bit)) {                                        \
      ERGM_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); \
      _done_mask.fetch_or(_bit, std::memory_order_release);                                            \
    }                                                                                                  \
  } while (0)

namespace ergm {
// Kernel launch with the PDL attribute (and an optional cluster dimension along x).
// ERGM_PDL=0 in the environment turns the attribute off (plain stream order).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  unsigned n = 0;
  if (pdl_enabled()) {
    attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attrs[n].id = cudaLaunchAttributeClusterDimension;
    attrs[n].val.clusterDim.x = (unsigned)cluster_x; attrs[n].val.clusterDim.y = 1; attrs[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attrs; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
// Encodes a 2-D bf16/fp32 tiled tensor map (128B swizzle).  dims/strides follow the
// cuTensorMapEncodeTiled convention: dim0 is the contiguous one.
int encode_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t dim0,
                   uint64_t dim1, uint64_t stride1_bytes, uint32_t box0, uint32_t box1);
int encode_tmap_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t dim0,
                   uint64_t dim1, uint64_t dim2, uint64_t stride1_bytes, uint64_t stride2_bytes,
                   uint32_t box0, uint32_t box1, uint32_t box2);
int num_sms();
}  // namespace ergm

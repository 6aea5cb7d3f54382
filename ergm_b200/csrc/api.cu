// Library probes of the C ABI (no compute).
#include "../../include/ergm_b200.h"
#include "common.cuh"
#include "dropout.cuh"

namespace ergm { const uint64_t* g_rng_step_ptr = nullptr; }

__global__ void rng_step_advance_kernel(unsigned long long* p, unsigned long long inc) { *p += inc; }

// Registers a device uint64 that every dropout kernel adds to its seed (NULL clears it).
extern "C" int ergm_set_rng_step_ptr(const uint64_t* dev_ptr) {
  ergm::g_rng_step_ptr = dev_ptr;
  return ERGM_OK;
}
extern "C" int ergm_rng_step_advance(uint64_t* dev_ptr, uint64_t inc, void* stream) {
  if (!dev_ptr) return ERGM_ERR_ARG;
  rng_step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(dev_ptr), inc);
  return (int)cudaGetLastError();
}

namespace ergm { void tmap_cache_stats(uint64_t* hits, uint64_t* misses); }
extern "C" int ergm_tmap_cache_stats(uint64_t* hits, uint64_t* misses) {
  if (!hits || !misses) return ERGM_ERR_ARG;
  ergm::tmap_cache_stats(hits, misses);
  return ERGM_OK;
}

extern "C" int ergm_abi_version(void) { return 2; }
extern "C" int ergm_device_sm_count(void) { return ergm::num_sms(); }

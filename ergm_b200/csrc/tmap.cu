// Host-side TMA tensor-map encoding.  The driver entry point is resolved at run
// time through the CUDA runtime so the library does not link against libcuda.
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include <cstdlib>

namespace ergm {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// Descriptor cache: a tensor map depends only on (base, element size, dims, strides, box),
// all of which are static per call site once the workspaces exist, so each distinct map is encoded once
// per process and afterwards copied (128 bytes) out of a mutex-guarded table.  Safe from any host thread.
struct TmapKey {
  uint64_t v[10];
  bool operator==(const TmapKey& o) const { return std::memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (uint64_t x : k.v) { h ^= x + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); }
    return (size_t)h;
  }
};
static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
static uint64_t g_tmap_hits = 0, g_tmap_misses = 0;

static bool tmap_lookup(const TmapKey& k, CUtensorMap* out) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  auto it = g_tmap_cache.find(k);
  if (it == g_tmap_cache.end()) { ++g_tmap_misses; return false; }
  ++g_tmap_hits;
  *out = it->second;
  return true;
}
static void tmap_insert(const TmapKey& k, const CUtensorMap& m) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  if (g_tmap_cache.size() > 16384) g_tmap_cache.clear();  // shape-polymorphic callers: bounded memory
  g_tmap_cache.emplace(k, m);
}
void tmap_cache_stats(uint64_t* hits, uint64_t* misses) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  *hits = g_tmap_hits; *misses = g_tmap_misses;
}

static CUtensorMapDataType dtype_of(int elem_bytes) {
  return elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
}

int encode_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t dim0,
                   uint64_t dim1, uint64_t stride1_bytes, uint32_t box0, uint32_t box1) {
  EncodeTiledFn fn = get_encode();
  if (!fn) return ERGM_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (stride1_bytes & 15)) return ERGM_ERR_ARG;
  const TmapKey key{{(uint64_t)reinterpret_cast<uintptr_t>(base), (uint64_t)elem_bytes | (2ull << 32), dim0, dim1, 0,
                     stride1_bytes, 0, box0, box1, 0}};
  if (tmap_lookup(key, out)) return ERGM_OK;
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {stride1_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dtype_of(elem_bytes), 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return ERGM_ERR_DRIVER;
  tmap_insert(key, *out);
  return ERGM_OK;
}

int encode_tmap_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t dim0,
                   uint64_t dim1, uint64_t dim2, uint64_t stride1_bytes, uint64_t stride2_bytes,
                   uint32_t box0, uint32_t box1, uint32_t box2) {
  EncodeTiledFn fn = get_encode();
  if (!fn) return ERGM_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (stride1_bytes & 15) || (stride2_bytes & 15))
    return ERGM_ERR_ARG;
  const TmapKey key{{(uint64_t)reinterpret_cast<uintptr_t>(base), (uint64_t)elem_bytes | (3ull << 32), dim0, dim1, dim2,
                     stride1_bytes, stride2_bytes, box0, box1, box2}};
  if (tmap_lookup(key, out)) return ERGM_OK;
  cuuint64_t dims[3] = {dim0, dim1, dim2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, dtype_of(elem_bytes), 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return ERGM_ERR_DRIVER;
  tmap_insert(key, *out);
  return ERGM_OK;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("ERGM_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace ergm

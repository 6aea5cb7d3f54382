// Host-side TMA tensor-map encoding.  The driver entry point is resolved at run
// time through the CUDA runtime so the library does not link against libcuda.
#include <mutex>

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

static CUtensorMapDataType dtype_of(int elem_bytes) {
  return elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
}

int encode_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t dim0,
                   uint64_t dim1, uint64_t stride1_bytes, uint32_t box0, uint32_t box1) {
  EncodeTiledFn fn = get_encode();
  if (!fn) return ERGM_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (stride1_bytes & 15)) return ERGM_ERR_ARG;
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {stride1_bytes};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dtype_of(elem_bytes), 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? ERGM_OK : ERGM_ERR_DRIVER;
}

int encode_tmap_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t dim0,
                   uint64_t dim1, uint64_t dim2, uint64_t stride1_bytes, uint64_t stride2_bytes,
                   uint32_t box0, uint32_t box1, uint32_t box2) {
  EncodeTiledFn fn = get_encode();
  if (!fn) return ERGM_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (stride1_bytes & 15) || (stride2_bytes & 15))
    return ERGM_ERR_ARG;
  cuuint64_t dims[3] = {dim0, dim1, dim2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, dtype_of(elem_bytes), 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? ERGM_OK : ERGM_ERR_DRIVER;
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

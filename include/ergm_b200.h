/*
 * ergm_b200 C ABI — the drop-in boundary of the B200-native ERGM hot path.
 *
 * The reference (LovesickPatience/ERGM) has no FFI layer: its model file
 * src/model.py reaches the GPU only through PyTorch library ops.  Each entry
 * point below replaces the library-op call sites of one stage of
 * src/model.py (file:line cited per function).  Conventions:
 *   - plain C: raw device pointers, explicit sizes / leading dimensions,
 *     a cudaStream_t (passed as void*), no C++ or torch types;
 *   - every function is asynchronous on `stream`, never allocates, never
 *     synchronises, and returns 0 on success, a negative ERGM_ERR_* code for
 *     an argument error or a positive cudaError_t;
 *   - bf16 tensors are `uint16_t`-sized, row-major, leading dimensions are in
 *     elements.
 */
#ifndef ERGM_B200_H_
#define ERGM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ERGM_OK 0
#define ERGM_ERR_ARG (-1)
#define ERGM_ERR_UNSUPPORTED (-2)
#define ERGM_ERR_DRIVER (-3)

/* library / device probes (no compute) */
int ergm_abi_version(void);
int ergm_device_sm_count(void);

/* ------------------------------------------------------------------------ */
/* GEMM: D[M,N] = epilogue(A[M,K] * B[K,N])  — bf16 operands, fp32 accumulate
 * in TMEM (tcgen05.mma, TMA-fed).  Replaces transformers Conv1D.forward
 * (addmm) used at model.py:218,219,222,244,263,265, nn.Linear lm_head at
 * model.py:698 and every dgrad / wgrad GEMM autograd derives from them.     */
/* ------------------------------------------------------------------------ */
enum {
  ERGM_MAJOR_K = 0,  /* operand is stored [rows(M or N), K], K contiguous      */
  ERGM_MAJOR_MN = 1  /* operand is stored [K, rows(M or N)], M/N contiguous    */
};
enum {
  ERGM_EPI_BIAS = 1,       /* + bias[N] (fp32)                                  */
  ERGM_EPI_GELU = 2,       /* gelu_new(x) (model.py:259,264)                    */
  ERGM_EPI_RESIDUAL = 4,   /* + residual[M,N] (fp32) (model.py:309,328,334)     */
  ERGM_EPI_ATOMIC = 8,     /* D += result via red.add.f32 (wgrad accumulate)    */
  ERGM_EPI_DROPOUT = 16,   /* inverted dropout on (acc+bias) before residual    */
  ERGM_EPI_PREACT = 32,    /* also store (acc+bias) before GELU to `preact`     */
  ERGM_EPI_EXACT = 64      /* exact tanhf instead of tanh.approx in GELU        */
};
enum { ERGM_DT_BF16 = 0, ERGM_DT_F32 = 1 };

typedef struct ergm_gemm_args {
  const void* a;       /* bf16 */
  const void* b;       /* bf16 */
  void* d;             /* bf16 or fp32, row-major [M, ldd] */
  const float* bias;   /* [N] or NULL */
  const float* residual; /* fp32 [M, ldr] or NULL */
  void* preact;        /* bf16 [M, ldd] or NULL */
  int64_t lda, ldb, ldd, ldr;
  int32_t M, N, K;
  int32_t a_major, b_major;
  int32_t d_dtype;
  int32_t epilogue;    /* ERGM_EPI_* bitmask */
  int32_t split_k;     /* >=1; >1 requires ERGM_EPI_ATOMIC and fp32 D */
  int32_t block_n;     /* 0 = auto, else 64 / 128 / 256 */
  float dropout_p;
  uint64_t seed, offset; /* Philox key / subsequence of this dropout site */
} ergm_gemm_args;

int ergm_gemm_bf16(const ergm_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ERGM_B200_H_ */

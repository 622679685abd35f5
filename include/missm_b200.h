/* missm_b200 -- C ABI of the B200-native (sm_100a) hot path of MissM-Benchmark.
 *
 * The reference (Fieldhunter/MissM-Benchmark) is pure Python and has no FFI of its own; this
 * header is the boundary a maintainer binds with ctypes.  Each entry
 * point cites the reference code whose arithmetic it replaces (file:line under the reference
 * tree; "transformers 4.3x" = the unvendored third-party CLIP modules it imports at
 * languagebind/image/modeling_image.py:11-12).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises, nothing allocates
 *     device memory, there is no global mutable state besides the last-error string
 *     (thread-local), so calls are re-entrant (autograd worker threads);
 *   - return value 0 = ok; non-zero = error, text in missm_last_error();
 *   - bf16 tensors are row-major with the stated leading dimension in ELEMENTS.
 */
#ifndef MISSM_B200_H_
#define MISSM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MISSM_ABI_VERSION 1

int missm_version(void);
const char* missm_last_error(void);

/* ---------------------------------------------------------------------------------------
 * tcgen05 GEMM:  C[m,n] = epilogue( sum_k A[m,k] * B[n,k] )     bf16 in, fp32 accumulate
 *
 * Replaces every nn.Linear on the path (transformers 4.3x CLIPAttention q/k/v/out_proj and
 * CLIPMLP fc1/fc2, called at languagebind/image/modeling_image.py:121,133,140,150; the
 * projection at languagebind/__init__.py:79) and their autograd dgrad/wgrad twins, and the
 * patch-embedding Conv2d (video/modeling_video.py:29-35,45) as a GEMM over patch rows.
 *
 *   a_mn = 0: A stored [M,K] (K contiguous, lda);  a_mn = 1: A stored [K,M] (M contiguous)
 *   b_mn = 0: B stored [N,K] (K contiguous, ldb);  b_mn = 1: B stored [K,N] (N contiguous)
 * so y = x W^T is (a_mn=0,b_mn=0), dgrad dX = dY W is (0,1) on the same W, and
 * wgrad dW = dY^T X is (1,1) -- no operand is ever transposed in memory.
 * ------------------------------------------------------------------------------------- */
enum missm_epilogue {
  MISSM_EPI_LINEAR = 0, /* C = (acc + bias[n]) * (n < scale_cols ? col_scale : 1)          */
  MISSM_EPI_GELU = 1,   /* u = acc + bias; aux_out = bf16(u); C = bf16(u * sigmoid(1.702u)) */
  MISSM_EPI_RESID = 2,  /* C(f32) = aux_in(f32)[m,n] + acc + bias[n]   (may alias C)        */
  MISSM_EPI_DGELU = 3,  /* C = bf16(acc * quickgelu'(aux_in(bf16)[m,n]))                    */
  MISSM_EPI_PATCH = 4   /* C(f32)[(m/P)*(P+1)+1+m%P, n] = acc + aux_in(f32)[1+m%P, n]       */
};

typedef struct missm_gemm_args {
  const void* A; /* bf16 */
  const void* B; /* bf16 */
  void* C;       /* bf16 or f32 (out_f32) */
  const float* bias;  /* [N] or NULL */
  const void* aux_in; /* see epilogue */
  void* aux_out;      /* see epilogue */
  int32_t M, N, K;
  int32_t lda, ldb, ldc;
  int32_t ld_aux_in, ld_aux_out;
  int32_t a_mn, b_mn;
  int32_t epilogue;
  int32_t out_f32;    /* 1: C is float32, 0: C is bf16 */
  int32_t scale_cols; /* LINEAR: columns [0,scale_cols) are multiplied by col_scale */
  float col_scale;
  int32_t patch_P;    /* PATCH: patches per sample */
  int32_t split_k;    /* 0 = auto (only LINEAR + out_f32 + no bias may split), 1 = never */
  int32_t force_bn;   /* 0 = auto, 128 or 256 = force tile N */
} missm_gemm_args;

int missm_gemm_bf16(const missm_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MISSM_B200_H_ */

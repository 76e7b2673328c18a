/* inverse_gpu.h -- the reference's batched-inverse GPU API, served by the B200 engine.
 *
 * Drop-in for reference include/inverse_gpu.h:7-31: the same 17 `extern "C"` symbols with the
 * same argument order and meaning.  Like the reference header it may be included after
 * <cublas_v2.h> and "types.h"; unlike it, it also works stand-alone (it declares the opaque
 * `cublasHandle_t` itself when cuBLAS was not included).  The handle is accepted and ignored:
 * nothing in this library calls cuBLAS, pass NULL if you have none.
 *
 * Two calling flavours:
 *   NAME_gpu(handle, n, As, aInvs, batchSize)
 *       HOST pointers, `batchSize` dense column-major n x n matrices back to back (lda = n).
 *       Synchronous and self contained: H2D, compute, D2H.  `As` is never written (the
 *       reference's Cholesky wrappers overwrite it with the factor,
 *       src/inverse_cholesky_gpu.cu:442,672,747 -- a side effect no caller uses and that
 *       corrupts inverse_bench's shared input; deliberately dropped).
 *       A singular / non-SPD matrix aborts the process with the reference CPU path's message
 *       (src/inverse.c:64,94); use the *_ex entry points of invgpu.h for per-matrix info[].
 *   NAME_device(handle, N, devAs, devAInvs, batchSize)
 *       arrays of per-matrix DEVICE pointers (lda = N).  The pointer arrays themselves may
 *       live in pinned host memory (as in the reference, src/gauss/batched_invert.cu:120) or in
 *       device memory.  Asynchronous on the legacy default stream, no sync before return.
 *
 * Which engine path serves which symbol is listed next to each prototype; reference
 * definitions are cited as file:line under /root/reference.
 */
#ifndef INVGPU_INVERSE_GPU_H
#define INVGPU_INVERSE_GPU_H

#include "types.h"

#ifndef CUBLAS_API_H_
struct cublasContext;
typedef struct cublasContext *cublasHandle_t;
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define INVGPU_HOST_ENTRY(name)   void name(cublasHandle_t handle, int n, Array As, Array aInvs, int batchSize)
#define INVGPU_DEVICE_ENTRY(name) void name(cublasHandle_t handle, int N, Array *devAs, Array *devAInvs, int batchSize)

/* General matrices: Gauss-Jordan with partial pivoting (one launch).
 * src/gauss/batched_invert.cu:99 (3N launches) and src/gauss/inverse_gpu.cu:60,16 (cuBLAS LU). */
INVGPU_HOST_ENTRY(inverse_gauss_batched_gpu);
INVGPU_HOST_ENTRY(inverse_lu_cuda_batched_gpu);
INVGPU_DEVICE_ENTRY(inverse_gauss_batched_device);     /* declared but never defined upstream (inverse_gpu.h:10) */
INVGPU_DEVICE_ENTRY(inverse_lu_cuda_batched_device);   /* devAs is left intact (cuBLAS leaves LU factors there) */

/* SPD matrices: Cholesky potrf -> trtri -> lauum, full symmetric inverse written.
 * src/inverse_cholesky_gpu.cu:397 (simple), :625 (mm), :699 (mm2), :200 (stride). */
INVGPU_HOST_ENTRY(inverse_cholesky_batched_gpu);
INVGPU_HOST_ENTRY(inverse_cholesky_mm_batched_gpu);
INVGPU_HOST_ENTRY(inverse_cholesky_mm2_batched_gpu);
INVGPU_HOST_ENTRY(inverse_cholesky_stride_batched_gpu);

/* devAs -> devAInvs (devAs untouched).  src/inverse_cholesky_gpu.cu:323, :607, :692. */
INVGPU_DEVICE_ENTRY(inverse_cholesky_batched_device);
INVGPU_DEVICE_ENTRY(inverse_cholesky_mm_batched_device);
INVGPU_DEVICE_ENTRY(inverse_cholesky_mm2_batched_device);

/* Factor only, in place in devAs: lower triangle = L, strictly upper zeroed.
 * src/inverse_cholesky_gpu.cu:356 and :614; devAInvs is unused. */
INVGPU_DEVICE_ENTRY(decompose_cholesky_batched_device);
INVGPU_DEVICE_ENTRY(decompose_cholesky_mm_batched_device);

/* The "stride" family works in place on devAInvs and ignores devAs (src/inverse_cholesky_gpu.cu:95-195):
 * decompose (SPD -> L), inverse_upper (L -> L^-1), multiply_upper (L^-1 -> L^-T L^-1, mirrored);
 * inverse_cholesky_stride = the three in sequence (here fused into one launch). */
INVGPU_DEVICE_ENTRY(decompose_cholesky_stride_batched_device);
INVGPU_DEVICE_ENTRY(inverse_upper_stride_batched_device);
INVGPU_DEVICE_ENTRY(multiply_upper_stride_batched_device);
INVGPU_DEVICE_ENTRY(inverse_cholesky_stride_batched_device);

#undef INVGPU_HOST_ENTRY
#undef INVGPU_DEVICE_ENTRY

#ifdef __cplusplus
}
#endif

#endif /* INVGPU_INVERSE_GPU_H */

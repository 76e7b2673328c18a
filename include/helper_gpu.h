/* helper_gpu.h -- device-side helpers of the drop-in API.
 *
 * Same names and behaviour as reference include/helper_gpu.h:4-69:
 *   batchedCudaMalloc   (src/helper.cu:103-118)  one pitched allocation for a whole batch, the per-matrix
 *                        device pointers written into a caller-provided (host or pinned-host) pointer array;
 *                        this is how every upstream caller of the `*_batched_device` flavour allocates
 *                        (src/gauss_bench.cu:160-167, src/gauss/batched_invert.cu:123-124)
 *   gpuErrchk(expr)      (include/helper_gpu.h:9-18)  "GPUassert: <string> <file>:<line>" on stderr,
 *                        cudaDeviceReset(), exit(code)
 *   cublasErrchk(expr)   (include/helper_gpu.h:60-69) only when cublas_v2.h was included first
 * Must be included after <cuda_runtime.h> and types.h, like upstream.
 */
#ifndef INVGPU_HELPER_GPU_H
#define INVGPU_HELPER_GPU_H

#include <stdio.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif
/* arraySize = BYTES per matrix; *pitch = distance in bytes between consecutive matrices (>= arraySize,
 * 512-byte aligned by cudaMallocPitch); devArrayPtr[i] = base + i * pitch.  Free with cudaFree(devArrayPtr[0]). */
cudaError_t batchedCudaMalloc(Array *devArrayPtr, size_t *pitch, size_t arraySize, int batchSize);
#ifdef __cplusplus
}
#endif

static inline void invgpu_gpu_assert(cudaError_t status, const char *file, int line) {
    if (status == cudaSuccess) return;
    fprintf(stderr, "GPUassert: %s %s:%d\n", cudaGetErrorString(status), file, line);
    cudaDeviceReset();
    exit((int)status);
}
#define gpuErrchk(ans) do { invgpu_gpu_assert((ans), __FILE__, __LINE__); } while (0)

#ifdef CUBLAS_API_H_
static inline const char *invgpu_cublas_status_name(cublasStatus_t s) {
    static const struct { cublasStatus_t code; const char *name; } table[] = {
        {CUBLAS_STATUS_SUCCESS, "CUBLAS_STATUS_SUCCESS"},
        {CUBLAS_STATUS_NOT_INITIALIZED, "CUBLAS_STATUS_NOT_INITIALIZED"},
        {CUBLAS_STATUS_ALLOC_FAILED, "CUBLAS_STATUS_ALLOC_FAILED"},
        {CUBLAS_STATUS_INVALID_VALUE, "CUBLAS_STATUS_INVALID_VALUE"},
        {CUBLAS_STATUS_ARCH_MISMATCH, "CUBLAS_STATUS_ARCH_MISMATCH"},
        {CUBLAS_STATUS_MAPPING_ERROR, "CUBLAS_STATUS_MAPPING_ERROR"},
        {CUBLAS_STATUS_EXECUTION_FAILED, "CUBLAS_STATUS_EXECUTION_FAILED"},
        {CUBLAS_STATUS_INTERNAL_ERROR, "CUBLAS_STATUS_INTERNAL_ERROR"},
        {CUBLAS_STATUS_NOT_SUPPORTED, "CUBLAS_STATUS_NOT_SUPPORTED"},
        {CUBLAS_STATUS_LICENSE_ERROR, "CUBLAS_STATUS_LICENSE_ERROR"},
    };
    for (unsigned i = 0; i < sizeof(table) / sizeof(table[0]); ++i)
        if (table[i].code == s) return table[i].name;
    return "<unknown>";
}
static inline void invgpu_cublas_assert(cublasStatus_t status, const char *file, int line) {
    if (status == CUBLAS_STATUS_SUCCESS) return;
    fprintf(stderr, "cuBLASassert: %s %s:%d\n", invgpu_cublas_status_name(status), file, line);
    cudaDeviceReset();
    exit((int)status);
}
#define cublasErrchk(ans) do { invgpu_cublas_assert((ans), __FILE__, __LINE__); } while (0)
#endif /* CUBLAS_API_H_ */

#endif /* INVGPU_HELPER_GPU_H */

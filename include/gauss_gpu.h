/* gauss_gpu.h -- GPU twins of the reference's GP mean / variance API.
 *
 * The reference exports only the CPU signatures (include/gauss_cpu.h:16-58); its GPU pipeline
 * `calcluateMean` / `calcluateVariance` is file-static inside the benchmark
 * (src/gauss_bench.cu:127-265, 275-409) and takes a leading cublasHandle_t.  These entry
 * points keep the CPU header's names (including its spelling), argument order and meaning,
 * with a GPU suffix, so gauss_bench calls them in place of the static functions:
 *
 *     Means[i]     = A_i^T (B_i + diag(C_i))^-1 D_i
 *     Variances[i] = E_i - A_i^T (B_i + diag(C_i))^-1 A_i        (sign of gauss_cpu.h:34)
 *
 * As, Cs, Ds: batchSize x n;  Bs: batchSize x n x n column-major (upper triangle is read,
 * like cblas_ssymv(Upper)/spotrf_("U") in src/gauss_cpu.c:52-54);  Es, Means, Variances:
 * batchSize scalars (the header comment "batchSize x n x 1" upstream is wrong,
 * src/gauss_cpu.c:66).  All HOST pointers; synchronous.  Unlike the CPU reference
 * (gauss_cpu.h:42) NO input is modified.  One fused kernel: diagonal added on load, Cholesky,
 * both triangular solves and the dot product in-kernel; no inverse ever reaches HBM.
 * A non-SPD (B + diag C) aborts with the reference's message (src/inverse.c:94); the *_ex
 * variants in invgpu.h return per-evaluation info[] instead.
 */
#ifndef INVGPU_GAUSS_GPU_H
#define INVGPU_GAUSS_GPU_H

#include "types.h"

#ifdef __cplusplus
extern "C" {
#endif

void calcluateMeanGPU(int n, Array As, Array Bs, Array Cs, Array Ds, Array Means, int batchSize);
void calcluateVarianceGPU(int n, Array As, Array Bs, Array Cs, Array Es, Array Variances, int batchSize);
/* Same results through the factor + two triangular solves formulation the reference's
 * -DGAUSS_SOLVE build intended (src/gauss_cpu.c:87-144, 221-277; broken upstream). Here both
 * formulations are the same fused kernel. */
void calcluateMeanSolveGPU(int n, Array As, Array Bs, Array Cs, Array Ds, Array Means, int batchSize);
void calcluateVarianceSolveGPU(int n, Array As, Array Bs, Array Cs, Array Es, Array Variances, int batchSize);

#ifdef __cplusplus
}
#endif

#endif /* INVGPU_GAUSS_GPU_H */

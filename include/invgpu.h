/* invgpu.h -- extended C ABI of the B200 batched dense-inversion engine (libinvgpu.so).
 *
 * Everything the legacy-named symbols (inverse_gpu.h, gauss_gpu.h) do is a thin wrapper over
 * these entry points.  Plain pointers and sizes only; no C++ or torch types.  Each entry
 * point cites the reference interface (file:line under /root/reference) it replaces.
 *
 * Conventions
 *   - matrices are column-major, lda = n; "dense" batches are back to back (matrix k at
 *     base + k*n*n), exactly what readMatricesFile produces (src/helper.cu:45).
 *   - `_f32` / `_f64` select the arithmetic type (the reference is fp32 only, types.h:4).
 *   - device-flavour calls are ASYNCHRONOUS on `stream` (a cudaStream_t passed as void*,
 *     NULL = legacy default stream) and never synchronise.
 *   - info (may be NULL): one int per matrix, LAPACK semantics.  SPD paths: spotrf info
 *     (k > 0: leading minor of order k is not positive definite).  General path: sgetrf info
 *     (k > 0: pivot k is exactly zero).  A flagged matrix's output is filled with NaN; the
 *     rest of the batch is processed normally (the reference aborts the process instead,
 *     src/inverse.c:94, src/gauss/inverse_gpu.cu:36).
 *   - return value: 0 on success, a positive cudaError_t, or a negative INVGPU_E* code.
 *   - there is NO CPU fallback anywhere: without a CUDA device every compute call fails.
 */
#ifndef INVGPU_H
#define INVGPU_H

#ifdef __cplusplus
extern "C" {
#endif

#define INVGPU_EARG         (-1)   /* bad argument (n < 1, batch < 0, null pointer) */
#define INVGPU_EUNSUPPORTED (-2)   /* n beyond what the selected path supports */
#define INVGPU_ESINGULAR    (-3)   /* host-flavour call without info[]: some matrix was flagged */

/* largest order every entry point of a family accepts (beyond it: INVGPU_EUNSUPPORTED) */
#define INVGPU_MAX_N_SPD_F32     256
#define INVGPU_MAX_N_SPD_F64     256
#define INVGPU_MAX_N_GENERAL_F32 256
#define INVGPU_MAX_N_GENERAL_F64 256

typedef void *invgpu_stream_t;
typedef long long invgpu_i64;

/* ---- library / device ------------------------------------------------------------------ */
const char *invgpu_version(void);
/* 1 when the library was built with `make lab=1`: the measured-but-not-default kernel generations behind the INVGPU_*
 * environment knobs are present (see csrc/tile_configs.h); 0 in a default build, where those knobs fall back to the defaults. */
int invgpu_has_lab(void);
/* 1 when invgpu_gp_host_* sends only the column prefixes (upper triangle) of the n x n matrices B for this order and element
 * size (columns of at least 256 bytes; no GP tier reads the rest; INVGPU_GP_UPPER_H2D=0 turns it off): what callers that
 * account bytes on the bus need to know (bench.py). */
int invgpu_gp_upper_h2d(int n, int dtype_bytes);
int invgpu_device_count(void);                 /* 0 when no CUDA device is usable */
int invgpu_set_device(int device);             /* cudaSetDevice for the calling host thread (multi-GPU sharding) */
const char *invgpu_error_string(int code);
/* number of kernels launched by this library on the calling thread's device since load
 * (used by bench.py's `gpu_launches`). */
invgpu_i64 invgpu_launch_count(void);
/* name of the kernel tier the dispatcher picks for (op, n, dtype_bytes): "warp", "cta", "generic" */
const char *invgpu_tier_name(int op, int n, int dtype_bytes);

/* ---- SPD inverse, dense strided device batches ------------------------------------------ *
 * A^-1 via Cholesky potrf -> trtri -> lauum; reads the upper triangle, writes both.
 * Replaces inverse_cholesky_batched_device & co (src/inverse_cholesky_gpu.cu:323-354,
 * 607-623, 692-696) and the CPU path inverse_chol_blas (src/inverse.c:89-98). */
int invgpu_spd_inverse_f32(const float *dA, float *dAinv, int n, invgpu_i64 batch, int *dInfo, invgpu_stream_t stream);
int invgpu_spd_inverse_f64(const double *dA, double *dAinv, int n, invgpu_i64 batch, int *dInfo, invgpu_stream_t stream);

/* Cholesky factor only: dL = lower factor, strictly upper zeroed (may alias dA).
 * Replaces decompose_cholesky_batched_device (src/inverse_cholesky_gpu.cu:356-369). */
int invgpu_spd_factor_f32(const float *dA, float *dL, int n, invgpu_i64 batch, int *dInfo, invgpu_stream_t stream);
int invgpu_spd_factor_f64(const double *dA, double *dL, int n, invgpu_i64 batch, int *dInfo, invgpu_stream_t stream);

/* ---- general inverse, dense strided device batches -------------------------------------- *
 * In-place Gauss-Jordan with partial pivoting.  Replaces `invert`
 * (src/gauss/batched_invert.cu:84-95) and the cuBLAS getrf/getriBatched pair
 * (src/gauss/inverse_gpu.cu:24-50). */
int invgpu_general_inverse_f32(const float *dA, float *dAinv, int n, invgpu_i64 batch, int *dInfo, invgpu_stream_t stream);
int invgpu_general_inverse_f64(const double *dA, double *dAinv, int n, invgpu_i64 batch, int *dInfo, invgpu_stream_t stream);

/* ---- LU factors with pivots / inverse from factors / multi-RHS solve ------------------------ *
 * The cublasSgetrfBatched / cublasSgetriBatched pair of the reference's fastest GPU path
 * (src/gauss/inverse_gpu.cu:24-33 getrf in place with PivotArray / infoArray :21-22, :39-50 getri out of place)
 * for callers that want the FACTORS, and the solve its CPU side formulates with spotrs_ (src/gauss_cpu.c:87-144),
 * generalised to non-symmetric systems.  LAPACK semantics: P A = L U in place (unit lower L, U), dPivots[k*n + i] =
 * 1-based row that row i was interchanged with, dInfo[k] = first i with U(i,i) exactly zero (the factorisation is
 * completed anyway; getri / gesv then give NaN for that matrix).  dPivots / dInfo are DEVICE arrays, may be NULL
 * in getrf / gesv.  n <= 256, nrhs <= 256.  B is n x nrhs column-major (ldb = n), batch dense; gesv leaves LU in dA
 * and X in dB.  `_ptrs`: cuBLAS-style arrays of per-matrix device pointers (lda = n). */
int invgpu_getrf_f32(float *dA, int n, int *dPivots, int *dInfo, invgpu_i64 batch, invgpu_stream_t stream);
int invgpu_getrf_f64(double *dA, int n, int *dPivots, int *dInfo, invgpu_i64 batch, invgpu_stream_t stream);
int invgpu_getrf_ptrs_f32(float *const *As, int n, int *dPivots, int *dInfo, int batch, invgpu_stream_t stream);
int invgpu_getrf_ptrs_f64(double *const *As, int n, int *dPivots, int *dInfo, int batch, invgpu_stream_t stream);
int invgpu_getri_f32(const float *dLU, const int *dPivots, float *dAinv, int n, int *dInfo, invgpu_i64 batch, invgpu_stream_t stream);
int invgpu_getri_f64(const double *dLU, const int *dPivots, double *dAinv, int n, int *dInfo, invgpu_i64 batch, invgpu_stream_t stream);
int invgpu_getri_ptrs_f32(float *const *LUs, const int *dPivots, float *const *Ainvs, int n, int *dInfo, int batch, invgpu_stream_t stream);
int invgpu_getri_ptrs_f64(double *const *LUs, const int *dPivots, double *const *Ainvs, int n, int *dInfo, int batch, invgpu_stream_t stream);
int invgpu_gesv_f32(float *dA, int *dPivots, float *dB, int n, int nrhs, int *dInfo, invgpu_i64 batch, invgpu_stream_t stream);
int invgpu_gesv_f64(double *dA, int *dPivots, double *dB, int n, int nrhs, int *dInfo, invgpu_i64 batch, invgpu_stream_t stream);

/* ---- pointer-array flavour (reference `Array *devAs`) ------------------------------------ *
 * As[k] / Ainvs[k] are device pointers; the arrays may be in pinned-host or device memory.
 * stages: bit mask 1 = potrf, 2 = trtri, 4 = lauum (7 = full inverse, 1 = factor);
 * As == Ainvs (in place) is allowed. */
int invgpu_spd_stages_ptrs_f32(float *const *As, float *const *Outs, int n, int batch, int stages, int *dInfo, invgpu_stream_t stream);
int invgpu_spd_stages_ptrs_f64(double *const *As, double *const *Outs, int n, int batch, int stages, int *dInfo, invgpu_stream_t stream);
int invgpu_general_inverse_ptrs_f32(float *const *As, float *const *Ainvs, int n, int batch, int *dInfo, invgpu_stream_t stream);
int invgpu_general_inverse_ptrs_f64(double *const *As, double *const *Ainvs, int n, int batch, int *dInfo, invgpu_stream_t stream);

/* ---- mixed dimensions: persistent-CTA scheduler ------------------------------------------- *
 * SPD inverse of `count` matrices of individual order ns[i] (1..256; fp64: as far as a packed
 * triangle fits one CTA's shared memory, n <= 236).  As / Ainvs / ns are HOST arrays; As[i] / Ainvs[i]
 * are DEVICE pointers (column-major, lda = ns[i]).  The work is bucketed by size (<= 32, <= 128,
 * <= 256 -- the reference README's "size-bucketed stream queues", README.md:41-44) and every bucket is
 * drained by a persistent grid pulling work with an atomic ticket, longest matrices first; the three
 * bucket kernels run concurrently.  Asynchronous with respect to `stream` once the work list is
 * uploaded; dInfo[i] (device, may be NULL) follows the caller's order. */
int invgpu_mixed_spd_inverse_f32(float *const *As, float *const *Ainvs, const int *ns, invgpu_i64 count, int *dInfo, invgpu_stream_t stream);
int invgpu_mixed_spd_inverse_f64(double *const *As, double *const *Ainvs, const int *ns, invgpu_i64 count, int *dInfo, invgpu_stream_t stream);

/* ---- fused GP mean / variance, dense device batches -------------------------------------- *
 * means[i] = A_i^T (B_i + diag C_i)^-1 D_i ; variances[i] = E_i - A_i^T (B_i + diag C_i)^-1 A_i.
 * ONE kernel replaces addDiagonal -> batchedInverse -> batchedMul -> batchedMul
 * (src/gauss_bench.cu:38,68,87,127-265,275-409).  dD/dMeans or dE/dVariances may be NULL to
 * compute only the other quantity; with both present the factorisation is shared. */
int invgpu_gp_f32(int n, const float *dA, const float *dB, const float *dC, const float *dD, const float *dE,
                  float *dMeans, float *dVariances, invgpu_i64 batch, int *dInfo, invgpu_stream_t stream);
int invgpu_gp_f64(int n, const double *dA, const double *dB, const double *dC, const double *dD, const double *dE,
                  double *dMeans, double *dVariances, invgpu_i64 batch, int *dInfo, invgpu_stream_t stream);

/* ---- host flavour with per-matrix info --------------------------------------------------- *
 * Same contract as the reference's *_gpu wrappers (host pointers, synchronous, H2D + compute
 * + D2H inside the call; e.g. src/gauss/batched_invert.cu:99-177) but chunked and pipelined
 * over a persistent pinned ring instead of per-call cudaHostAlloc/cudaMallocPitch/cudaFree.
 * info may be NULL (then a flagged matrix makes the call return INVGPU_ESINGULAR). */
int invgpu_spd_inverse_host_f32(const float *As, float *aInvs, int n, invgpu_i64 batch, int *info);
int invgpu_spd_inverse_host_f64(const double *As, double *aInvs, int n, invgpu_i64 batch, int *info);
int invgpu_general_inverse_host_f32(const float *As, float *aInvs, int n, invgpu_i64 batch, int *info);
int invgpu_general_inverse_host_f64(const double *As, double *aInvs, int n, invgpu_i64 batch, int *info);
int invgpu_gp_host_f32(int n, const float *As, const float *Bs, const float *Cs, const float *Ds, const float *Es,
                       float *Means, float *Variances, invgpu_i64 batch, int *info);
int invgpu_gp_host_f64(int n, const double *As, const double *Bs, const double *Cs, const double *Ds, const double *Es,
                       double *Means, double *Variances, invgpu_i64 batch, int *info);

/* Transfer probe: the host pipeline above with the kernel replaced by a device-to-device copy -- `batch` units of
 * `unit_bytes` go host -> device -> host through the same chunks, streams and ring.  What it measures is the
 * ceiling of every host-flavour call on this box (bench.py `e2e_ceiling`, tools/xfer_bench); it is the engine's
 * counterpart of the reference's transfer micro-benchmarks (src/bench.cu:26-158). */
int invgpu_xfer_roundtrip_host(const void *in, void *out, unsigned long long unit_bytes, invgpu_i64 batch);
/* NUMA node of a CUDA device (sysfs), -1 when the platform does not say. */
int invgpu_device_numa_node(int device);

/* pinned host allocations for callers that want the zero-staging fast path of the host flavour */
void *invgpu_host_alloc(unsigned long long bytes);
void invgpu_host_free(void *p);
/* drop the cached device workspace / pinned ring of the calling thread's device */
void invgpu_release_workspace(void);

#ifdef __cplusplus
}
#endif

#endif /* INVGPU_H */

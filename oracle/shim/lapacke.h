/* Fortran-ABI LAPACK prototypes the reference calls directly although it
 * includes <lapacke.h> (src/inverse.c:63-66,92-97; src/gauss_cpu.c:118,122).
 * Resolved by the image's OpenBLAS 0.3.15 (LP64, un-prefixed symbols).
 * Test infrastructure only. */
#ifndef ORACLE_SHIM_LAPACKE_H
#define ORACLE_SHIM_LAPACKE_H
#ifdef __cplusplus
extern "C" {
#endif
void spotrf_(const char *uplo, const int *n, float *a, const int *lda, int *info);
void spotri_(const char *uplo, const int *n, float *a, const int *lda, int *info);
void spotrs_(const char *uplo, const int *n, const int *nrhs, const float *a, const int *lda,
             float *b, const int *ldb, int *info);
void sgetrf_(const int *m, const int *n, float *a, const int *lda, int *ipiv, int *info);
void sgetri_(const int *n, float *a, const int *lda, const int *ipiv, float *work,
             const int *lwork, int *info);
#ifdef __cplusplus
}
#endif
#endif

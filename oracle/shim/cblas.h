/* Minimal CBLAS declarations so the reference's CPU sources compile against
 * the OpenBLAS 0.3.15 bundled in this image (no system cblas.h exists).
 * Only what /root/reference/src/{gauss_cpu.c,inverse.c,inverse_bench.c,gauss_bench.cu}
 * call.  Test infrastructure only. */
#ifndef ORACLE_SHIM_CBLAS_H
#define ORACLE_SHIM_CBLAS_H
#ifdef __cplusplus
extern "C" {
#endif
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
enum CBLAS_UPLO { CblasUpper = 121, CblasLower = 122 };
float cblas_sdot(int n, const float *x, int incx, const float *y, int incy);
float cblas_sasum(int n, const float *x, int incx);
void cblas_saxpy(int n, float alpha, const float *x, int incx, float *y, int incy);
void cblas_scopy(int n, const float *x, int incx, float *y, int incy);
void cblas_sscal(int n, float alpha, float *x, int incx);
void cblas_ssymv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, float alpha,
                 const float *a, int lda, const float *x, int incx, float beta, float *y, int incy);
void cblas_ssyrk(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans,
                 int n, int k, float alpha, const float *a, int lda, float beta, float *c, int ldc);
#ifdef __cplusplus
}
#endif
#endif

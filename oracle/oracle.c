/*
 * oracle.c -- CPU ORACLE.  THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the arithmetic on the reference's hot path
 * (batched dense inversion and the GP mean/variance), used ONLY as the checker
 * in tests/, in __graft_entry__.smoke() and as the timed `cpu_baseline` /
 * `--impl reference` leg of bench.py.  Nothing under cuda_matrix_inversion_b200/
 * links, loads or calls it; the product path has no CPU fallback.
 *
 * Where the algorithm lives: the reference's CPU path is a thin wrapper over
 * LAPACK/CBLAS (`spotrf_/spotri_/sgetrf_/sgetri_/spotrs_`, `cblas_ssymv/sdot`;
 * reference src/inverse.c:63-66,92-97, src/gauss_cpu.c:54-72) -- a third-party
 * dependency that is NOT in /root/reference and is unpinned there (the
 * Makefile just links -llapack -lblas, Makefile:135).  The restatement below
 * therefore follows (a) the reference's own dependency-free statement of the
 * SPD algorithm, src/inverse_cholesky_cpu.c:17-85, (b) its Gauss-Jordan kernel
 * sequence, src/gauss/batched_invert.cu:17-95, (c) the call sequence of
 * src/gauss_cpu.c:46-72, and (d) LAPACK's published unblocked algorithms for
 * the partial-pivot LU inverse.
 *
 * Parity pinning: tests/test_oracle.py checks this file against every golden
 * vector the reference ships for the path (tests/golden/reference/:
 * inverse_100_{8,16,32}/aInv.mats, gaussian_100_NxN/{means,variances}.mats,
 * simpleMean/chol.mats -> cholinv.mats) and against the reference's own CPU
 * sources compiled unmodified here against OpenBLAS 0.3.15 (oracle/_ref, built
 * by oracle/Makefile).  fp64 has no reference counterpart (the reference is
 * fp32-only, include/types.h:4): the fp64 instantiation is pinned against
 * numpy/LAPACK fp64 in the same test file -- "fp64 parity unpinned by the
 * reference" (DESIGN.md).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL float
#define NAME(x) CAT(x, _f32)
#define SQRT sqrtf
#define FABS fabsf
#include "oracle_impl.inc"
#undef REAL
#undef NAME
#undef SQRT
#undef FABS

#define REAL double
#define NAME(x) CAT(x, _f64)
#define SQRT sqrt
#define FABS fabs
#include "oracle_impl.inc"
#undef REAL
#undef NAME
#undef SQRT
#undef FABS

/*
 * `.mats` reader, restating `readMatricesFile` (reference src/helper.cu:15-52):
 * header "numMatrices m n", then numMatrices*m*n numbers row by row; each
 * matrix is stored column-major, element (i,j) at [j*m + i] (:45).
 * Two-call protocol: dst == NULL -> only the header is parsed.
 * Returns 0 on success, -1 open failure, -2 bad header, -3 truncated data.
 */
int orc_read_mats(const char *path, int *num, int *m, int *n, double *dst)
{
    FILE *fp = fopen(path, "r");
    if (!fp) return -1;
    int k_, m_, n_;
    if (fscanf(fp, "%d %d %d", &k_, &m_, &n_) != 3) { fclose(fp); return -2; }
    *num = k_; *m = m_; *n = n_;
    if (dst) {
        for (int k = 0; k < k_; ++k) {
            double *cur = dst + (size_t)k * m_ * n_;
            for (int i = 0; i < m_; ++i)
                for (int j = 0; j < n_; ++j)
                    if (fscanf(fp, "%lf", &cur[(size_t)j * m_ + i]) != 1) { fclose(fp); return -3; }
        }
    }
    fclose(fp);
    return 0;
}

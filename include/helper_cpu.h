/* helper_cpu.h -- `.mats` file I/O and printing helpers of the drop-in API.
 *
 * Same four symbols as reference include/helper_cpu.h:33-36 (defined in src/helper.cu:15-101),
 * plus a writer.  File format: SURVEY.md Appendix B -- "numMatrices m n" then the matrices
 * row by row as text; in memory every matrix is column-major, element (i,j) at [j*m + i].
 * Failures print a message and exit(EXIT_FAILURE) like the reference's `ensure`
 * (include/helper_cpu.h:12-21).
 */
#ifndef INVGPU_HELPER_CPU_H
#define INVGPU_HELPER_CPU_H

#include "types.h"

/* reference include/helper_cpu.h:4 -- size guard of readMatricesFile */
#define MAX_MATRIX_BYTE_READ 67108864

/* reference include/helper_cpu.h:6-23 -- the error convention of the whole code base: message on stderr, then
 * exit(EXIT_FAILURE).  `ensure` also reports errno when it is set; callers need <stdio.h>, <stdlib.h>, <errno.h>. */
#ifndef fail
#define fail(...)                                           \
    do {                                                    \
        fprintf(stderr, "%s:%d\t", __FILE__, __LINE__);     \
        fprintf(stderr, __VA_ARGS__);                       \
        fprintf(stderr, "\r\n");                            \
        exit(EXIT_FAILURE);                                 \
    } while (0)
#endif
#ifndef ensure
#define ensure(condition, ...)                                                        \
    do {                                                                              \
        if (!(condition)) {                                                           \
            fprintf(stderr, "ENSURE FAILED %s:%d\r\n", __FILE__, __LINE__);           \
            fprintf(stderr, __VA_ARGS__);                                             \
            fprintf(stderr, "\r\n");                                                  \
            if (errno) perror("possible reason for failure from ERRNO");              \
            exit(EXIT_FAILURE);                                                       \
        }                                                                             \
    } while (0)
#endif
#ifndef div_ceil
#define div_ceil(x, y) (1 + (((x) - 1) / (y)))   /* positive x, y */
#endif

#ifdef __cplusplus
extern "C" {
#endif

void readMatricesFile(const char *path, int *numMatrices, int *m, int *n, Array *matrices);
void replicateMatrices(Array *matrices, const int M, const int N, const int numMatrices, const int numReplications);
void printMatrix(Array a, int M, int N);
void printMatrixList(Array a, int N, int batchSize);

/* new: inverse of readMatricesFile (digits = significant digits, 0 -> 9 which round-trips fp32) */
int writeMatricesFile(const char *path, int numMatrices, int m, int n, const DataType *matrices, int digits);

#ifdef __cplusplus
}
#endif

#endif /* INVGPU_HELPER_CPU_H */

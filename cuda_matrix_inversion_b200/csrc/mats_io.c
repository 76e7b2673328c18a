/* mats_io.c -- `.mats` text I/O and the small host helpers of the drop-in API
 * (include/helper_cpu.h).  Behaviour follows reference src/helper.cu:15-101:
 * header "numMatrices m n", matrices row by row in the file, column-major in memory
 * (element (i,j) at [j*m + i]), one malloc'd block for the whole list, a 64 MiB size guard,
 * and failures that print and exit(EXIT_FAILURE) (include/helper_cpu.h:12-21).
 * Unlike the reference a short file is an error (its `ensure(ret, ...)` accepts EOF,
 * SURVEY.md App. A-9). */
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/helper_cpu.h"

#define MATS_FAIL(...)                                              \
    do {                                                            \
        fprintf(stderr, "ENSURE FAILED %s:%d\r\n", __FILE__, __LINE__); \
        fprintf(stderr, __VA_ARGS__);                               \
        fprintf(stderr, "\r\n");                                    \
        if (errno) perror("possible reason for failure from ERRNO"); \
        exit(EXIT_FAILURE);                                         \
    } while (0)

void readMatricesFile(const char *path, int *numMatrices, int *m, int *n, Array *matrices)
{
    FILE *fp = fopen(path, "r");
    if (!fp) MATS_FAIL("could not open matrix file %s", path);

    int count = 0, rows = 0, cols = 0;
    if (fscanf(fp, "%d %d %d", &count, &rows, &cols) != 3 || count < 0 || rows < 0 || cols < 0)
        MATS_FAIL("could not read number of matrices from file %s", path);

    const size_t per = (size_t)rows * (size_t)cols;
    const size_t bytes = sizeof(DataType) * per * (size_t)count;
    if (bytes > MAX_MATRIX_BYTE_READ)
        MATS_FAIL("cannot read file %s because the allocated array would be bigger than 0x%lX bytes",
                  path, (unsigned long)bytes);

    DataType *block = (DataType *)malloc(bytes ? bytes : 1);
    if (!block) MATS_FAIL("could not allocate 0x%lX bytes of memory for file %s", (unsigned long)bytes, path);

    for (int k = 0; k < count; ++k) {
        DataType *mat = block + (size_t)k * per;
        for (int i = 0; i < rows; ++i) {            /* the file is row-major ... */
            for (int j = 0; j < cols; ++j) {
                double v;
                if (fscanf(fp, "%lf", &v) != 1)
                    MATS_FAIL("could not read matrix from file %s, stuck at matrix %d element %d, %d", path, k, i, j);
                mat[(size_t)j * rows + i] = (DataType)v;   /* ... memory is column-major */
            }
        }
    }
    fclose(fp);
    *numMatrices = count; *m = rows; *n = cols; *matrices = block;
}

void replicateMatrices(Array *matrices, const int M, const int N, const int numMatrices, const int numReplications)
{
    const size_t list = sizeof(DataType) * (size_t)M * (size_t)N * (size_t)numMatrices;
    char *tiled = (char *)malloc(list * (size_t)(numReplications > 0 ? numReplications : 1) + 1);
    if (!tiled) MATS_FAIL("Could not allocate memory for the replicated array (%lu bytes).",
                          (unsigned long)(list * (size_t)numReplications));
    for (int r = 0; r < numReplications; ++r) memcpy(tiled + (size_t)r * list, *matrices, list);
    free(*matrices);
    *matrices = (Array)tiled;
}

void printMatrix(Array a, int M, int N)
{
    for (int i = 0; i < M; ++i) {
        for (int j = 0; j < N; ++j) printf("%f\t", a[(size_t)j * M + i]);
        printf("\n");
    }
    printf("\n");
}

void printMatrixList(Array a, int N, int batchSize)
{
    for (int k = 0; k < batchSize; ++k) {
        printf("=============== <%d> ===============\n", k + 1);
        printMatrix(a + (size_t)k * N * N, N, N);
    }
}

int writeMatricesFile(const char *path, int numMatrices, int m, int n, const DataType *matrices, int digits)
{
    FILE *fp = fopen(path, "w");
    if (!fp) return -1;
    if (digits <= 0) digits = 9;
    fprintf(fp, "%d %d %d\n", numMatrices, m, n);
    for (int k = 0; k < numMatrices; ++k) {
        const DataType *mat = matrices + (size_t)k * m * n;
        for (int i = 0; i < m; ++i) {
            for (int j = 0; j < n; ++j)
                fprintf(fp, j + 1 < n ? "%.*g\t" : "%.*g", digits, (double)mat[(size_t)j * m + i]);
            fprintf(fp, "\n");
        }
    }
    return fclose(fp) ? -1 : 0;
}

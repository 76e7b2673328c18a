/* bench_common.h -- shared pieces of the reference-compatible CLIs (inverse_bench, gauss_bench):
 * monotonic-clock timers with Welford running mean/variance (what reference include/timer.h:99-132
 * accumulates), the report lines of SURVEY.md Appendix C, and the optional hooks described below.
 *
 * Optional trailing arguments (the reference invocations `DIR REPS DUPS [-csv]` keep working):
 *   --gpus N        shard the batch contiguously over N devices, one host thread per device
 *                   (no collective: every thread writes its own slice of the host output)
 *   --cpu-lib PATH  dlopen a build of the reference's CPU path (its own symbols
 *                   inverse_lu_blas / inverse_lu_blas_omp / calcluateMeanCPU / calcluateVarianceCPU)
 *                   and time it for the *_cpu rows.  Without it those rows are omitted: this
 *                   program has no CPU implementation of its own.
 *   --json          one extra JSON line with throughput and the roofline fraction
 *   --pageable      keep the GPU calls' host buffers in pageable malloc memory (default: pinned, see bench_buffer)
 *   --dump PATH     write the raw output array of the last algorithm (tests compare runs bit for bit)
 */
#ifndef INVGPU_BENCH_COMMON_H
#define INVGPU_BENCH_COMMON_H

#include <math.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
    struct timespec t0;
    double last_ms, total_ms, mean, m2;
    long n;
} bench_timer;

static inline void bt_start(bench_timer *t) { clock_gettime(CLOCK_MONOTONIC, &t->t0); }
static inline void bt_stop(bench_timer *t) {
    struct timespec t1;
    clock_gettime(CLOCK_MONOTONIC, &t1);
    t->last_ms = (t1.tv_sec - t->t0.tv_sec) * 1e3 + (t1.tv_nsec - t->t0.tv_nsec) * 1e-6;
    t->n += 1;
    t->total_ms += t->last_ms;
    const double delta = t->last_ms - t->mean;
    t->mean += delta / (double)t->n;
    t->m2 += delta * (t->last_ms - t->mean);
}
static inline double bt_variance(const bench_timer *t) { return t->n > 1 ? t->m2 / (double)(t->n - 1) : 0.0; }

/* Appendix C line formats (reference src/inverse_bench.c:54-71, src/gauss_bench.cu:504-529) */
static inline void bench_report(const char *name, int numMatrices, int n, int numReps, const bench_timer *t,
                                double avg_err, bool csv)
{
    if (csv) {
        if (numReps > 1) printf("%d %d %d %s %e %e %e %e\n", numMatrices, n, numReps, name, t->total_ms, t->mean, bt_variance(t), avg_err);
        else printf("%d %d %d %s %e %e\n", numMatrices, n, numReps, name, t->total_ms, avg_err);
    } else {
        if (numReps > 1)
            printf("%s - %d %dx%d matrices, replicated %d times, runtime %.4f ms (%.4f ms average, %.4f ms variance), average error %.4e\n",
                   name, numMatrices, n, n, numReps, t->total_ms, t->mean, bt_variance(t), avg_err);
        else
            printf("%s - %d %dx%d matrices, replicated %d times, runtime %.4f ms, average error %.4e\n",
                   name, numMatrices, n, n, numReps, t->total_ms, avg_err);
    }
}

#define BENCH_ENSURE(cond, ...)                                              \
    do {                                                                     \
        if (!(cond)) {                                                       \
            fprintf(stderr, "ENSURE FAILED %s:%d\r\n", __FILE__, __LINE__);  \
            fprintf(stderr, __VA_ARGS__);                                    \
            fprintf(stderr, "\r\n");                                         \
            exit(EXIT_FAILURE);                                              \
        }                                                                    \
    } while (0)

/* sum |x - y| over count elements (the reference's cblas_saxpy + cblas_sasum pair) */
static inline double l1_distance(const float *x, const float *y, size_t count)
{
    double s = 0;
    for (size_t i = 0; i < count; ++i) s += fabs((double)x[i] - (double)y[i]);
    return s;
}

/* Buffers handed to the *_gpu entry points: pinned host memory from the engine's allocator (DMA'd directly by the host
 * pipeline, no staging copy; with --gpus N that is what lets the devices' transfers run side by side), pageable malloc
 * memory when the allocator is unavailable or --pageable is given (what the reference's CLIs pass). */
void *invgpu_host_alloc(unsigned long long bytes);
void invgpu_host_free(void *p);
static inline void *bench_buffer(size_t bytes, bool pageable, bool *pinned)
{
    void *p = pageable ? NULL : invgpu_host_alloc(bytes);
    *pinned = p != NULL;
    return p ? p : malloc(bytes);
}
static inline void bench_buffer_free(void *p, bool pinned) { if (pinned) invgpu_host_free(p); else free(p); }

typedef struct {
    bool csv, json, pageable;
    int gpus;
    const char *cpu_lib;
    const char *dump;
} bench_opts;

static inline bench_opts parse_opts(int argc, char const *argv[])
{
    bench_opts o = {false, false, false, 1, NULL, NULL};
    for (int i = 4; i < argc; ++i) {
        if (!strncmp("-csv", argv[i], 4)) o.csv = true;
        else if (!strcmp("--json", argv[i])) o.json = true;
        else if (!strcmp("--pageable", argv[i])) o.pageable = true;
        else if (!strcmp("--gpus", argv[i]) && i + 1 < argc) o.gpus = atoi(argv[++i]);
        else if (!strcmp("--cpu-lib", argv[i]) && i + 1 < argc) o.cpu_lib = argv[++i];
        else if (!strcmp("--dump", argv[i]) && i + 1 < argc) o.dump = argv[++i];
    }
    if (o.gpus < 1) o.gpus = 1;
    return o;
}

#endif

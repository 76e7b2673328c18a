/* inverse_bench -- drop-in for the reference's `inverse_bench TEST_FOLDER TEST_REPLICATIONS
 * MATRIX_DUPLICATES [-csv]` (reference src/inverse_bench.c:276-303): reads DIR/a.mats and
 * DIR/aInv.mats, tiles them DUPS times, runs every inverse REPS times and prints one line per
 * algorithm in the reference's order and format:
 *     [lu_blas_cpu, lu_blas_omp_cpu,]  chol_gpu, chol_mm2_gpu, gauss_batched_gpu, lu_cuda_batched_gpu
 * The GPU rows call the same symbols as upstream (include/inverse_gpu.h), now served by the B200
 * engine; timing wraps the whole host call (H2D + compute + D2H), as upstream did (report.tex:104).
 * avg_err = sum |inv - aInv| / numMatrices after the last repetition (inverse_bench.c:49-51).
 * Plain C host code over the C ABI; see bench_common.h for the optional trailing flags.
 */
#include <dlfcn.h>
#include <omp.h>

#include "../../include/types.h"
#include "../../include/helper_cpu.h"
#include "../../include/inverse_gpu.h"
#include "../../include/invgpu.h"
#include "bench_common.h"

typedef void (*host_inverse_fn)(cublasHandle_t, int, Array, Array, int);

static void read_test(const char *dir, int *numMatrices, int *n, Array *a, Array *aInv)
{
    char path[1024];
    int kA, mA, nA, kI, mI, nI;
    snprintf(path, sizeof path, "%s/a.mats", dir);
    readMatricesFile(path, &kA, &mA, &nA, a);
    snprintf(path, sizeof path, "%s/aInv.mats", dir);
    readMatricesFile(path, &kI, &mI, &nI, aInv);
    BENCH_ENSURE(kA == kI, "test in directory %s invalid, number of matrices in files not matching\r\n"
                           "numMatricesA(%d) numMatricesAInv(%d)", dir, kA, kI);
    BENCH_ENSURE(mA == mI && nA == nI && mA == nA, "test in directory %s invalid, dimensions not matching\r\n"
                 "mA(%d) mAInv(%d)\r\nnA(%d) nAInv(%d)", dir, mA, mI, nA, nI);
    *numMatrices = kA;
    *n = mA;
}

/* one algorithm: REPS timed calls of fn on a fresh copy of the input, sharded over `gpus` devices */
static double run_gpu(host_inverse_fn fn, int gpus, int n, int numMatrices, int numReps, const float *a,
                      float *work, float *inv, const float *aInv, bench_timer *t)
{
    const size_t per = (size_t)n * n;
    /* rep -1 is an untimed warm-up: CUDA context creation, kernel loading and the first allocation of the
       host pipeline are not what the reference times either (it creates its cuBLAS handle before the loops,
       src/inverse_bench.c:282-284) */
    for (int rep = -1; rep < numReps; ++rep) {
        memcpy(work, a, per * numMatrices * sizeof(float));
        if (rep >= 0) bt_start(t);
        if (gpus == 1) {
            fn(NULL, n, work, inv, numMatrices);
        } else {
            #pragma omp parallel num_threads(gpus)
            {
                const int g = omp_get_thread_num();
                const long lo = (long)numMatrices * g / gpus, hi = (long)numMatrices * (g + 1) / gpus;
                invgpu_set_device(g);
                if (hi > lo) fn(NULL, n, work + lo * per, inv + lo * per, (int)(hi - lo));
            }
        }
        if (rep >= 0) bt_stop(t);
    }
    return l1_distance(inv, aInv, per * numMatrices) / numMatrices;
}

int main(int argc, char const *argv[])
{
    BENCH_ENSURE(argc >= 4, "Usage: inverse_bench TEST_FOLDER TEST_REPLICATIONS MATRIX_DUPLICATES [-csv] "
                            "[--gpus N] [--cpu-lib PATH] [--json]");
    const int numReps = atoi(argv[2]), numDuplicates = atoi(argv[3]);
    bench_opts opt = parse_opts(argc, argv);
    int numMatrices, n;
    Array a, aInv;
    read_test(argv[1], &numMatrices, &n, &a, &aInv);
    replicateMatrices(&a, n, n, numMatrices, numDuplicates);
    replicateMatrices(&aInv, n, n, numMatrices, numDuplicates);
    numMatrices *= numDuplicates;
    const size_t total = (size_t)numMatrices * n * n;
    bool pin_inv, pin_work;
    float *inv = (float *)bench_buffer(total * sizeof(float), opt.pageable, &pin_inv);
    float *work = (float *)bench_buffer(total * sizeof(float), opt.pageable, &pin_work);
    BENCH_ENSURE(inv && work, "Could not allocate the result buffers");

    /* optional CPU rows, from a user-supplied build of the reference CPU path */
    if (opt.cpu_lib) {
        void *h = dlopen(opt.cpu_lib, RTLD_NOW | RTLD_GLOBAL);
        BENCH_ENSURE(h, "could not load --cpu-lib %s: %s", opt.cpu_lib, dlerror());
        void (*lu1)(Array, Array, int) = (void (*)(Array, Array, int))dlsym(h, "inverse_lu_blas");
        void (*luomp)(Array, int, int) = (void (*)(Array, int, int))dlsym(h, "inverse_lu_blas_omp");
        BENCH_ENSURE(lu1 && luomp, "%s does not export inverse_lu_blas / inverse_lu_blas_omp", opt.cpu_lib);
        bench_timer t1 = {0}, t2 = {0};
        for (int rep = 0; rep < numReps; ++rep) {
            memcpy(inv, a, total * sizeof(float));
            bt_start(&t1);
            for (int i = 0; i < numMatrices; ++i) lu1(inv + (size_t)i * n * n, work, n);
            bt_stop(&t1);
        }
        bench_report("lu_blas_cpu", numMatrices, n, numReps, &t1, l1_distance(inv, aInv, total) / numMatrices, opt.csv);
        for (int rep = 0; rep < numReps; ++rep) {
            memcpy(inv, a, total * sizeof(float));
            bt_start(&t2);
            luomp(inv, n, numMatrices);
            bt_stop(&t2);
        }
        bench_report("lu_blas_omp_cpu", numMatrices, n, numReps, &t2, l1_distance(inv, aInv, total) / numMatrices, opt.csv);
    }

    BENCH_ENSURE(invgpu_device_count() >= opt.gpus, "%d CUDA device(s) requested, %d usable: this program has no CPU path",
                 opt.gpus, invgpu_device_count());
    static const struct { const char *name; host_inverse_fn fn; } algos[] = {
        {"chol_gpu", inverse_cholesky_batched_gpu},
        {"chol_mm2_gpu", inverse_cholesky_mm2_batched_gpu},
        {"gauss_batched_gpu", inverse_gauss_batched_gpu},
        {"lu_cuda_batched_gpu", inverse_lu_cuda_batched_gpu},
    };
    double best_ms = 1e300;
    for (size_t k = 0; k < sizeof algos / sizeof algos[0]; ++k) {
        bench_timer t = {0};
        const double err = run_gpu(algos[k].fn, opt.gpus, n, numMatrices, numReps, a, work, inv, aInv, &t);
        bench_report(algos[k].name, numMatrices, n, numReps, &t, err, opt.csv);
        if (t.mean < best_ms) best_ms = t.mean;
    }
    if (opt.json)
        printf("{\"bench\": \"inverse_bench\", \"n\": %d, \"numMatrices\": %d, \"gpus\": %d, \"best_ms\": %.6f, "
               "\"inversions_per_s\": %.6e, \"end_to_end\": true}\n",
               n, numMatrices, opt.gpus, best_ms, numMatrices / (best_ms * 1e-3));
    if (opt.dump) {
        FILE *f = fopen(opt.dump, "wb");
        BENCH_ENSURE(f && fwrite(inv, sizeof(float), total, f) == total, "could not write %s", opt.dump);
        fclose(f);
    }
    bench_buffer_free(work, pin_work); bench_buffer_free(inv, pin_inv); free(a); free(aInv);
    return 0;
}

/* gauss_bench -- drop-in for the reference's `gauss_bench TEST_FOLDER TEST_REPLICATIONS
 * MATRIX_DUPLICATES [-csv]` (reference src/gauss_bench.cu:577-702): reads DIR/{a,b,c,d,e,means,
 * variances}.mats, checks the shapes like upstream (:441-463), tiles them DUPS times and times the
 * GP mean and variance REPS times.  Output: two lines per side in the reference's format
 * (SURVEY.md Appendix C): [means_cpu, variances_cpu,] means_gpu, variances_gpu, with
 * avg_err = sum_reps sum |out - golden| / numMatrices / numReps (:493-529).
 * The GPU rows call calcluateMeanGPU / calcluateVarianceGPU (include/gauss_gpu.h): ONE fused kernel
 * instead of upstream's addDiagonal -> getrf/getriBatched -> gemv -> dot pipeline; the timing wraps
 * the whole host call, transfers included, as upstream did.
 */
#include <dlfcn.h>
#include <omp.h>

#include "../../include/types.h"
#include "../../include/helper_cpu.h"
#include "../../include/gauss_gpu.h"
#include "../../include/invgpu.h"
#include "bench_common.h"

typedef void (*gp_fn)(int, Array, Array, Array, Array, Array, int);

static void read_test(const char *dir, int *numMatrices, int *n, Array *a, Array *b, Array *c, Array *d, Array *e,
                      Array *means, Array *variances)
{
    static const char *names[7] = {"a", "b", "c", "d", "e", "means", "variances"};
    Array *dst[7] = {a, b, c, d, e, means, variances};
    int num[7], m[7], nn[7];
    char path[1024];
    for (int i = 0; i < 7; ++i) {
        snprintf(path, sizeof path, "%s/%s.mats", dir, names[i]);
        readMatricesFile(path, &num[i], &m[i], &nn[i], dst[i]);
    }
    for (int i = 1; i < 7; ++i)
        BENCH_ENSURE(num[i] == num[0], "test in directory %s invalid, number of matrices in files not matching\r\n"
                     "%s.mats(%d) a.mats(%d)", dir, names[i], num[i], num[0]);
    BENCH_ENSURE(m[0] == m[1] && m[1] == m[2] && m[2] == m[3] && m[4] == 1 && m[5] == 1 && m[6] == 1 &&
                 nn[0] == 1 && nn[1] == m[1] && nn[2] == 1 && nn[3] == 1 && nn[4] == 1 && nn[5] == 1 && nn[6] == 1,
                 "test in directory %s invalid, dimensions not matching\r\n"
                 "mA(%d) mB(%d) mC(%d) mD(%d)\r\nmE(%d) mMeans(%d) mVariances(%d)\r\n"
                 "nA(%d) nB(%d) nC(%d) nD(%d)\r\nnE(%d) nMeans(%d) nVariances(%d)",
                 dir, m[0], m[1], m[2], m[3], m[4], m[5], m[6], nn[0], nn[1], nn[2], nn[3], nn[4], nn[5], nn[6]);
    *numMatrices = num[0];
    *n = m[0];
}

typedef struct { Array a, b, c, d, e; } gp_inputs;

static void call_sharded(gp_fn fn, int gpus, int n, const gp_inputs *in, Array rhs, size_t rhs_unit, Array out, int numMatrices)
{
    if (gpus == 1) { fn(n, in->a, in->b, in->c, rhs, out, numMatrices); return; }
    #pragma omp parallel num_threads(gpus)
    {
        const int g = omp_get_thread_num();
        const long lo = (long)numMatrices * g / gpus, hi = (long)numMatrices * (g + 1) / gpus;
        invgpu_set_device(g);
        if (hi > lo)
            fn(n, in->a + lo * n, in->b + lo * (long)n * n, in->c + lo * n, rhs + lo * (long)rhs_unit, out + lo, (int)(hi - lo));
    }
}

int main(int argc, char const *argv[])
{
    BENCH_ENSURE(argc >= 4, "Usage: gauss_bench TEST_FOLDER TEST_REPLICATIONS MATRIX_DUPLICATES [-csv] "
                            "[--gpus N] [--cpu-lib PATH] [--json]");
    const int numReps = atoi(argv[2]), numDuplicates = atoi(argv[3]);
    bench_opts opt = parse_opts(argc, argv);
    int numMatrices, n;
    Array _a, _b, _c, _d, _e, _means, _variances;
    read_test(argv[1], &numMatrices, &n, &_a, &_b, &_c, &_d, &_e, &_means, &_variances);
    replicateMatrices(&_a, n, 1, numMatrices, numDuplicates);
    replicateMatrices(&_b, n, n, numMatrices, numDuplicates);
    replicateMatrices(&_c, n, 1, numMatrices, numDuplicates);
    replicateMatrices(&_d, n, 1, numMatrices, numDuplicates);
    replicateMatrices(&_e, 1, 1, numMatrices, numDuplicates);
    replicateMatrices(&_means, 1, 1, numMatrices, numDuplicates);
    replicateMatrices(&_variances, 1, 1, numMatrices, numDuplicates);
    numMatrices *= numDuplicates;

    const size_t nv = (size_t)numMatrices * n, nm = (size_t)numMatrices * n * n;
    gp_inputs in;
    bool pin[7];
    in.a = (Array)bench_buffer(nv * sizeof(float), opt.pageable, &pin[0]); in.b = (Array)bench_buffer(nm * sizeof(float), opt.pageable, &pin[1]);
    in.c = (Array)bench_buffer(nv * sizeof(float), opt.pageable, &pin[2]); in.d = (Array)bench_buffer(nv * sizeof(float), opt.pageable, &pin[3]);
    in.e = (Array)bench_buffer((size_t)numMatrices * sizeof(float), opt.pageable, &pin[4]);
    Array means_out = (Array)bench_buffer((size_t)numMatrices * sizeof(float), opt.pageable, &pin[5]);
    Array variances_out = (Array)bench_buffer((size_t)numMatrices * sizeof(float), opt.pageable, &pin[6]);
    BENCH_ENSURE(in.a && in.b && in.c && in.d && in.e && means_out && variances_out, "Could not allocate memory");
#define GP_SETUP()                                                                              \
    do { memcpy(in.a, _a, nv * sizeof(float)); memcpy(in.b, _b, nm * sizeof(float));            \
         memcpy(in.c, _c, nv * sizeof(float)); memcpy(in.d, _d, nv * sizeof(float));            \
         memcpy(in.e, _e, (size_t)numMatrices * sizeof(float)); } while (0)

    if (opt.cpu_lib) {   /* the reference CPU path destroys Bs and Cs, hence the setup before every call */
        void *h = dlopen(opt.cpu_lib, RTLD_NOW | RTLD_GLOBAL);
        BENCH_ENSURE(h, "could not load --cpu-lib %s: %s", opt.cpu_lib, dlerror());
        gp_fn mean_cpu = (gp_fn)dlsym(h, "calcluateMeanCPU"), var_cpu = (gp_fn)dlsym(h, "calcluateVarianceCPU");
        BENCH_ENSURE(mean_cpu && var_cpu, "%s does not export calcluateMeanCPU / calcluateVarianceCPU", opt.cpu_lib);
        bench_timer tm = {0}, tv = {0};
        double em = 0, ev = 0;
        for (int rep = 0; rep < numReps; ++rep) {
            GP_SETUP(); bt_start(&tm); mean_cpu(n, in.a, in.b, in.c, in.d, means_out, numMatrices); bt_stop(&tm);
            em += l1_distance(means_out, _means, numMatrices);
            GP_SETUP(); bt_start(&tv); var_cpu(n, in.a, in.b, in.c, in.e, variances_out, numMatrices); bt_stop(&tv);
            ev += l1_distance(variances_out, _variances, numMatrices);
        }
        bench_report("means_cpu", numMatrices, n, numReps, &tm, em / numMatrices / numReps, opt.csv);
        bench_report("variances_cpu", numMatrices, n, numReps, &tv, ev / numMatrices / numReps, opt.csv);
    }

    BENCH_ENSURE(invgpu_device_count() >= opt.gpus, "%d CUDA device(s) requested, %d usable: this program has no CPU path",
                 opt.gpus, invgpu_device_count());
    bench_timer tm = {0}, tv = {0};
    double em = 0, ev = 0;
    /* rep -1 is an untimed warm-up (CUDA context, kernel loading, first allocation of the host pipeline): the
       reference creates its cuBLAS handle before its timing loops as well (src/gauss_bench.cu:660-661) */
    for (int rep = -1; rep < numReps; ++rep) {
        GP_SETUP();
        if (rep >= 0) bt_start(&tm);
        call_sharded(calcluateMeanGPU, opt.gpus, n, &in, in.d, n, means_out, numMatrices);
        if (rep >= 0) { bt_stop(&tm); em += l1_distance(means_out, _means, numMatrices); }
        GP_SETUP();
        if (rep >= 0) bt_start(&tv);
        call_sharded(calcluateVarianceGPU, opt.gpus, n, &in, in.e, 1, variances_out, numMatrices);
        if (rep >= 0) { bt_stop(&tv); ev += l1_distance(variances_out, _variances, numMatrices); }
    }
    bench_report("means_gpu", numMatrices, n, numReps, &tm, em / numMatrices / numReps, opt.csv);
    bench_report("variances_gpu", numMatrices, n, numReps, &tv, ev / numMatrices / numReps, opt.csv);
    if (opt.json)
        printf("{\"bench\": \"gauss_bench\", \"n\": %d, \"numMatrices\": %d, \"gpus\": %d, \"mean_ms\": %.6f, "
               "\"means_per_s\": %.6e, \"end_to_end\": true}\n", n, numMatrices, opt.gpus, tm.mean, numMatrices / (tm.mean * 1e-3));
    return 0;
}

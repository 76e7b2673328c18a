// xfer_bench -- host<->device transfer micro-benchmark (the engine's counterpart of the reference's
// src/bench.cu:26-158: pageable / pinned / chunked copies), extended to what the end-to-end number of
// this engine needs: N GPUs CONCURRENTLY, both directions at once, NUMA placement on / off, and the
// library's own host pipeline with the kernel replaced by a device copy (invgpu_xfer_roundtrip_host).
//
//   xfer_bench [--gpus N] [--mb M] [--reps R] [--quick]
//
// One CSV line per measurement on stdout:
//   gpus,alloc,numa,chunk_mb,what,GBps_per_gpu_min,GBps_per_gpu_mean,GBps_aggregate
// where `what` is h2d | d2h | bidir (each direction's bytes counted once; bidir = sum of both) |
// pipeline (invgpu_xfer_roundtrip_host: in + out bytes / wall time).  `aggregate` is total bytes of all
// GPUs / the wall time between two barriers around all of them (the honest multi-GPU figure).
#include <cuda_runtime.h>
#include <pthread.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../cuda_matrix_inversion_b200/csrc/host_numa.h"
#include "../include/invgpu.h"

using namespace invgpu;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

enum Alloc { A_PINNED, A_WC, A_REGISTERED, A_PAGEABLE };
static const char *alloc_name(Alloc a) { return a == A_PINNED ? "pinned" : a == A_WC ? "pinned_wc" : a == A_REGISTERED ? "registered" : "pageable"; }

static void *host_buffer(Alloc a, size_t bytes, int dev, bool numa, bool for_input) {
    int node = -1;
    if (numa && numa_node_count() >= 2) node = numa_node_of_device(dev);
    if (node >= 0) numa_prefer_node(node);
    void *p = nullptr;
    if (a == A_PINNED) CK(cudaHostAlloc(&p, bytes, cudaHostAllocDefault));
    else if (a == A_WC) CK(cudaHostAlloc(&p, bytes, for_input ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
    else {
        p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) { perror("mmap"); exit(2); }
        madvise(p, bytes, MADV_HUGEPAGE);
        memset(p, 1, bytes);                                   // first touch under the policy
        if (a == A_REGISTERED) CK(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
    }
    if (a == A_PINNED || a == A_WC) memset(p, 1, bytes);
    if (node >= 0) numa_prefer_node(-1);
    return p;
}
static void host_release(Alloc a, void *p, size_t bytes) {
    if (a == A_PINNED || a == A_WC) { cudaFreeHost(p); return; }
    if (a == A_REGISTERED) cudaHostUnregister(p);
    munmap(p, bytes);
}

struct Shared {
    pthread_barrier_t bar;
    int gpus;
    size_t bytes;
    int reps;
    std::vector<double> secs;      // per GPU
    double wall = 0;
};

struct Case { Alloc alloc; bool numa; size_t chunk; int what; };   // what: 0 h2d, 1 d2h, 2 bidir, 3 pipeline

static void worker(int g, Shared *sh, const Case *cs) {
    CK(cudaSetDevice(g));
    invgpu_set_device(g);
    cpu_set_t saved;
    sched_getaffinity(0, sizeof(saved), &saved);
    if (cs->numa) numa_bind_thread_to_device(g);
    const size_t bytes = sh->bytes;
    void *hin = host_buffer(cs->alloc, bytes, g, cs->numa, true);
    void *hout = host_buffer(cs->alloc, bytes, g, cs->numa, false);
    void *din = nullptr, *dout = nullptr;
    cudaStream_t s0, s1;
    if (cs->what != 3) {
        CK(cudaMalloc(&din, bytes)); CK(cudaMalloc(&dout, bytes));
        CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    }
    auto pass = [&]() {
        if (cs->what == 3) {
            int rc = invgpu_xfer_roundtrip_host(hin, hout, 4096, (invgpu_i64)(bytes / 4096));
            if (rc) { fprintf(stderr, "invgpu_xfer_roundtrip_host: %s\n", invgpu_error_string(rc)); exit(2); }
            return;
        }
        for (size_t off = 0; off < bytes; off += cs->chunk) {
            const size_t len = std::min(cs->chunk, bytes - off);
            if (cs->what == 0 || cs->what == 2) CK(cudaMemcpyAsync((char *)din + off, (char *)hin + off, len, cudaMemcpyHostToDevice, s0));
            if (cs->what == 1 || cs->what == 2) CK(cudaMemcpyAsync((char *)hout + off, (char *)dout + off, len, cudaMemcpyDeviceToHost, s1));
        }
        CK(cudaStreamSynchronize(s0)); CK(cudaStreamSynchronize(s1));
    };
    pass();                                                         // warm-up (page tables, ring allocation)
    pthread_barrier_wait(&sh->bar);
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < sh->reps; ++r) pass();
    const auto t1 = std::chrono::steady_clock::now();
    sh->secs[g] = std::chrono::duration<double>(t1 - t0).count();
    pthread_barrier_wait(&sh->bar);
    if (g == 0) sh->wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (cs->what != 3) { cudaFree(din); cudaFree(dout); cudaStreamDestroy(s0); cudaStreamDestroy(s1); }
    else invgpu_release_workspace();
    host_release(cs->alloc, hin, bytes); host_release(cs->alloc, hout, bytes);
    sched_setaffinity(0, sizeof(saved), &saved);
}

static void run_case(int gpus, size_t bytes, int reps, const Case &cs) {
    Shared sh; sh.gpus = gpus; sh.bytes = bytes; sh.reps = reps; sh.secs.assign(gpus, 0);
    pthread_barrier_init(&sh.bar, nullptr, gpus);
    std::vector<std::thread> th;
    for (int g = 0; g < gpus; ++g) th.emplace_back(worker, g, &sh, &cs);
    for (auto &t : th) t.join();
    pthread_barrier_destroy(&sh.bar);
    const double per_pass = (cs.what >= 2 ? 2.0 : 1.0) * (double)bytes;
    double mn = 1e30, mean = 0;
    for (int g = 0; g < gpus; ++g) { const double v = per_pass * reps / sh.secs[g] / 1e9; mn = std::min(mn, v); mean += v / gpus; }
    static const char *names[] = {"h2d", "d2h", "bidir", "pipeline"};
    printf("%d,%s,%d,%zu,%s,%.2f,%.2f,%.2f\n", gpus, alloc_name(cs.alloc), (int)cs.numa, cs.chunk >> 20, names[cs.what], mn, mean,
           per_pass * reps * gpus / sh.wall / 1e9);
    fflush(stdout);
}

int main(int argc, char **argv) {
    int gpus = 0, reps = 3; size_t mb = 1024; bool quick = false;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--mb") && i + 1 < argc) mb = (size_t)atol(argv[++i]);
        else if (!strcmp(argv[i], "--reps") && i + 1 < argc) reps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--quick")) quick = true;
        else { fprintf(stderr, "usage: xfer_bench [--gpus N] [--mb M] [--reps R] [--quick]\n"); return 1; }
    }
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) { fprintf(stderr, "xfer_bench: no CUDA device\n"); return 2; }
    if (gpus <= 0 || gpus > have) gpus = have;
    fprintf(stderr, "# %d GPU(s) visible, using up to %d; %d NUMA node(s); host CPUs %ld\n", have, gpus, numa_node_count(), sysconf(_SC_NPROCESSORS_ONLN));
    for (int g = 0; g < have; ++g) {
        char bus[32] = {0}; cudaDeviceGetPCIBusId(bus, sizeof(bus), g);
        fprintf(stderr, "# gpu %d pci %s numa_node %d\n", g, bus, numa_node_of_device(g));
    }
    printf("gpus,alloc,numa,chunk_mb,what,GBps_per_gpu_min,GBps_per_gpu_mean,GBps_aggregate\n");
    const size_t bytes = mb << 20;
    std::vector<int> counts;
    for (int n = 1; n <= gpus; n *= 2) counts.push_back(n);
    if (counts.back() != gpus) counts.push_back(gpus);
    for (int n : counts) {
        for (int numa = 1; numa >= 0; --numa) {
            if (numa == 0 && numa_node_count() < 2 && n > 1) continue;          // identical by construction
            for (int what = 0; what < 3; ++what) run_case(n, bytes, reps, Case{A_PINNED, numa != 0, (size_t)32 << 20, what});
            run_case(n, bytes, reps, Case{A_PINNED, numa != 0, (size_t)32 << 20, 3});
            if (quick) continue;
            run_case(n, bytes, reps, Case{A_REGISTERED, numa != 0, (size_t)32 << 20, 2});
            run_case(n, bytes, reps, Case{A_WC, numa != 0, (size_t)32 << 20, 2});
            if (numa) {
                run_case(n, bytes, reps, Case{A_PINNED, true, (size_t)8 << 20, 2});
                run_case(n, bytes, reps, Case{A_PINNED, true, (size_t)128 << 20, 2});
                run_case(n, bytes, reps, Case{A_PINNED, true, bytes, 2});
                if (n == 1) run_case(n, bytes, 1, Case{A_PAGEABLE, true, (size_t)32 << 20, 3});
            }
        }
    }
    return 0;
}

// host_numa.h -- NUMA placement of the pinned staging memory and of the host thread that feeds a GPU.
//
// The end-to-end path (host buffers in, host buffers out) is bound by host<->device DMA, and on a
// two-socket 8-GPU box a pinned page that lives on the other socket crosses the inter-socket link on
// every transfer.  Everything here is best effort and silent: on a VM that hides the topology
// (numa_node = -1, a single node) the calls do nothing.  Plain Linux syscalls, no libnuma.
//   INVGPU_NUMA=0 switches all of it off (tools/xfer_bench measures both).
#pragma once

#include <cuda_runtime.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace invgpu {

static inline bool numa_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("INVGPU_NUMA"); v = (e && !strcmp(e, "0")) ? 0 : 1; }
    return v == 1;
}

// NUMA node of a CUDA device from sysfs (/sys/bus/pci/devices/<bus id>/numa_node); -1 = unknown.
static inline int numa_node_of_device(int dev) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char *p = bus; *p; ++p) if (*p >= 'A' && *p <= 'Z') *p += 'a' - 'A';
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

static inline int numa_node_count() {
    int n = 0;
    for (;; ++n) {
        char path[96];
        snprintf(path, sizeof(path), "/sys/devices/system/node/node%d", n);
        if (access(path, F_OK) != 0) break;
    }
    return n;
}

// CPUs of a node ("0-31,64-95") as a cpu_set_t; false when unknown.
static inline bool numa_node_cpus(int node, cpu_set_t *set) {
    char path[96];
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = fopen(path, "r");
    if (!f) return false;
    char buf[4096] = {0};
    const bool ok = fgets(buf, sizeof(buf), f) != nullptr;
    fclose(f);
    if (!ok) return false;
    CPU_ZERO(set);
    int any = 0;
    for (char *p = buf; *p && *p != '\n';) {
        char *end;
        long a = strtol(p, &end, 10), b = a;
        if (end == p) break;
        p = end;
        if (*p == '-') { b = strtol(p + 1, &end, 10); p = end; }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, set); ++any; }
        if (*p == ',') ++p;
    }
    return any > 0;
}

// MPOL_PREFERRED (1) on `node` for the calling thread's future page allocations; node < 0 restores the default.
static inline void numa_prefer_node(int node) {
#ifdef SYS_set_mempolicy
    if (node < 0) { syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0UL); return; }
    unsigned long mask[16] = {0};
    if (node >= (int)(sizeof(mask) * 8)) return;
    mask[node / (8 * sizeof(unsigned long))] |= 1UL << (node % (8 * sizeof(unsigned long)));
    syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, (unsigned long)(sizeof(mask) * 8));
#else
    (void)node;
#endif
}

// Restrict the calling thread to the CPUs of the device's node (only when the box has more than one node and
// the thread's current mask is not already a subset).  Returns the node used or -1.
static inline int numa_bind_thread_to_device(int dev) {
    if (!numa_enabled()) return -1;
    const int node = numa_node_of_device(dev);
    if (node < 0 || numa_node_count() < 2) return -1;
    cpu_set_t want, have, both;
    if (!numa_node_cpus(node, &want)) return -1;
    if (sched_getaffinity(0, sizeof(have), &have) != 0) return -1;
    CPU_AND(&both, &want, &have);
    if (CPU_COUNT(&both) == 0) return -1;               // the launcher pinned us elsewhere on purpose: keep it
    sched_setaffinity(0, sizeof(both), &both);
    return node;
}

// cudaHostAlloc whose pages are first-touched on the device's node.
static inline cudaError_t numa_host_alloc(void **p, size_t bytes, int dev, unsigned flags) {
    int node = -1;
    if (numa_enabled() && numa_node_count() >= 2) node = numa_node_of_device(dev);
    if (node >= 0) numa_prefer_node(node);
    const cudaError_t e = cudaHostAlloc(p, bytes, flags);
    if (node >= 0) numa_prefer_node(-1);
    return e;
}

}  // namespace invgpu

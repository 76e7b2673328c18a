// engine.cuh -- host-side runtime shared by the C-ABI layer: per-device state, launch
// accounting, persistent-grid sizing and the pinned-ring host<->device pipeline that replaces
// the reference's per-call cudaHostAlloc / cudaMallocPitch / cudaMemcpy2D / cudaFree
// (e.g. src/gauss/batched_invert.cu:120-176, src/gauss_bench.cu:160-264).
#pragma once

#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace invgpu {

extern std::atomic<long long> g_launches;

#define INVGPU_TRY(expr)                                   \
    do {                                                   \
        cudaError_t e__ = (expr);                          \
        if (e__ != cudaSuccess) return (int)e__;           \
    } while (0)

struct DeviceState {
    int dev = -1;
    int sms = 0;
    size_t smem_optin = 0;
    // host pipeline resources (lazily created, grow-only)
    static const int kSlots = 3;
    void *d_ws[kSlots] = {nullptr, nullptr, nullptr};
    size_t d_ws_bytes = 0;
    void *h_ring[kSlots] = {nullptr, nullptr, nullptr};
    size_t h_ring_bytes = 0;
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[kSlots], ev_comp[kSlots], ev_out[kSlots];
    static const int kMixedTiers = 9;                         // 16 / 24 / 32 / 48 / 64 / 96 / 128 / 192 / 256
    cudaEvent_t ev_mixed[kMixedTiers];                        // one per tier of the mixed-dimension scheduler
    bool streams_ready = false;
    // scratch of the fused GP tile kernels (natural-order info recomputation of flagged matrices)
    void *gp_scratch = nullptr;
    size_t gp_scratch_bytes = 0;
    // work lists of the mixed-dimension scheduler (device copy + pinned staging)
    void *d_mixed = nullptr, *h_mixed = nullptr;
    size_t mixed_bytes = 0;
};

DeviceState *device_state(int *err);   // state of the calling thread's current device
std::mutex &engine_mutex();

// Persistent grid: enough CTAs to fill every SM at the kernel's occupancy, never more than the
// work needs.  148 SMs x resident CTAs per SM on B200.
template <typename Kern>
static int persistent_grid(Kern kern, int block, size_t smem, i64 blocks_needed, const DeviceState *ds, int *grid) {
    if (smem > ds->smem_optin) return -2;
    if (smem > 48 * 1024)
        INVGPU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    INVGPU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block, smem));
    if (occ < 1) return -2;
    i64 g = (i64)occ * ds->sms;
    if (g > blocks_needed) g = blocks_needed;
    if (g < 1) g = 1;
    *grid = (int)g;
    return 0;
}

}  // namespace invgpu

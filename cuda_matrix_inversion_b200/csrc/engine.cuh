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
    std::atomic<int> dev{-1};         // published (release) after sms / smem_optin are written
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
    // scratch of the fused GP tile kernels (natural-order info recomputation of flagged matrices): one buffer per
    // launch stream, so that GP calls on different streams never share it; a buffer that has to grow is retired
    // (freed by invgpu_release_workspace), never freed under a launch that may still read it
    struct Scratch { cudaStream_t st; void *p; size_t bytes; };
    std::vector<Scratch> gp_scratch;
    std::vector<void *> retired;
    std::mutex scratch_mu;
    // work lists of the mixed-dimension scheduler: a ring of (device copy + pinned staging) pairs, each guarded by
    // the event recorded after the LAST tier kernel that reads it
    static const int kMixedRing = 4;
    struct MixedBuf { void *d = nullptr, *h = nullptr; size_t bytes = 0; cudaEvent_t done = nullptr, uploaded = nullptr; unsigned long long seq = 0; };
    MixedBuf mixed[kMixedRing];
    unsigned mixed_next = 0;
    // one lock per device: the host pipeline and the mixed scheduler of different devices run concurrently
    std::mutex mu;
    int numa_node = -1;
};

DeviceState *device_state(int *err);   // state of the calling thread's current device
std::mutex &init_mutex();          // first-time initialisation of a DeviceState only

// Per-stream device scratch (natural-order info recomputation of flagged GP matrices; working copies of the any-n
// kernels when they exceed shared memory).  Launches on one stream are ordered, so they may share a buffer.
static int ensure_gp_scratch(DeviceState *ds, size_t need, cudaStream_t st, void **out) {
    std::lock_guard<std::mutex> lk(ds->scratch_mu);
    DeviceState::Scratch *s = nullptr;
    for (auto &e : ds->gp_scratch) if (e.st == st) s = &e;
    if (!s) { ds->gp_scratch.push_back(DeviceState::Scratch{st, nullptr, 0}); s = &ds->gp_scratch.back(); }
    if (s->bytes < need) {
        if (s->p) ds->retired.push_back(s->p);        // an earlier launch on this stream may still be running
        s->p = nullptr; s->bytes = 0;
        INVGPU_TRY(cudaMalloc(&s->p, need));
        s->bytes = need;
    }
    *out = s->p;
    return 0;
}

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

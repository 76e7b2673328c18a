// tensor-core tier (tcgen05 / TMEM), fp32 n = 128: kernels of tc_kernels.cuh and their launchers
#include <stdlib.h>
#include "tile_launch.cuh"
#include "tc_kernels.cuh"

namespace invgpu {

template <int PW>
static int launch_tc_gp128_pw(GpIO<float> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds, int per_sm) {
    auto kern = tc::tc_gp128_kernel<PW, 4>;
    i64 g = (i64)per_sm * ds->sms;
    if (g > batch) g = batch;
    kern<<<(int)g, 128, tc::GpGeo<PW>::SMEM_BYTES, st>>>(io, batch, dInfo);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int launch_tc_gp128(GpIO<float> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    // 4 CTAs per SM by construction (<= 128 registers x 128 threads, <= 37 KB shared memory, 128 of the 512 TMEM columns
    // each).  The occupancy API answers 1 for this kernel on CUDA 12.9 (profiles/r2_tc_gp128_v1_summary.md), so the
    // persistent grid is sized by hand; INVGPU_TC_CTAS_PER_SM / INVGPU_TC_PANEL (16 | 32) override (experiments).
    static int per_sm = -1, pw = -1;
    if (per_sm < 0) { const char *e = getenv("INVGPU_TC_CTAS_PER_SM"); per_sm = (e && atoi(e) > 0) ? atoi(e) : 4; }
    if (pw < 0) { const char *e = getenv("INVGPU_TC_PANEL"); pw = (e && atoi(e) == 16) ? 16 : 32; }
    if (pw == 16) return launch_tc_gp128_pw<16>(io, batch, dInfo, st, ds, per_sm);
    return launch_tc_gp128_pw<32>(io, batch, dInfo, st, ds, per_sm);
}

}  // namespace invgpu

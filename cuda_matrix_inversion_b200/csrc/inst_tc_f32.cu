// tensor-core tier (tcgen05 / TMEM), fp32 n = 128: kernels of tc_kernels.cuh and their launchers
#include <stdlib.h>
#include "tile_launch.cuh"
#include "tc_kernels.cuh"

namespace invgpu {

template <int PW, int NM>
static int launch_tc_gp128_cfg(GpIO<float> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds, int per_sm) {
    auto kern = tc::tc_gp128_kernel<PW, NM, 4 / NM>;
    const size_t smem = tc::GpGeoM<PW, NM>::SMEM_BYTES;
    if (smem > 48 * 1024) INVGPU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    i64 g = (i64)per_sm * ds->sms;
    const i64 groups = (batch + NM - 1) / NM;
    if (g > groups) g = groups;
    kern<<<(int)g, 128, smem, st>>>(io, batch, dInfo);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int launch_tc_gp128(GpIO<float> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    // Residency by construction: NM evaluations per CTA x (4 / NM) CTAs per SM = the four 128-column accumulators that fit
    // the 512 TMEM columns (<= 128 x NM registers, <= 37 x NM KB shared memory per CTA; 8 KB operands + 18 KB stage + 2 KB at the default panel 16).  The occupancy API answers 1 for
    // these kernels on CUDA 12.9 (profiles/r2_tc_gp128_summary.md), so the persistent grid is sized by hand.
    // Measured (200 000 evaluations, right-hand sides outside the pivot loop): 1 evaluation per CTA, panel 16: 0.243 of the roofline
    // (default), panel 32: 0.221; 2 per CTA: 0.157.
    // INVGPU_TC_GROUP (1 | 2 evaluations per CTA), INVGPU_TC_PANEL (16 | 32), INVGPU_TC_CTAS_PER_SM: experiments.
    static int per_sm = -1, pw = -1, nm = -1;
    if (nm < 0) { const char *e = getenv("INVGPU_TC_GROUP"); nm = (e && atoi(e) == 2) ? 2 : 1; }
    if (per_sm < 0) { const char *e = getenv("INVGPU_TC_CTAS_PER_SM"); per_sm = (e && atoi(e) > 0) ? atoi(e) : 4 / nm; }
    if (pw < 0) { const char *e = getenv("INVGPU_TC_PANEL"); pw = (e && atoi(e) == 32) ? 32 : 16; }
    if (nm == 1) return pw == 16 ? launch_tc_gp128_cfg<16, 1>(io, batch, dInfo, st, ds, per_sm) : launch_tc_gp128_cfg<32, 1>(io, batch, dInfo, st, ds, per_sm);
    return pw == 16 ? launch_tc_gp128_cfg<16, 2>(io, batch, dInfo, st, ds, per_sm) : launch_tc_gp128_cfg<32, 2>(io, batch, dInfo, st, ds, per_sm);
}

}  // namespace invgpu

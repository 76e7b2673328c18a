// tile_launch.cuh -- host launchers of the register-tiled kernels.  Declared everywhere, defined
// (and explicitly instantiated) only in the inst_*.cu translation units.
#pragma once

#include "engine.cuh"
#include "generic_smem.cuh"   // SPD_* stage bits

namespace invgpu {

template <typename T, int N, int TR, int TC, bool PERM, int STAGES, int MINB>
int launch_tile_spd(StridedIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

template <typename T, int N, int TR, int TC, int MINB>
int launch_tile_gp(GpIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

template <typename T, int N, int ROWS, typename IO, int MINB>
int launch_gj(IO io, int n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

template <typename T, int N, int TR, int TC, bool STAGE, int MINB>
int launch_onesweep(StridedIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

template <typename T, int N, int TR, int TC, bool UNROLL, int MINB>
int launch_sweep(StridedIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

template <typename T, int N, int TR, int TC, bool UNROLL, int MINB>
int launch_sweep_gp(GpIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

template <typename T, int N, int TR, int TC, int MINB>
int launch_sweep_pad(PadIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

}  // namespace invgpu

#ifdef INVGPU_TILE_DEFINE
#include "tile_kernels.cuh"
#include "gj_kernels.cuh"
#include "onesweep_kernels.cuh"
#include "sweep_kernels.cuh"

namespace invgpu {

template <typename T, int N, int TR, int TC, bool UNROLL, int MINB>
int launch_sweep(StridedIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using SG = SweepGeo<N, TR, TC>;
    auto kern = sweep_spd_kernel<T, N, TR, TC, UNROLL, StridedIO<T>, MINB>;
    const size_t smem = (size_t)SG::MPB * SG::WORDS * sizeof(T);
    int grid = 0;
    int rc = persistent_grid(kern, SG::BLOCK, smem, (batch + SG::MPB - 1) / SG::MPB, ds, &grid);
    if (rc) return rc;
    kern<<<grid, SG::BLOCK, smem, st>>>(io, batch, dInfo);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

template <typename T, int N, int TR, int TC, int MINB>
int launch_sweep_pad(PadIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using SG = SweepGeo<N, TR, TC>;
    auto kern = sweep_spd_kernel<T, N, TR, TC, false, PadIO<T>, MINB>;
    const size_t smem = (size_t)SG::MPB * SG::WORDS * sizeof(T);
    int grid = 0;
    int rc = persistent_grid(kern, SG::BLOCK, smem, (batch + SG::MPB - 1) / SG::MPB, ds, &grid);
    if (rc) return rc;
    kern<<<grid, SG::BLOCK, smem, st>>>(io, batch, dInfo);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

// scratch for the (rare) natural-order info recomputation of flagged GP matrices
static int ensure_gp_scratch(DeviceState *ds, size_t need) {
    if (ds->gp_scratch_bytes < need) {
        static std::mutex scratch_mutex;              // not engine_mutex(): the host pipeline holds that one
        std::lock_guard<std::mutex> lk(scratch_mutex);
        if (ds->gp_scratch_bytes < need) {
            if (ds->gp_scratch) cudaFree(ds->gp_scratch);
            ds->gp_scratch = nullptr; ds->gp_scratch_bytes = 0;
            INVGPU_TRY(cudaMalloc(&ds->gp_scratch, need));
            ds->gp_scratch_bytes = need;
        }
    }
    return 0;
}

template <typename T, int N, int TR, int TC, bool UNROLL, int MINB>
int launch_sweep_gp(GpIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using SG = SweepGeo<N, TR, TC>;
    auto kern = sweep_gp_kernel<T, N, TR, TC, UNROLL, MINB>;
    const size_t smem = (size_t)SG::MPB * SG::WORDS * sizeof(T);
    int grid = 0;
    int rc = persistent_grid(kern, SG::BLOCK, smem, (batch + SG::MPB - 1) / SG::MPB, ds, &grid);
    if (rc) return rc;
    rc = ensure_gp_scratch(ds, (size_t)grid * SG::MPB * N * N * sizeof(T));
    if (rc) return rc;
    kern<<<grid, SG::BLOCK, smem, st>>>(io, batch, dInfo, (T *)ds->gp_scratch);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

template <typename T, int N, int TR, int TC, bool STAGE, int MINB>
int launch_onesweep(StridedIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using G = TileGeo<N, TR, TC, false>;
    auto kern = onesweep_spd_kernel<T, N, TR, TC, STAGE, StridedIO<T>, MINB>;
    constexpr int WORDS = ((2 * N + 31) / 32) * 32 + (G::LANES < 32 ? 8 : 0) + (STAGE ? TileStage<T, N, TR, TC>::MATRIX_WORDS : 0);
    const size_t smem = (size_t)G::MPB * WORDS * sizeof(T);
    int grid = 0;
    int rc = persistent_grid(kern, G::BLOCK, smem, (batch + G::MPB - 1) / G::MPB, ds, &grid);
    if (rc) return rc;
    kern<<<grid, G::BLOCK, smem, st>>>(io, batch, dInfo);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

template <typename T, int N, int ROWS, typename IO, int MINB>
int launch_gj(IO io, int n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using G = GjGeo<T, N, ROWS>;
    auto kern = gj_rowlane_kernel<T, N, ROWS, IO, MINB>;
    const size_t smem = (size_t)G::MPB * G::WORDS * sizeof(T);
    int grid = 0;
    int rc = persistent_grid(kern, G::BLOCK, smem, (batch + G::MPB - 1) / G::MPB, ds, &grid);
    if (rc) return rc;
    kern<<<grid, G::BLOCK, smem, st>>>(io, n, batch, dInfo);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

template <typename T, int N, int TR, int TC, bool PERM, int STAGES, int MINB>
int launch_tile_spd(StridedIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using G = TileGeo<N, TR, TC, PERM>;
    auto kern = tile_spd_kernel<T, N, TR, TC, PERM, StridedIO<T>, STAGES, MINB>;
    const size_t smem = (size_t)G::MPB * G::SMEM_WORDS * sizeof(T);
    int grid = 0;
    int rc = persistent_grid(kern, G::BLOCK, smem, (batch + G::MPB - 1) / G::MPB, ds, &grid);
    if (rc) return rc;
    kern<<<grid, G::BLOCK, smem, st>>>(io, batch, dInfo);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

template <typename T, int N, int TR, int TC, int MINB>
int launch_tile_gp(GpIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using G = TileGeo<N, TR, TC, true>;
    auto kern = tile_gp_kernel<T, N, TR, TC, MINB>;
    constexpr int WORDS = 2 * (N + 4) + (G::LANES < 32 ? 8 : 0) + ((2 * (N + 4)) % 32 == 0 ? 0 : 32 - (2 * (N + 4)) % 32);
    const size_t smem = (size_t)G::MPB * WORDS * sizeof(T);
    int grid = 0;
    int rc = persistent_grid(kern, G::BLOCK, smem, (batch + G::MPB - 1) / G::MPB, ds, &grid);
    if (rc) return rc;
    rc = ensure_gp_scratch(ds, (size_t)grid * G::MPB * N * N * sizeof(T));
    if (rc) return rc;
    kern<<<grid, G::BLOCK, smem, st>>>(io, batch, dInfo, (T *)ds->gp_scratch);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

}  // namespace invgpu

#define INVGPU_TILE_INSTANTIATE(T, N, TR, TC, PERM, STAGES, MINB) \
    template int invgpu::launch_tile_spd<T, N, TR, TC, PERM, STAGES, MINB>(invgpu::StridedIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_GJ_INSTANTIATE(T, N, ROWS, MINB) \
    template int invgpu::launch_gj<T, N, ROWS, invgpu::StridedIO<T>, MINB>(invgpu::StridedIO<T>, int, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *); \
    template int invgpu::launch_gj<T, N, ROWS, invgpu::PtrIO<T>, MINB>(invgpu::PtrIO<T>, int, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_ONESWEEP_INSTANTIATE(T, N, TR, TC, STAGE, MINB) \
    template int invgpu::launch_onesweep<T, N, TR, TC, STAGE, MINB>(invgpu::StridedIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_SWEEP_PAD_INSTANTIATE(T, N, TR, TC, MINB) \
    template int invgpu::launch_sweep_pad<T, N, TR, TC, MINB>(invgpu::PadIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_SWEEP_GP_INSTANTIATE(T, N, TR, TC, UNROLL, MINB) \
    template int invgpu::launch_sweep_gp<T, N, TR, TC, UNROLL, MINB>(invgpu::GpIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_SWEEP_INSTANTIATE(V, T, N, TR, TC, UNROLL, MINB) \
    template int invgpu::launch_sweep<T, N, TR, TC, UNROLL, MINB>(invgpu::StridedIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_TILE_INSTANTIATE_GP(T, N, TR, TC, MINB) \
    template int invgpu::launch_tile_gp<T, N, TR, TC, MINB>(invgpu::GpIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#endif

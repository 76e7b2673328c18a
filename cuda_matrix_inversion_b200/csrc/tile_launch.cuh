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

template <typename T, int N, int TR, int TC, typename IO, int MINB>
int launch_gj_tile(IO io, int n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

template <typename T, int N, int ROWS, typename IO, int MINB>
int launch_gj_roll(IO io, int n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

template <typename T, int N, int CW, typename IO, int MINB>
int launch_gj_roll2d(IO io, int n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

// INVGPU_GJR2_WS=1 (lab builds) selects the warp-specialised kernel (4 FMA warps + a pivot warp) instead of the four-warp one
static inline bool gj_roll2d_ws_enabled() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("INVGPU_GJR2_WS"); on = (e && *e == '1') ? 1 : 0; }
    return on != 0;
}

template <typename T, int N, typename IO, int MINB>
int launch_gj_roll2d_ws(IO io, int n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

template <typename T, int N, int CL, int WARPS, int MINB>
int launch_gj_colsplit(StridedIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

template <typename T, int N, int TR, int TC, bool STAGE, int MINB>
int launch_onesweep(StridedIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

template <typename T, int N, int TR, int TC, bool UNROLL, int MINB, int BLK>
int launch_sweep(StridedIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

template <typename T, int N, int TR, int TC, bool UNROLL, int MINB, int BLK>
int launch_sweep_gp(GpIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

template <typename T, int N, int TR, int TC, int MINB>
int launch_sweep_pad(PadIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

// returns INVGPU_TMA_UNAVAILABLE when the batch cannot be described to the TMA unit (strides, size, driver)
#define INVGPU_TMA_UNAVAILABLE (-1001)
template <typename T, int N, int TR, int TC, bool UNROLL, int MINB, bool DIRECT_OUT, bool INTERLEAVE>
int launch_sweep_tma(StridedIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

template <typename T, int NBUF, int MINB, bool GENERAL>
int launch_spd8_tma(StridedIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

template <typename T, int N, int WARPS, int MINB>
int launch_spd_thread_bulk(StridedIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

template <typename T, int N, int WARPS, int MINB>
int launch_gp_thread(GpIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

// tensor-core tier (tc_kernels.cuh, defined in inst_tc_f32.cu): fused GP mean / variance, n = 128 fp32
int launch_tc_gp128(GpIO<float> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

}  // namespace invgpu

#ifdef INVGPU_TILE_DEFINE
#include <cuda.h>
#include <type_traits>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include "tile_kernels.cuh"
#include "gj_kernels.cuh"
#include "gj_roll_kernels.cuh"
#include "gj_roll2d_kernels.cuh"
#include "onesweep_kernels.cuh"
#include "sweep_kernels.cuh"
#include "gj_tile_kernels.cuh"

namespace invgpu {

template <typename T, int N, int TR, int TC, bool UNROLL, int MINB, int BLK>
int launch_sweep(StridedIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using SG = SweepGeo<N, TR, TC>;
    auto kern = sweep_spd_kernel<T, N, TR, TC, UNROLL, StridedIO<T>, MINB, BLK>;
    const size_t smem = (size_t)SG::MPB * (BLK == 2 ? SG::WORDS2 : SG::WORDS) * sizeof(T);
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

// ---- TMA descriptors: cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// the batch as a 2-D tensor [N rows (contiguous) x N*batch columns], box = `mats` whole matrices, 128-byte swizzle
template <typename T, int N>
static int make_batch_tensor_map(void *out128, const T *base, i64 batch, int mats) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return INVGPU_TMA_UNAVAILABLE;
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)N * (cuuint64_t)batch};
    const cuuint64_t strides[1] = {(cuuint64_t)N * sizeof(T)};
    const cuuint32_t box[2] = {(cuuint32_t)N, (cuuint32_t)(N * mats)};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    CUtensorMap m;
    const CUresult r = enc(&m, dt, 2, const_cast<T *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return INVGPU_TMA_UNAVAILABLE;
    memcpy(out128, &m, 128);
    return 0;
}

// any dense batch as a 2-D tensor of 128-byte lines, box = `box_lines` lines, 128-byte swizzle
template <typename T>
static int make_lines_tensor_map(void *out128, const T *base, i64 total_lines, int box_lines) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return INVGPU_TMA_UNAVAILABLE;
    constexpr int EPL = 128 / (int)sizeof(T);
    const cuuint64_t dims[2] = {(cuuint64_t)EPL, (cuuint64_t)total_lines};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {(cuuint32_t)EPL, (cuuint32_t)box_lines};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    CUtensorMap m;
    const CUresult r = enc(&m, dt, 2, const_cast<T *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return INVGPU_TMA_UNAVAILABLE;
    memcpy(out128, &m, 128);
    return 0;
}

template <typename T, int N, int WARPS, int MINB>
int launch_spd_thread_bulk(StridedIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using G = ThreadBulkGeo<T, N, WARPS>;
    auto kern = spd_thread_bulk_kernel<T, N, WARPS, MINB>;
    int grid = 0;
    int rc = persistent_grid(kern, G::BLOCK, G::SMEM, (batch + G::MPB - 1) / G::MPB, ds, &grid);
    if (rc) return rc;
    kern<<<grid, G::BLOCK, G::SMEM, st>>>(io.in, io.in_stride, io.out, io.out_stride, batch, dInfo);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

template <typename T, int N, int WARPS, int MINB>
int launch_gp_thread(GpIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using G = ThreadBulkGeo<T, N, WARPS>;
    auto kern = gp_thread_kernel<T, N, WARPS, MINB>;
    int grid = 0;
    int rc = persistent_grid(kern, G::BLOCK, G::SMEM, (batch + G::MPB - 1) / G::MPB, ds, &grid);
    if (rc) return rc;
    kern<<<grid, G::BLOCK, G::SMEM, st>>>(io, batch, dInfo);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

template <typename T, int NBUF, int MINB, bool GENERAL>
int launch_spd8_tma(StridedIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using G = Small8Geo<T, NBUF>;
    if (io.in_stride != 64 || io.out_stride != 64) return INVGPU_TMA_UNAVAILABLE;
    const i64 lines = batch * G::LINES_PER_MAT;
    if (lines > 0x7fffffffLL) return INVGPU_TMA_UNAVAILABLE;
    TmaMaps maps;
    int rc = make_lines_tensor_map<T>(maps.in, io.in, lines, 32 * G::LINES_PER_MAT);
    if (rc) return rc;
    rc = make_lines_tensor_map<T>(maps.out, io.out, lines, 32 * G::LINES_PER_MAT);
    if (rc) return rc;
    auto kern = spd8_tma_kernel<T, NBUF, MINB, typename std::conditional<GENERAL, Gj8Math<T>, Spd8Math<T>>::type>;
    int grid = 0;
    rc = persistent_grid(kern, G::BLOCK, G::SMEM, (batch + G::MPB - 1) / G::MPB, ds, &grid);
    if (rc) return rc;
    static int trace = -1;
    if (trace < 0) { const char *e = getenv("INVGPU_TRACE"); trace = (e && atoi(e) > 0) ? 1 : 0; }
    if (trace) fprintf(stderr, "[invgpu] spd8_tma_kernel<%s, nbuf=%d, %s> grid %d smem %zu\n", sizeof(T) == 4 ? "f32" : "f64", NBUF, GENERAL ? "gauss-jordan" : "spd", grid, (size_t)G::SMEM);
    kern<<<grid, G::BLOCK, G::SMEM, st>>>(maps, io.in, io.out, batch, dInfo);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

template <typename T, int N, int TR, int TC, bool UNROLL, int MINB, bool DIRECT_OUT, bool INTERLEAVE>
int launch_sweep_tma(StridedIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using SG = SweepGeo<N, TR, TC>;
    using TG = SweepTmaGeo<T, N, TR, TC>;
    if (io.in_stride != (i64)N * N || io.out_stride != (i64)N * N) return INVGPU_TMA_UNAVAILABLE;   // columns must be equidistant
    if ((i64)N * batch > 0x7fffffffLL || TG::MPW * N > 256) return INVGPU_TMA_UNAVAILABLE;          // 32-bit box coordinates
    TmaMaps maps;
    int rc = make_batch_tensor_map<T, N>(maps.in, io.in, batch, TG::MPW);
    if (rc) return rc;
    rc = make_batch_tensor_map<T, N>(maps.out, io.out, batch, TG::MPW);
    if (rc) return rc;
    rc = make_batch_tensor_map<T, N>(maps.in1, io.in, batch, 1);
    if (rc) return rc;
    rc = make_batch_tensor_map<T, N>(maps.out1, io.out, batch, 1);
    if (rc) return rc;
    auto kern = sweep_spd_tma_kernel<T, N, TR, TC, UNROLL, MINB, DIRECT_OUT, INTERLEAVE>;
    constexpr size_t SMEM = TG::template smem<INTERLEAVE>();
    int grid = 0;
    rc = persistent_grid(kern, SG::BLOCK, SMEM, (batch + SG::MPB - 1) / SG::MPB, ds, &grid);
    if (rc) return rc;
    static int trace = -1;
    if (trace < 0) { const char *e = getenv("INVGPU_TRACE"); trace = (e && atoi(e) > 0) ? 1 : 0; }
    if (trace) fprintf(stderr, "[invgpu] sweep_spd_tma_kernel<n=%d, %dx%d, direct_out=%d, interleave=%d> grid %d smem %zu\n", N, TR, TC, (int)DIRECT_OUT, (int)INTERLEAVE, grid, SMEM);
    kern<<<grid, SG::BLOCK, SMEM, st>>>(maps, io.in, io.in_stride, io.out, io.out_stride, batch, dInfo);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

template <typename T, int N, int TR, int TC, bool UNROLL, int MINB, int BLK>
int launch_sweep_gp(GpIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using SG = SweepGeo<N, TR, TC>;
    auto kern = sweep_gp_kernel<T, N, TR, TC, UNROLL, MINB, BLK>;
    const size_t smem = (size_t)SG::MPB * (BLK == 2 ? SG::WORDS2 : SG::WORDS) * sizeof(T);
    int grid = 0;
    int rc = persistent_grid(kern, SG::BLOCK, smem, (batch + SG::MPB - 1) / SG::MPB, ds, &grid);
    if (rc) return rc;
    void *scratch = nullptr;
    rc = ensure_gp_scratch(ds, (size_t)grid * SG::MPB * N * N * sizeof(T), st, &scratch);
    if (rc) return rc;
    kern<<<grid, SG::BLOCK, smem, st>>>(io, batch, dInfo, (T *)scratch);
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

template <typename T, int N, int ROWS, typename IO, int MINB>
int launch_gj_roll(IO io, int n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using G = GjRollGeo<T, N, ROWS>;
    auto kern = (n == N) ? gj_roll_kernel<T, N, ROWS, IO, MINB, true> : gj_roll_kernel<T, N, ROWS, IO, MINB, false>;
    const size_t smem = (size_t)G::MPB * G::WORDS * sizeof(T);
    int grid = 0;
    int rc = persistent_grid(kern, G::BLOCK, smem, (batch + G::MPB - 1) / G::MPB, ds, &grid);
    if (rc) return rc;
    kern<<<grid, G::BLOCK, smem, st>>>(io, n, batch, dInfo);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

template <typename T, int N, int CW, typename IO, int MINB>
int launch_gj_roll2d(IO io, int n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using G = GjRoll2dGeo<T, N, CW>;
    auto kern = (n == N) ? gj_roll2d_kernel<T, N, CW, IO, MINB, true> : gj_roll2d_kernel<T, N, CW, IO, MINB, false>;
    const size_t smem = (size_t)G::WORDS * sizeof(T);
    int grid = 0;
    int rc = persistent_grid(kern, G::BLOCK, smem, batch, ds, &grid);
    if (rc) return rc;
    kern<<<grid, G::BLOCK, smem, st>>>(io, n, batch, dInfo);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

template <typename T, int N, typename IO, int MINB>
int launch_gj_roll2d_ws(IO io, int n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using G = GjRoll2dGeo<T, N, 32>;
    auto kern = (n == N) ? gj_roll2d_ws_kernel<T, N, IO, MINB, true> : gj_roll2d_ws_kernel<T, N, IO, MINB, false>;
    const size_t smem = (size_t)(G::WORDS + N) * sizeof(T);
    int grid = 0;
    int rc = persistent_grid(kern, 256, smem, batch, ds, &grid);
    if (rc) return rc;
    kern<<<grid, 256, smem, st>>>(io, n, batch, dInfo);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

template <typename T, int N, int CL, int WARPS, int MINB>
int launch_gj_colsplit(StridedIO<T> io, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using G = GjcGeo<T, N, CL, WARPS>;
    auto kern = gj_colsplit_kernel<T, N, CL, WARPS, MINB>;
    int grid = 0;
    int rc = persistent_grid(kern, G::BLOCK, G::SMEM, (batch + G::MPB - 1) / G::MPB, ds, &grid);
    if (rc) return rc;
    kern<<<grid, G::BLOCK, G::SMEM, st>>>(io.in, io.in_stride, io.out, io.out_stride, batch, dInfo);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

template <typename T, int N, int TR, int TC, typename IO, int MINB>
int launch_gj_tile(IO io, int n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    using G = GjtGeo<T, N, TR, TC>;
    auto kern = gj_tile_kernel<T, N, TR, TC, IO, MINB>;
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
    void *scratch = nullptr;
    rc = ensure_gp_scratch(ds, (size_t)grid * G::MPB * N * N * sizeof(T), st, &scratch);
    if (rc) return rc;
    kern<<<grid, G::BLOCK, smem, st>>>(io, batch, dInfo, (T *)scratch);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

}  // namespace invgpu

#define INVGPU_TILE_INSTANTIATE(T, N, TR, TC, PERM, STAGES, MINB) \
    template int invgpu::launch_tile_spd<T, N, TR, TC, PERM, STAGES, MINB>(invgpu::StridedIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_GJ_INSTANTIATE(T, N, ROWS, MINB) \
    template int invgpu::launch_gj<T, N, ROWS, invgpu::StridedIO<T>, MINB>(invgpu::StridedIO<T>, int, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *); \
    template int invgpu::launch_gj<T, N, ROWS, invgpu::PtrIO<T>, MINB>(invgpu::PtrIO<T>, int, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_GJR_INSTANTIATE(T, N, ROWS, MINB) \
    template int invgpu::launch_gj_roll<T, N, ROWS, invgpu::StridedIO<T>, MINB>(invgpu::StridedIO<T>, int, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *); \
    template int invgpu::launch_gj_roll<T, N, ROWS, invgpu::PtrIO<T>, MINB>(invgpu::PtrIO<T>, int, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_GJR2_INSTANTIATE(T, N, CW, MINB) \
    template int invgpu::launch_gj_roll2d<T, N, CW, invgpu::StridedIO<T>, MINB>(invgpu::StridedIO<T>, int, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *); \
    template int invgpu::launch_gj_roll2d<T, N, CW, invgpu::PtrIO<T>, MINB>(invgpu::PtrIO<T>, int, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_GJR2WS_INSTANTIATE(T, N, MINB) \
    template int invgpu::launch_gj_roll2d_ws<T, N, invgpu::StridedIO<T>, MINB>(invgpu::StridedIO<T>, int, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *); \
    template int invgpu::launch_gj_roll2d_ws<T, N, invgpu::PtrIO<T>, MINB>(invgpu::PtrIO<T>, int, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_GJC_INSTANTIATE(T, N, CL, WARPS, MINB) \
    template int invgpu::launch_gj_colsplit<T, N, CL, WARPS, MINB>(invgpu::StridedIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_GJT_INSTANTIATE(T, N, TR, TC, MINB) \
    template int invgpu::launch_gj_tile<T, N, TR, TC, invgpu::StridedIO<T>, MINB>(invgpu::StridedIO<T>, int, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *); \
    template int invgpu::launch_gj_tile<T, N, TR, TC, invgpu::PtrIO<T>, MINB>(invgpu::PtrIO<T>, int, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_ONESWEEP_INSTANTIATE(T, N, TR, TC, STAGE, MINB) \
    template int invgpu::launch_onesweep<T, N, TR, TC, STAGE, MINB>(invgpu::StridedIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_GP_THREAD_INSTANTIATE(T, N, WARPS, MINB) \
    template int invgpu::launch_gp_thread<T, N, WARPS, MINB>(invgpu::GpIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_THREAD_BULK_INSTANTIATE(T, N, WARPS, MINB) \
    template int invgpu::launch_spd_thread_bulk<T, N, WARPS, MINB>(invgpu::StridedIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_SPD8_TMA_INSTANTIATE(T, NBUF, MINB) \
    template int invgpu::launch_spd8_tma<T, NBUF, MINB, false>(invgpu::StridedIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *); \
    template int invgpu::launch_spd8_tma<T, NBUF, MINB, true>(invgpu::StridedIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_SWEEP_TMA_INSTANTIATE(V, T, N, TR, TC, UNROLL, MINB, DIRECT_OUT, INTERLEAVE) \
    template int invgpu::launch_sweep_tma<T, N, TR, TC, UNROLL, MINB, DIRECT_OUT, INTERLEAVE>(invgpu::StridedIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_SWEEP_PAD_INSTANTIATE(T, N, TR, TC, MINB) \
    template int invgpu::launch_sweep_pad<T, N, TR, TC, MINB>(invgpu::PadIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_SWEEP_GP_INSTANTIATE(V, T, N, TR, TC, UNROLL, MINB, BLK) \
    template int invgpu::launch_sweep_gp<T, N, TR, TC, UNROLL, MINB, BLK>(invgpu::GpIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_SWEEP_INSTANTIATE(V, T, N, TR, TC, UNROLL, MINB, BLK) \
    template int invgpu::launch_sweep<T, N, TR, TC, UNROLL, MINB, BLK>(invgpu::StridedIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#define INVGPU_TILE_INSTANTIATE_GP(T, N, TR, TC, MINB) \
    template int invgpu::launch_tile_gp<T, N, TR, TC, MINB>(invgpu::GpIO<T>, invgpu::i64, int *, cudaStream_t, invgpu::DeviceState *);
#endif

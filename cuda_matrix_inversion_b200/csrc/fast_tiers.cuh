// fast_tiers.cuh -- dispatch into the register-tiled tiers (tile_kernels.cuh, instantiated in the
// inst_*.cu units); returns INVGPU_NO_FAST_PATH when a shape has no specialised kernel, in which
// case the any-n shared-memory kernels of generic_smem.cuh serve it.
#pragma once

#include <type_traits>

#include "engine.cuh"
#include "generic_smem.cuh"
#include "tile_configs.h"
#include "tile_launch.cuh"

#define INVGPU_NO_FAST_PATH (-1000)
// orders up to this bound stay on the lane = row kernel (gj_kernels.cuh), measured faster there on B200:
// fp32 n = 16: 0.289 vs 0.147 of the HBM roofline, 32: 0.203 vs 0.074; fp64 32: 0.159 vs 0.124 (the tile kernel won
// against the first lane = row kernel, 0.138 vs 0.120, and loses against the lean one)
// fp32 n = 64: 0.108 (two rows per lane) vs 0.068
#define INVGPU_GJT_MIN_N(T) (sizeof(T) == 4 ? 64 : 32)
#ifndef INVGPU_SWEEP_MIN_N
#define INVGPU_SWEEP_MIN_N 16
#endif

namespace invgpu {

template <typename T>
static bool dense_aligned(const StridedIO<T> &io, int n) {
    return ((uintptr_t)io.in % 16 == 0) && ((uintptr_t)io.out % 16 == 0) &&
           (io.in_stride * (i64)sizeof(T)) % 16 == 0 && (io.out_stride * (i64)sizeof(T)) % 16 == 0 &&
           io.in_stride >= (i64)n * n && io.out_stride >= (i64)n * n;
}

// one `if` per instantiated configuration, generated from the X-macro lists
#define INVGPU_TILE_TRY(TT, N, TR, TC, PERM, STAGES_, MINB)                                         \
    if (std::is_same<T, TT>::value && n == N && STAGES == STAGES_)                                   \
        return launch_tile_spd<TT, N, TR, TC, PERM, STAGES_, MINB>(                                  \
            *reinterpret_cast<StridedIO<TT> *>(&io), batch, dInfo, st, ds);

#define INVGPU_ONESWEEP_TRY(TT, N, TR, TC, STAGE, MINB)                                             \
    if (std::is_same<T, TT>::value && n == N)                                                        \
        return launch_onesweep<TT, N, TR, TC, STAGE, MINB>(*reinterpret_cast<StridedIO<TT> *>(&io), batch, dInfo, st, ds);

#define INVGPU_SWEEP_TRY(V, TT, N, TR, TC, UNROLL, MINB, BLK)                                       \
    if (std::is_same<T, TT>::value && n == N && V == variant)                                        \
        return launch_sweep<TT, N, TR, TC, UNROLL, MINB, BLK>(*reinterpret_cast<StridedIO<TT> *>(&io), batch, dInfo, st, ds);

#define INVGPU_SWEEP_TMA_TRY(V, TT, N, TR, TC, UNROLL, MINB, DIRECT_OUT, INTERLEAVE)                 \
    if (std::is_same<T, TT>::value && n == N && V == variant) {                                      \
        const int rc_tma = launch_sweep_tma<TT, N, TR, TC, UNROLL, MINB, DIRECT_OUT, INTERLEAVE>(*reinterpret_cast<StridedIO<TT> *>(&io), batch, dInfo, st, ds); \
        if (rc_tma != INVGPU_TMA_UNAVAILABLE) return rc_tma;                                          \
        variant = 0;                                                                                 \
    }

// INVGPU_SPD_KERNEL=threesweep selects the three-sweep tile kernels for the sizes both families cover
static bool prefer_onesweep() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("INVGPU_SPD_KERNEL"); v = (e && !strcmp(e, "threesweep")) ? 0 : 1; }
    return v == 1;
}

template <typename T> static bool dense_aligned_any(const StridedIO<T> &io, int n) { return dense_aligned(io, n); }
template <typename T> static bool dense_aligned_any(const PtrIO<T> &, int) { return false; }

template <typename T, int STAGES>
static int fast_spd_dense(StridedIO<T> io, int n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    if (!dense_aligned(io, n)) return INVGPU_NO_FAST_PATH;
    if (STAGES == SPD_INVERSE && prefer_onesweep()) {
        // n >= INVGPU_SWEEP_MIN_N: the look-ahead sweep kernel (sweep_kernels.cuh); below: the unrolled one-sweep
        // kernel.  INVGPU_SWEEP_VARIANT=V picks another instantiated thread grid, INVGPU_SPD_KERNEL=onesweep
        // keeps the older kernel for every size it covers (experiments, tools/kbench.py).
        static int env_variant = -1, old = -1;
        if (env_variant < 0) { const char *e = getenv("INVGPU_SWEEP_VARIANT"); env_variant = e ? atoi(e) : 0; }
        int variant = env_variant;
        if (old < 0) { const char *e = getenv("INVGPU_SPD_KERNEL"); old = (e && !strcmp(e, "onesweep")) ? 1 : 0; }
        if (!old && n == 8 && variant != 9) {                      // one thread per matrix + TMA tile I/O
#define INVGPU_SPD8_TRY(TT, NBUF, MINB)                                                              \
            if (std::is_same<T, TT>::value) {                                                         \
                const int rc8 = launch_spd8_tma<TT, NBUF, MINB, false>(*reinterpret_cast<StridedIO<TT> *>(&io), batch, dInfo, st, ds); \
                if (rc8 != INVGPU_TMA_UNAVAILABLE) return rc8;                                         \
            }
            INVGPU_SPD8_TMA_ALL(INVGPU_SPD8_TRY)
        }
        if (!old && variant != 9) {                                // one thread per matrix + per-lane bulk copies
#define INVGPU_THREAD_BULK_TRY(TT, N, WARPS, MINB)                                                   \
            if (std::is_same<T, TT>::value && n == N)                                                 \
                return launch_spd_thread_bulk<TT, N, WARPS, MINB>(*reinterpret_cast<StridedIO<TT> *>(&io), batch, dInfo, st, ds);
            INVGPU_THREAD_BULK_ALL(INVGPU_THREAD_BULK_TRY)
        }
        if (!old && n >= INVGPU_SWEEP_MIN_N) {
            if (variant == 9) variant = 0;                         // 9 = the default grids without TMA tile I/O
            else { INVGPU_SWEEP_TMA_ALL(INVGPU_SWEEP_TMA_TRY) }    // falls through when the batch is not TMA-describable
            INVGPU_SWEEP_ALL(INVGPU_SWEEP_TRY)
        }
        INVGPU_ONESWEEP_ALL(INVGPU_ONESWEEP_TRY)
        if (!old) { variant = 0; INVGPU_SWEEP_ALL(INVGPU_SWEEP_TRY) }
    }
    INVGPU_TILE_SPD_ALL(INVGPU_TILE_TRY)
    return INVGPU_NO_FAST_PATH;
}

// smallest padded sweep tier >= n that is instantiated for T (0: none)
#define INVGPU_SWEEP_PAD_PICK(TT, N, TR, TC, MINB) \
    if (std::is_same<T, TT>::value && n <= N && (best == 0 || N < best)) best = N;
template <typename T>
static int padded_tier_for(int n) {
    int best = 0;
    INVGPU_SWEEP_PAD_ALL(INVGPU_SWEEP_PAD_PICK)
    return best;
}
template <typename T>
static int fast_padded(PadIO<T> io, int tier_n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds);

// Inverse of a batch the aligned dense kernels cannot take (pointer arrays, odd orders, unaligned strides):
// the padded sweep tier -- blockdiag(A, I) in the register tile, bounds-checked scalar I/O.
template <typename T>
static int fast_spd_padded(PadIO<T> pio, int n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    static int off = -1;                              // INVGPU_SPD_KERNEL=generic keeps such batches on the shared-memory tier
    if (off < 0) { const char *e = getenv("INVGPU_SPD_KERNEL"); off = (e && !strcmp(e, "generic")) ? 1 : 0; }
    const int tier = off ? 0 : padded_tier_for<T>(n);
    if (!tier) return INVGPU_NO_FAST_PATH;
    pio.n = n;
    return fast_padded<T>(pio, tier, batch, dInfo, st, ds);
}

template <typename T, typename IO, int STAGES>
struct FastSpd {   // pointer-array batches: inverse on the padded tiers, staged calls on the generic tier
    static int run(IO io, int n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
        if (STAGES != SPD_INVERSE) return INVGPU_NO_FAST_PATH;
        PadIO<T> pio;
        pio.in_ptrs = io.in; pio.out_ptrs = io.out;
        return fast_spd_padded<T>(pio, n, batch, dInfo, st, ds);
    }
};
template <typename T, int STAGES>
struct FastSpd<T, StridedIO<T>, STAGES> {
    static int run(StridedIO<T> io, int n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
        const int rc = fast_spd_dense<T, STAGES>(io, n, batch, dInfo, st, ds);
        if (rc != INVGPU_NO_FAST_PATH || STAGES != SPD_INVERSE) return rc;
        PadIO<T> pio;
        pio.in_base = io.in; pio.out_base = io.out; pio.in_stride = io.in_stride; pio.out_stride = io.out_stride;
        return fast_spd_padded<T>(pio, n, batch, dInfo, st, ds);
    }
};

template <typename T, typename IO, int STAGES>
static int fast_spd(IO io, int n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    return FastSpd<T, IO, STAGES>::run(io, n, batch, dInfo, st, ds);
}

// smallest instantiated padded order >= n wins (the lists are ascending in N)
#define INVGPU_GJ_TRY(TT, N, ROWS, MINB)                                                            \
    if (std::is_same<T, TT>::value && n <= N)                                                        \
        return launch_gj<TT, N, ROWS, typename IOCast<IO, TT>::type, MINB>(                          \
            *reinterpret_cast<typename IOCast<IO, TT>::type *>(&io), n, batch, dInfo, st, ds);

template <typename IO, typename TT> struct IOCast;
template <typename T, typename TT> struct IOCast<StridedIO<T>, TT> { typedef StridedIO<TT> type; };
template <typename T, typename TT> struct IOCast<PtrIO<T>, TT> { typedef PtrIO<TT> type; };

#define INVGPU_GJR_TRY(TT, N, ROWS, MINB)                                                           \
    if (std::is_same<T, TT>::value && n <= N)                                                        \
        return launch_gj_roll<TT, N, ROWS, typename IOCast<IO, TT>::type, MINB>(                     \
            *reinterpret_cast<typename IOCast<IO, TT>::type *>(&io), n, batch, dInfo, st, ds);

#define INVGPU_GJR2_TRY(TT, N, CW, MINB)                                                            \
    if (std::is_same<T, TT>::value && n <= N)                                                        \
        return launch_gj_roll2d<TT, N, CW, typename IOCast<IO, TT>::type, MINB>(                         \
            *reinterpret_cast<typename IOCast<IO, TT>::type *>(&io), n, batch, dInfo, st, ds);

#define INVGPU_GJR2WS_TRY(TT, N, MINB)                                                              \
    if (std::is_same<T, TT>::value && n <= N && gj_roll2d_ws_enabled())                              \
        return launch_gj_roll2d_ws<TT, N, typename IOCast<IO, TT>::type, MINB>(                      \
            *reinterpret_cast<typename IOCast<IO, TT>::type *>(&io), n, batch, dInfo, st, ds);

// 2-D tile Gauss-Jordan: smallest instantiated padded order >= n wins (ascending lists)
#define INVGPU_GJT_TRY(TT, N, TR, TC, MINB)                                                         \
    if (std::is_same<T, TT>::value && n <= N)                                                        \
        return launch_gj_tile<TT, N, TR, TC, typename IOCast<IO, TT>::type, MINB>(                   \
            *reinterpret_cast<typename IOCast<IO, TT>::type *>(&io), n, batch, dInfo, st, ds);

template <typename T, typename IO>
static int fast_general(IO io, int n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    static int which = -1;                            // INVGPU_GJ_KERNEL=rowlane | generic | colsplit | tile: other tiers (experiments)
    if (which < 0) {
        const char *e = getenv("INVGPU_GJ_KERNEL");
        which = (e && !strcmp(e, "rowlane")) ? 1 : (e && !strcmp(e, "generic")) ? 2 : (e && !strcmp(e, "colsplit")) ? 3 : (e && !strcmp(e, "tile")) ? 4 : 0;
    }
    if (which == 2) return INVGPU_NO_FAST_PATH;
    const bool use_roll = which == 0 && n <= (sizeof(T) == 4 ? 64 : 32);
    if constexpr (std::is_same<IO, StridedIO<T>>::value) {        // dense batches of order exactly 8: one thread per matrix + TMA
        if (which == 0 && n == 8 && dense_aligned(io, 8)) {
#define INVGPU_GJ8_TRY(TT, NBUF, MINB)                                                               \
            if (std::is_same<T, TT>::value) {                                                         \
                const int rc8 = launch_spd8_tma<TT, NBUF, MINB, true>(*reinterpret_cast<StridedIO<TT> *>(&io), batch, dInfo, st, ds); \
                if (rc8 != INVGPU_TMA_UNAVAILABLE) return rc8;                                         \
            }
            INVGPU_SPD8_TMA_ALL(INVGPU_GJ8_TRY)
        }
        // dense batches of order exactly 16 / 32: column-split lanes, matrix in registers (INVGPU_GJ_KERNEL=colsplit only:
        // the lean lane = row kernel is faster, tile_configs.h)
#define INVGPU_GJC_TRY(TT, N, CL, WARPS, MINB)                                                       \
        if ((which == 3 || (which == 0 && INVGPU_GJC_DEFAULT(TT))) && std::is_same<T, TT>::value && n == N && dense_aligned(io, N)) \
            return launch_gj_colsplit<TT, N, CL, WARPS, MINB>(*reinterpret_cast<StridedIO<TT> *>(&io), batch, dInfo, st, ds);
        INVGPU_GJC_ALL(INVGPU_GJC_TRY)
    }
    if (use_roll) { INVGPU_GJR_ALL(INVGPU_GJR_TRY) }    // rolled lane = row kernel: every n <= 64 that the n = 8 TMA kernel did not take
    if (which == 0 && sizeof(T) == 4 && n > 64) { INVGPU_GJR2WS_ALL(INVGPU_GJR2WS_TRY) INVGPU_GJR2_ALL(INVGPU_GJR2_TRY) }   // one CTA per matrix, rolled (fp32 65 .. 128)
    if (((which == 0 || which == 3) && n > INVGPU_GJT_MIN_N(T)) || (which == 4 && n > 16)) { INVGPU_GJT_ALL(INVGPU_GJT_TRY) }
    INVGPU_GJ_ALL(INVGPU_GJ_TRY)
    return INVGPU_NO_FAST_PATH;
}

#define INVGPU_TILE_TRY_GP(TT, N, TR, TC, MINB)                                                     \
    if (std::is_same<T, TT>::value && n == N)                                                        \
        return launch_tile_gp<TT, N, TR, TC, MINB>(*reinterpret_cast<GpIO<TT> *>(&io), batch, dInfo, st, ds);

#define INVGPU_SWEEP_TRY_GP(V, TT, N, TR, TC, UNROLL, MINB, BLK)                                    \
    if (std::is_same<T, TT>::value && n == N && V == variant)                                        \
        return launch_sweep_gp<TT, N, TR, TC, UNROLL, MINB, BLK>(*reinterpret_cast<GpIO<TT> *>(&io), batch, dInfo, st, ds);

template <typename T>
static int fast_gp(GpIO<T> io, int n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    const uintptr_t all = (uintptr_t)io.a | (uintptr_t)io.b | (uintptr_t)io.c | (uintptr_t)io.d;
    if (all % 16 != 0 || ((size_t)n * sizeof(T)) % 16 != 0) return INVGPU_NO_FAST_PATH;
    static int old = -1, no_tc = -1;                  // INVGPU_GP_KERNEL=tile keeps the three-phase tile kernels, =sweep the CUDA-core sweep
    if (old < 0) { const char *e = getenv("INVGPU_GP_KERNEL"); old = (e && !strcmp(e, "tile")) ? 1 : 0; no_tc = (e && !strcmp(e, "sweep")) ? 1 : 0; }
    if constexpr (std::is_same<T, float>::value) {    // n = 128 fp32: blocked Cholesky with the tcgen05 trailing update (tc_kernels.cuh)
        if (n == 128 && !old && !no_tc) return launch_tc_gp128(io, batch, dInfo, st, ds);
    }
    static int variant = -1;                          // INVGPU_SWEEP_VARIANT=V: another instantiated configuration
    if (variant < 0) { const char *e = getenv("INVGPU_SWEEP_VARIANT"); variant = e ? atoi(e) : 0; }
#define INVGPU_GP_THREAD_TRY(TT, N, WARPS, MINB)                                                    \
    if (std::is_same<T, TT>::value && n == N)                                                        \
        return launch_gp_thread<TT, N, WARPS, MINB>(*reinterpret_cast<GpIO<TT> *>(&io), batch, dInfo, st, ds);
    if (!old) { INVGPU_GP_THREAD_ALL(INVGPU_GP_THREAD_TRY) }
    if (!old) { INVGPU_SWEEP_GP_ALL(INVGPU_SWEEP_TRY_GP) }
    INVGPU_TILE_GP_ALL(INVGPU_TILE_TRY_GP)
    return INVGPU_NO_FAST_PATH;
}

// mixed-dimension batches: the padded sweep tier of order exactly N, if instantiated for T
#define INVGPU_SWEEP_PAD_TRY(TT, N, TR, TC, MINB)                                                   \
    if (std::is_same<T, TT>::value && tier_n == N)                                                   \
        return launch_sweep_pad<TT, N, TR, TC, MINB>(*reinterpret_cast<PadIO<TT> *>(&io), batch, dInfo, st, ds);

template <typename T>
static int fast_padded(PadIO<T> io, int tier_n, i64 batch, int *dInfo, cudaStream_t st, DeviceState *ds) {
    INVGPU_SWEEP_PAD_ALL(INVGPU_SWEEP_PAD_TRY)
    return INVGPU_NO_FAST_PATH;
}

// op: 0 = spd inverse, 1 = general inverse, 2 = gp; which tier serves a dense aligned batch
#define INVGPU_TILE_NAME(TT, N, TR, TC, PERM, STAGES_, MINB) \
    if (op == 0 && n == N && dtype_bytes == (int)sizeof(TT) && STAGES_ == 7) return TR * TC <= 32 ? "warp-tile" : "cta-tile";
#define INVGPU_TILE_NAME_GP(TT, N, TR, TC, MINB) \
    if (op == 2 && n == N && dtype_bytes == (int)sizeof(TT)) return TR * TC <= 32 ? "warp-tile" : "cta-tile";
#define INVGPU_GJC_NAME(TT, N, CL, WARPS, MINB) \
    if (op == 1 && n == N && dtype_bytes == (int)sizeof(TT) && INVGPU_GJC_DEFAULT(TT)) return "gj-colsplit";
#define INVGPU_GJ8_NAME(TT, NBUF, MINB) \
    if (op == 1 && n == 8 && dtype_bytes == (int)sizeof(TT)) return "thread-tma";
#define INVGPU_GJT_NAME(TT, N, TR, TC, MINB) \
    if (op == 1 && n > INVGPU_GJT_MIN_N(TT) && n <= N && dtype_bytes == (int)sizeof(TT)) return TR * TC <= 32 ? "gj-tile-warp" : "gj-tile-cta";
#define INVGPU_GJ_NAME(TT, N, ROWS, MINB) \
    if (op == 1 && n <= N && dtype_bytes == (int)sizeof(TT)) return "warp-rowlane";
#define INVGPU_GJR_NAME(TT, N, ROWS, MINB) \
    if (op == 1 && n <= N && dtype_bytes == (int)sizeof(TT)) return "warp-rowlane-rolled";
#define INVGPU_GJR2_NAME(TT, N, CW, MINB) \
    if (op == 1 && n > 64 && n <= N && dtype_bytes == (int)sizeof(TT)) return "cta-roll2d";
#define INVGPU_THREAD_BULK_NAME(TT, N, WARPS, MINB) \
    if (op == 0 && n == N && dtype_bytes == (int)sizeof(TT)) return "thread-bulk";
#define INVGPU_SPD8_NAME(TT, NBUF, MINB) \
    if (op == 0 && n == 8 && dtype_bytes == (int)sizeof(TT)) return "thread-tma";
#define INVGPU_SWEEP_NAME(V, TT, N, TR, TC, UNROLL, MINB, BLK) \
    if (op == 0 && V == 0 && n == N && n >= INVGPU_SWEEP_MIN_N && dtype_bytes == (int)sizeof(TT)) return TR * TC <= 32 ? "sweep-warp" : "sweep-cta";
#define INVGPU_GP_THREAD_NAME(TT, N, WARPS, MINB) \
    if (op == 2 && n == N && dtype_bytes == (int)sizeof(TT)) return "thread-bulk";
#define INVGPU_SWEEP_GP_NAME(V, TT, N, TR, TC, UNROLL, MINB, BLK) \
    if (op == 2 && V == 0 && n == N && dtype_bytes == (int)sizeof(TT)) return TR * TC <= 32 ? "sweep-warp" : "sweep-cta";
static const char *fast_tier_name(int op, int n, int dtype_bytes) {
    if (op == 2 && n == 128 && dtype_bytes == 4) return "tcgen05-blocked";
    INVGPU_SPD8_TMA_ALL(INVGPU_SPD8_NAME)
    INVGPU_THREAD_BULK_ALL(INVGPU_THREAD_BULK_NAME)
    INVGPU_SWEEP_ALL(INVGPU_SWEEP_NAME)
    INVGPU_GP_THREAD_ALL(INVGPU_GP_THREAD_NAME)
    INVGPU_SWEEP_GP_ALL(INVGPU_SWEEP_GP_NAME)
    INVGPU_SPD8_TMA_ALL(INVGPU_GJ8_NAME)
    INVGPU_GJC_ALL(INVGPU_GJC_NAME)
    INVGPU_GJR_ALL(INVGPU_GJR_NAME)
    INVGPU_GJR2_ALL(INVGPU_GJR2_NAME)
    INVGPU_GJT_ALL(INVGPU_GJT_NAME)
    INVGPU_GJ_ALL(INVGPU_GJ_NAME)
    INVGPU_TILE_SPD_ALL(INVGPU_TILE_NAME)
    INVGPU_TILE_GP_ALL(INVGPU_TILE_NAME_GP)
    if (op == 0 && ((dtype_bytes == 4 && padded_tier_for<float>(n)) || (dtype_bytes == 8 && padded_tier_for<double>(n)))) return "sweep-padded";
    return "generic";
}

}  // namespace invgpu

// tile_configs.h -- the register-tiled kernels that are instantiated (one translation unit per
// group, compiled in parallel), as X-macro lists.
//   SPD : X(T, N, TR, TC, PERM, STAGES, MINB)     GP : X(T, N, TR, TC, MINB)
//   TR x TC = thread grid per matrix, MINB = resident CTAs per SM the register allocator must allow
#pragma once

// INVGPU_LAB (make lab=1): also build the kernel generations and configurations that were measured and NOT made the
// default -- the three-sweep tile inverse, the unrolled one-sweep kernels above n = 8, sweep variants 1 / 3 / 4 (other thread
// grids, 2x2 block pivots), TMA variants 6 / 7, the unrolled lane = row and the column-split Gauss-Jordan kernels.  They stay
// selectable through the INVGPU_* environment knobs of INTEGRATION.md and keep their parity tests (skipped without the flag);
// a default build ships only what the dispatcher can reach.
#ifndef INVGPU_LAB
#define INVGPU_LAB 0
#endif
#if INVGPU_LAB
#define INVGPU_LAB_ONLY(...) __VA_ARGS__
#else
#define INVGPU_LAB_ONLY(...)
#endif

#ifndef INVGPU_F32_N32_MINB
#define INVGPU_F32_N32_MINB 2
#endif
#ifndef INVGPU_F32_N32_TR
#define INVGPU_F32_N32_TR 4
#define INVGPU_F32_N32_TC 2
#endif

// fp32 SPD inverse, warp tiers
#define INVGPU_TILE_SPD_F32_INV(X) INVGPU_LAB_ONLY(             \
    X(float, 8, 1, 1, true, 7, 4)                               \
    X(float, 16, 2, 2, true, 7, 4)                              \
    X(float, 32, INVGPU_F32_N32_TR, INVGPU_F32_N32_TC, true, 7, INVGPU_F32_N32_MINB)            \
    X(float, 64, 8, 4, true, 7, 3))
// fp32 SPD inverse, CTA tier
#define INVGPU_TILE_SPD_F32_INV_CTA(X) INVGPU_LAB_ONLY(         \
    X(float, 128, 16, 16, true, 7, 2))

// fp64 SPD inverse
#define INVGPU_TILE_SPD_F64_INV(X) INVGPU_LAB_ONLY(             \
    X(double, 8, 1, 1, true, 7, 2)                              \
    X(double, 16, 2, 2, true, 7, 2)                             \
    X(double, 32, 4, 4, true, 7, 2)                             \
    X(double, 64, 8, 8, true, 7, 4))
#define INVGPU_TILE_SPD_F64_INV_CTA(X) INVGPU_LAB_ONLY(         \
    X(double, 128, 16, 16, true, 7, 1))

// fused GP mean / variance
#define INVGPU_TILE_GP_F32(X)                                   \
    X(float, 32, 4, 2, 3)                                       \
    X(float, 64, 8, 4, 3)                                       \
    X(float, 128, 16, 16, 2)
#define INVGPU_TILE_GP_F64(X)                                   \
    X(double, 16, 2, 2, 2)                                      \
    X(double, 32, 4, 4, 2)                                      \
    X(double, 64, 8, 8, 4)                                      \
    X(double, 128, 16, 16, 1)

#define INVGPU_TILE_SPD_ALL(X)                                  \
    INVGPU_TILE_SPD_F32_INV(X) INVGPU_TILE_SPD_F32_INV_CTA(X)   \
    INVGPU_TILE_SPD_F64_INV(X) INVGPU_TILE_SPD_F64_INV_CTA(X)
#define INVGPU_TILE_GP_ALL(X) INVGPU_TILE_GP_F32(X) INVGPU_TILE_GP_F64(X)

// general inverse, lane = row Gauss-Jordan (gj_kernels.cuh):  X(T, N, ROWS, MINB); N = padded order
// Measured on B200 (fp32, dense, fraction of the HBM roofline; ROWS x CTAs per SM):
//   n = 16: 1 x 4: 0.203, 1 x 6: 0.226, 2 x 4: 0.238, 2 x 6: 0.255 (default)      column-split lanes (INVGPU_GJC): 0.227
//   n = 32: 1 x 4: 0.147, 1 x 6: 0.161 (default), 2 x 4: 0.154, 2 x 5: 0.134 (spills)   column-split lanes: 0.122
// (before the exact-order load / store path; with it the defaults reach 0.289 at n = 16 and 0.203 at n = 32)
// 8 CTAs per SM (64 registers, spills): n = 16: 0.252, n = 32: 0.206 -- not worth it
// fp64: n = 16: 1 x 4: 0.284, 2 x 4: 0.311 (default; column-split lanes 0.245); n = 32: 1 x 3: 0.151, 1 x 4: 0.159 (default; 4 x 4 tile kernel 0.124)
#ifndef INVGPU_GJ32_MINB
#define INVGPU_GJ32_MINB 6
#endif
#ifndef INVGPU_GJ16_MINB
#define INVGPU_GJ16_MINB 6
#endif
#ifndef INVGPU_GJ32_ROWS
#define INVGPU_GJ32_ROWS 1
#endif
#ifndef INVGPU_GJ16_ROWS
#define INVGPU_GJ16_ROWS 2
#endif
// n = 64 fp32: two rows per lane, one warp per matrix, 250 registers: 0.108 vs 0.068 for the 8 x 4 tile kernel (n = 48: 0.054 vs 0.045)
#define INVGPU_GJ64_F32(X) X(float, 64, 2, 2)
#define INVGPU_GJ_F32(X) INVGPU_LAB_ONLY(X(float, 8, 1, 4) X(float, 16, INVGPU_GJ16_ROWS, INVGPU_GJ16_MINB) X(float, 32, INVGPU_GJ32_ROWS, INVGPU_GJ32_MINB) INVGPU_GJ64_F32(X))
#ifndef INVGPU_GJ16_ROWS_F64
#define INVGPU_GJ16_ROWS_F64 2
#endif
#ifndef INVGPU_GJ32_MINB_F64
#define INVGPU_GJ32_MINB_F64 4
#endif
#define INVGPU_GJ_F64(X) INVGPU_LAB_ONLY(X(double, 8, 1, 4) X(double, 16, INVGPU_GJ16_ROWS_F64, 4) X(double, 32, 1, INVGPU_GJ32_MINB_F64))
#define INVGPU_GJ_ALL(X) INVGPU_GJ_F32(X) INVGPU_GJ_F64(X)

// lane = row Gauss-Jordan with the ROLLED pivot loop (gj_roll_kernels.cuh: rotating register window, deferred row
// scaling):  X(T, N, ROWS, MINB); the smallest N >= n serves an order n.  The default general-inverse tier for n <= 64.
#ifndef INVGPU_GJR64_MINB
#define INVGPU_GJR64_MINB 3      // 168 registers, 70 B of spills: 0.202 of the roofline vs 0.197 with 2 CTAs per SM (198 registers)
#endif
#ifndef INVGPU_GJR32_ROWS
#define INVGPU_GJR32_ROWS 1
#endif
#ifndef INVGPU_GJR32_MINB
#define INVGPU_GJR32_MINB 6
#endif
#ifndef INVGPU_GJR16_ROWS
#define INVGPU_GJR16_ROWS 2
#endif
#define INVGPU_GJR_F32(X) X(float, 8, 1, 4) X(float, 16, INVGPU_GJR16_ROWS, 6) X(float, 32, INVGPU_GJR32_ROWS, INVGPU_GJR32_MINB) X(float, 64, 2, INVGPU_GJR64_MINB)
#define INVGPU_GJR_F64(X) X(double, 8, 1, 4) X(double, 16, 2, 4) X(double, 32, 1, 4)
#define INVGPU_GJR_ALL(X) INVGPU_GJR_F32(X) INVGPU_GJR_F64(X)

// one CTA per matrix, 32 lanes x N / CW warps, rolled pivot loop (gj_roll2d_kernels.cuh), fp32 64 < n <= 128:  X(T, N, CW, MINB)
// measured on B200, 16 384 x 128x128: CW 32 / 3 CTAs per SM 4.21 ms (default; 4.14 ms with the two-reduction pivot search), CW 32 / 2 CTAs 4.25 ms, CW 16 (8 warps) / 2 CTAs 4.87 ms;
// the tile kernel it replaces: 7.08 ms
#ifndef INVGPU_GJR2_CW
#define INVGPU_GJR2_CW 32
#endif
#ifndef INVGPU_GJR2_MINB
#define INVGPU_GJR2_MINB 3
#endif
#define INVGPU_GJR2_F32(X) X(float, 128, INVGPU_GJR2_CW, INVGPU_GJR2_MINB)
#define INVGPU_GJR2_ALL(X) INVGPU_GJR2_F32(X)
// the warp-specialised form (4 FMA warps + 1 pivot warp, setmaxnreg register hand-over, INVGPU_GJR2_WS=1):  X(T, N, MINB)
// measured 4.60 ms against 4.21-4.26 ms of the four-warp kernel (profiles/r2_gj_roll2d_128_summary.md): lab builds only
#ifndef INVGPU_GJR2WS_MINB
#define INVGPU_GJR2WS_MINB 2
#endif
#define INVGPU_GJR2WS_F32(X) INVGPU_LAB_ONLY(X(float, 128, INVGPU_GJR2WS_MINB))
#define INVGPU_GJR2WS_ALL(X) INVGPU_GJR2WS_F32(X)

// SPD inverse, one-sweep Cholesky (onesweep_kernels.cuh), warp tiers:  X(T, N, TR, TC, STAGE, MINB)
#ifndef INVGPU_OS_F32_N32_MINB
#define INVGPU_OS_F32_N32_MINB 5
#endif
#ifndef INVGPU_OS_F32_N32_TR
#define INVGPU_OS_F32_N32_TR 4
#define INVGPU_OS_F32_N32_TC 4
#endif
// (n = 64 stays on the three-sweep kernel: without compile-time pruning of the shrinking trailing
//  matrix the one-sweep form does more FMAs, and at n = 64 the kernel is no longer latency-bound)
#define INVGPU_ONESWEEP_F32(X) X(float, 8, 1, 1, true, 4) INVGPU_LAB_ONLY(X(float, 16, 2, 2, false, 4) X(float, 32, INVGPU_OS_F32_N32_TR, INVGPU_OS_F32_N32_TC, false, INVGPU_OS_F32_N32_MINB))
#define INVGPU_ONESWEEP_F64(X) X(double, 8, 1, 1, true, 2) INVGPU_LAB_ONLY(X(double, 16, 2, 2, false, 2) X(double, 32, 4, 4, false, 2))
#define INVGPU_ONESWEEP_ALL(X) INVGPU_ONESWEEP_F32(X) INVGPU_ONESWEEP_F64(X)

// SPD inverse, look-ahead sweep kernel (sweep_kernels.cuh):  X(V, T, N, TR, TC, UNROLL, MINB, BLK)
// V = variant number; V == 0 is what the dispatcher uses, the others are reachable with
// INVGPU_SWEEP_VARIANT=V (tools/kbench.py experiments).  BLK = 1: scalar pivots, 2: 2x2 block pivots.
#define INVGPU_SWEEP_F32(X)                                                                     \
    X(0, float, 16, 2, 2, false, 4, 1)                                                          \
    X(0, float, 32, 2, 4, false, 3, 1) INVGPU_LAB_ONLY(X(1, float, 32, 4, 2, false, 3, 1) X(4, float, 32, 2, 4, false, 3, 2)) \
    X(0, float, 64, 4, 4, false, 2, 1) INVGPU_LAB_ONLY(X(3, float, 64, 8, 4, false, 3, 1) X(4, float, 64, 4, 4, false, 2, 2)) \
    X(0, float, 128, 8, 8, false, 2, 1) INVGPU_LAB_ONLY(X(3, float, 128, 16, 8, false, 3, 1) X(4, float, 128, 8, 8, false, 2, 2))
#define INVGPU_SWEEP_F64(X)                                                                     \
    X(0, double, 16, 2, 2, false, 2, 1) X(0, double, 32, 4, 4, false, 2, 1) X(0, double, 64, 8, 8, false, 4, 1) X(0, double, 128, 16, 16, false, 1, 1) \
    INVGPU_LAB_ONLY(X(4, double, 64, 8, 8, false, 4, 2) X(4, double, 128, 16, 16, false, 1, 2))
// Measured on B200, fraction of the HBM roofline (tools/kbench.py, gpurun_out/o_kbench.log), BLK = 1 vs BLK = 2:
//   fp32 n = 32: 0.51 vs 0.44   64: 0.28 vs 0.25   128: 0.18 vs 0.165   fp64 64: 0.229 vs 0.232   128: 0.112 vs 0.116
// BLK = 0 (no look-ahead: publish -> barrier -> update through one rolled body per range, half the code):
//   fp32 n = 32: 0.38 (2x4 lanes), 0.43 (4x4 lanes) vs 0.51;  n = 64: 0.23 vs 0.28  -- the look-ahead is worth 25-35 %.
// Tile size (fp32): 16 x 16 tiles with the strictly-upper blocks never materialised (160 accumulator registers, 255 in all)
// beat 8 x 16 tiles: n = 64 on 4 x 4 lanes 0.359 vs 0.284 (8 x 4 lanes), n = 128 on 8 x 8 threads 0.227 vs 0.182 (16 x 8):
// 5 FMAs per broadcast operand word instead of 4 -- the kernels are bound by shared-memory bandwidth.  (n = 32 on 2 x 2
// lanes does not gain: 0.47.)
// Occupancy (fp32, 8 x 16 tiles, CTAs per SM 2 / 3 / 4): n = 64: 0.275 / 0.283 / 0.168 (spills), n = 128: 0.171 / 0.182 / 0.100 (spills).
// n = 32 fp32 on 2 x 2 lanes (16 x 16 tiles, strictly-upper blocks pruned, 254 registers, 2 CTAs per SM): 0.47 vs 0.51.
// The 2x2 block pivots halve barriers and dependency chains but double the live operand registers (x1, x2, y1, y2)
// and the shared-memory bytes per step; the kernels are not chain-bound enough for that to pay.  Thread grids:
// wide tiles (TR > TC) win by 3-7 % at n = 64 / 128 (row owners of a pivot sit in one quarter-warp).
#define INVGPU_SWEEP_ALL(X) INVGPU_SWEEP_F32(X) INVGPU_SWEEP_F64(X)

// the same kernel with TMA tile I/O (sweep_spd_tma_kernel; columns of exactly 128 bytes):  X(V, T, N, TR, TC, UNROLL, MINB, DIRECT_OUT, INTERLEAVE)
// Measured on B200 (2^19 matrices, fraction of the HBM roofline): direct global access 0.512 (sweep_spd_kernel,
// INVGPU_SWEEP_VARIANT=9 skips the TMA kernels), TMA in + out 0.508 (6), TMA in with prefetch + direct stores 0.507 (7), TMA in + out with INTERLEAVED
// lanes 0.577 (0, the default; direct stores with interleaved lanes: 0.448, 4 x 2 lanes interleaved: 0.456; FULL symmetric
// storage -- pivot column alone as the broadcast vector, no row part, no mirrored stores, a third more FMAs: 0.271): the kernel is bound by shared-memory bandwidth, and with the global side on the TMA
// unit the lane map can be chosen for the publish stores alone.
#ifndef INVGPU_TMA_N32_MINB
#define INVGPU_TMA_N32_MINB 3
#endif
#define INVGPU_SWEEP_TMA_F32(X) X(0, float, 32, 2, 4, false, INVGPU_TMA_N32_MINB, false, true) INVGPU_LAB_ONLY(X(6, float, 32, 2, 4, false, 3, false, false) X(7, float, 32, 2, 4, false, 3, true, false))
// fp64 with interleaved lanes and per-lane bulk-copy tile I/O (padded slots instead of swizzle; tried, not kept):
// n = 32: 0.382 vs 0.418 direct, n = 16: 0.553 vs 0.574 -- at half-rate DFMA and 192-204 registers it does not pay.
#define INVGPU_SWEEP_TMA_F64(X)
#define INVGPU_SWEEP_TMA_ALL(X) INVGPU_SWEEP_TMA_F32(X) INVGPU_SWEEP_TMA_F64(X)

// fused GP mean / variance on the sweep machinery (sweep_gp_kernel):  X(V, T, N, TR, TC, UNROLL, MINB, BLK)
// Measured on B200 against the three-phase tile kernels above (fraction of the HBM roofline, sweep vs tile):
// fp32 n = 32: 0.45 vs 0.57, 64: 0.23 vs 0.39, 128: 0.118 vs 0.092; fp64 32/64/128: equal within 5 %.
// Only the CTA tier gains (one barrier per pivot instead of the rolled potrf's per-pivot chain), so only
// that one is dispatched (8 x 8 threads with 16 x 16 tiles: 0.183, with 2x2 block pivots 0.177; 16 x 8 threads with 2x2 block
// pivots: 0.123; 8 x 16 scalar pivots: 0.118; the sweep on 4 x 4 lanes at n = 64: 0.325 vs 0.386 for the tile kernel); the warp tiers stay on the fully unrolled tile kernels with exact static pruning.
#define INVGPU_SWEEP_GP_F32(X) X(0, float, 128, 8, 8, false, 2, 1) INVGPU_LAB_ONLY(X(3, float, 128, 16, 8, false, 3, 2))
#define INVGPU_SWEEP_GP_F64(X)
#define INVGPU_SWEEP_GP_ALL(X) INVGPU_SWEEP_GP_F32(X) INVGPU_SWEEP_GP_F64(X)

// mixed-dimension batches on the sweep kernel with the padded IO policy:  X(T, N, TR, TC, MINB); tiers in ascending N
// (intermediate tiers 24 / 48 / 96 / 192 on square grids: their strictly-upper 4x4 blocks are never materialised, so a
//  12 x 12 tile costs 96 registers; they cut the (N / n)^3 padding waste of a tier from 3.0 to 1.7 on average)
#ifndef INVGPU_PAD256_TR
#define INVGPU_PAD256_TR 16
#define INVGPU_PAD256_TC 16
#define INVGPU_PAD256_MINB 1
#endif
// (192 tier at 2 CTAs per SM -- 128 registers, 2.6 KB of spills: the 500 k mixed batch takes 27.9 ms instead of 24.0 with the tiers
//  serialised; it stays at one CTA per SM, 201 registers)
#ifndef INVGPU_PAD192_TR
#define INVGPU_PAD192_TR 16
#define INVGPU_PAD192_TC 16
#define INVGPU_PAD192_MINB 1
#endif
#ifndef INVGPU_PAD128_TR
#define INVGPU_PAD128_TR 8
#define INVGPU_PAD128_TC 8
#define INVGPU_PAD128_MINB 2
#endif
#define INVGPU_SWEEP_PAD_F32(X) X(float, 16, 2, 2, 4) X(float, 24, 2, 2, 3) X(float, 32, 2, 4, 3) X(float, 48, 4, 4, 3) X(float, 64, 4, 4, 2) \
    X(float, 96, 8, 8, 4) X(float, 128, INVGPU_PAD128_TR, INVGPU_PAD128_TC, INVGPU_PAD128_MINB) \
    X(float, 192, INVGPU_PAD192_TR, INVGPU_PAD192_TC, INVGPU_PAD192_MINB) X(float, 256, INVGPU_PAD256_TR, INVGPU_PAD256_TC, INVGPU_PAD256_MINB)
#define INVGPU_SWEEP_PAD_F64(X) X(double, 16, 2, 2, 2) X(double, 32, 4, 4, 2) X(double, 64, 8, 8, 4) X(double, 128, 16, 16, 1)
#define INVGPU_SWEEP_PAD_ALL(X) INVGPU_SWEEP_PAD_F32(X) INVGPU_SWEEP_PAD_F64(X)

// general inverse, 2-D register tile Gauss-Jordan (gj_tile_kernels.cuh):  X(T, N, TR, TC, MINB); N = padded order, ascending
// (thread grids measured on B200, fraction of the HBM roofline: n = 64 fp32 8x4: 0.067, 2x16: 0.057; n = 128 fp32 16x8: 0.039,
//  4x32: 0.047 -- the kernel is instruction-cache bound, a one-block-wide tile has 4 step bodies instead of N / TC)
#define INVGPU_GJT_F32(X) X(float, 64, 8, 4, 3) X(float, 128, 4, 32, 3)
#define INVGPU_GJT_F64(X) X(double, 32, 4, 4, 2) X(double, 64, 8, 8, 4) X(double, 128, 16, 16, 1)
#define INVGPU_GJT_ALL(X) INVGPU_GJT_F32(X) INVGPU_GJT_F64(X)

// n = 8: one thread per matrix, sweep in registers, TMA tile I/O (spd8_tma_kernel):  X(T, NBUF, MINB)
#define INVGPU_SPD8_TMA_F32(X) X(float, 2, 3)
#define INVGPU_SPD8_TMA_F64(X) X(double, 1, 3)
#define INVGPU_SPD8_TMA_ALL(X) INVGPU_SPD8_TMA_F32(X) INVGPU_SPD8_TMA_F64(X)

// one thread per matrix with per-lane 1-D bulk copies into padded slots (spd_thread_bulk_kernel):  X(T, N, WARPS, MINB)
#define INVGPU_THREAD_BULK_F32(X) X(float, 16, 2, 3)
#define INVGPU_THREAD_BULK_F64(X)
#define INVGPU_THREAD_BULK_ALL(X) INVGPU_THREAD_BULK_F32(X) INVGPU_THREAD_BULK_F64(X)

// fused GP mean / variance, one thread per evaluation (gp_thread_kernel):  X(T, N, WARPS, MINB)
#define INVGPU_GP_THREAD_F32(X) X(float, 8, 4, 4) X(float, 16, 2, 3)
#define INVGPU_GP_THREAD_F64(X) X(double, 8, 4, 3)
#define INVGPU_GP_THREAD_ALL(X) INVGPU_GP_THREAD_F32(X) INVGPU_GP_THREAD_F64(X)

// general inverse, column-split lanes, matrix in registers (gj_colsplit_kernel):  X(T, N, CL, WARPS, MINB)
#define INVGPU_GJC_F32(X) INVGPU_LAB_ONLY(X(float, 16, 8, 4, 2) X(float, 32, 4, 4, 2))
#define INVGPU_GJC_F64(X) INVGPU_LAB_ONLY(X(double, 16, 4, 4, 2))
#define INVGPU_GJC_ALL(X) INVGPU_GJC_F32(X) INVGPU_GJC_F64(X)
#define INVGPU_GJC_DEFAULT(TT) false                  // INVGPU_GJ_KERNEL=colsplit only: fp32 0.227 / 0.122 vs 0.289 / 0.203 for the lean lane = row kernel, fp64 n = 16 0.245 vs 0.311

// gj_roll2d_kernels.cuh -- Gauss-Jordan inverse with partial pivoting for 64 < n <= 128 (fp32): ONE CTA per matrix, the
// matrix in registers as a (32 lanes) x (N / 32 warps) grid of threads, ROLLED pivot loop.
//
// Same job as gj_tile_kernel (gj_tile_kernels.cuh: replaces the reference's `invert` launch loop,
// src/gauss/batched_invert.cu:84-95, and its cuBLAS getrf/getriBatched path, src/gauss/inverse_gpu.cu:24-50).  ncu on the
// tile kernel at n = 128 (profiles/r2_gj_tile128_summary.md): 1212 executed warp instructions per pivot of which 256 are
// FFMA2 -- the rest is the serial arg-max of the four owner lanes over their 32 row slots, the 16-way switch that finds the
// register slot of the pivot row, operand scaling and 4-byte scattered I/O; barrier / wait / short_scoreboard are the top
// stalls.  This kernel keeps the per-pivot overhead off the FMA threads:
//
//  * warp w owns the 32 natural columns 32 w .. 32 w + 31, lane l owns the ROWS = N / 32 consecutive rows ROWS l ..: a
//    thread holds ROWS x 32 elements as horizontal FFMA2 pairs.  Per pivot it needs ROWS multipliers (one 128-bit load) and
//    the 32 pivot-row entries of its columns (eight 128-bit broadcast loads, warp-uniform address) for 16 ROWS FFMA2.
//  * ROTATING WINDOW (as gj_roll_kernels.cuh): the OWNER warp's 32-column window moves down one register per pivot -- the FMA
//    destination simply is the neighbouring register, two pivots per loop iteration keep the pairs aligned -- so its pivot
//    column is always register pair 0 and the loop body is two step bodies for any N.  The window is cyclic: position 0
//    re-enters at position 31, exactly where the new inverse column belongs; a warp owns 32 consecutive pivots = one full
//    turn, so every window is in natural order whenever ownership changes hands and at the end.  The other warps update in
//    place (GJ2D_OWNER_ONLY_ROTATION; rotating all of them measured the same at n = 128, 7 % slower padded).
//  * ONE uniform update for everything:  a_ic += z_i * row_c  with  z_i = -a_ik / pivot  (z = 0 in the pivot row: rows are
//    never scaled inside the loop, each is multiplied once at the end by the reciprocal of its own pivot).  The owner warp
//    restarts its pivot column as e_p (1 in the pivot row, 0 elsewhere) BEFORE the pivot row is published, so the published
//    row carries row_k = 1 and the same FMA produces the new column (0 + z_i * 1 = z_i; pivot row 1 + 0 * 1 = 1).
//  * pivot search by the owner WARP: all 32 lanes x ROWS slots at once, |a| as an unsigned key, one REDUX max + ROWS
//    ballots; the first maximum in row order wins (isamax / the oracle).  Multipliers, pivot row index and 1 / pivot go
//    through shared memory: two CTA barriers per pivot, three CTAs per SM interleave.
//
// Rows are never swapped (implicit pivoting); with piv[k] = pivot row of step k and mystep[r] = step of row r the registers
// finally hold  M[r][k] = Ainv[mystep[r]][piv[k]]  up to the deferred row scale.  info: k (1-based) if no non-zero pivot
// exists for column k (sgetrf's "U(k,k) is exactly zero"; a NaN column counts as singular); flagged outputs are NaN.
// Runtime order n <= N: the matrix is embedded as blockdiag(A, I).
#pragma once

#include "gj_roll_kernels.cuh"

namespace invgpu {

template <typename T, int N, int CW_>
struct GjRoll2dGeo {
    static constexpr int ROWS = N / 32;                 // rows per lane
    static constexpr int CW = CW_, H = CW / 2;          // window: CW columns = CW / 2 pairs per thread
    static constexpr int WARPS = N / CW;                // column groups of CW columns
    static constexpr int BLOCK = 32 * WARPS;
    // multipliers z[N] | pivot row [N] | meta (row index, 1 / pivot, singular flag, -) | piv[N] (pivot row of step k) |
    // step_of_row[N] | 1 / pivot of row [N]   (the last two are read once, by the final store: registers are short here)
    static constexpr int WORDS = 2 * N + 4 + 3 * N;
    static_assert(N % 32 == 0 && ROWS == 4 && sizeof(T) == 4 && (CW == 16 || CW == 32), "built for fp32, N = 128 (the multipliers of a lane are one 128-bit word)");
};

// A register pair as ONE 64-bit value.  Held as two floats (GjPair) the halves of a pair drifted apart in the register
// allocation of this kernel -- ptxas re-assembled every FFMA2 operand with two moves (355 moves for 128 FFMA2 per loop
// iteration); a .b64 virtual register is an aligned pair by construction.
#ifndef GJ2D_TWO_REDUX
#define GJ2D_TWO_REDUX 1
#endif
#ifndef GJ2D_OWNER_ONLY_ROTATION
#define GJ2D_OWNER_ONLY_ROTATION 1
#endif
typedef unsigned long long P64;
__device__ __forceinline__ P64 p64_pack(float x, float y) { P64 p; asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(x), "f"(y)); return p; }
__device__ __forceinline__ float p64_lo(P64 p) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p)); return x; }
__device__ __forceinline__ float p64_hi(P64 p) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p)); return y; }
// replace one half of a pair.  Plain integer arithmetic on purpose: the unpack / re-pack asm pair (mov.b64 {x, y}, p; mov.b64 p, {x, e})
// was miscompiled by ptxas 12.9 in gj_roll2d_ws_kernel -- the temporary feeding e was allocated ON TOP of the live low half of p
// (correct PTX, wrong SASS: every even column of the result was off; found with tools/dbg_gjws.py)
__device__ __forceinline__ P64 p64_set_lo(P64 p, float e) { return (p & 0xffffffff00000000ull) | (P64)__float_as_uint(e); }
__device__ __forceinline__ P64 p64_set_hi(P64 p, float e) { return (p & 0x00000000ffffffffull) | ((P64)__float_as_uint(e) << 32); }
__device__ __forceinline__ P64 p64_fma(P64 zz, P64 r, P64 a) { P64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(zz), "l"(r), "l"(a)); return d; }
// 64-bit shared store with the row slot as an immediate operand: the four branches of gj2d_publish differ in an operand that
// cannot become a phi, so the compiler cannot sink them into one block with a run-time register index (which would put
// the whole tile into local memory -- it did)
template <int Q> __device__ __forceinline__ void store_p64_slot(float *p, P64 v) {
    // (a .v2.f32 store: ptxas merges neighbouring .b64 stores into 128-bit ones, whose quad alignment flips with every
    // window move -- four staging moves per store plus a physical rotation of the window at the loop end)
    asm volatile("{ .reg .f32 lo, hi;  mov.b64 {lo, hi}, %1;  st.volatile.shared.v2.f32 [%0], {lo, hi}; }  // row slot %2"
                 ::"r"((unsigned)__cvta_generic_to_shared(p)), "l"(v), "n"(Q) : "memory");
}
template <int N, int H, int Q>
__device__ __forceinline__ void gj2d_publish(P64 (&ap)[N / 32][H], float *dst, int pq) {
    if (pq == Q) {
        #pragma unroll
        for (int i = 0; i < H; ++i) store_p64_slot<Q>(dst + 2 * i, ap[Q][i]);
    } else if constexpr (Q + 1 < N / 32) gj2d_publish<N, H, Q + 1>(ap, dst, pq);
}

// one pivot step; B = 0: the pivot column is pair 0 .x, the window stays; B = 1: pair 0 .y, the window moves on by one pair
template <typename T, int N, int CW, int B>
__device__ __forceinline__ void gj2d_step(P64 (&ap)[N / 32][CW / 2], T *zline, T *rowline, T *meta, int *piv, int k, int w, int lane,
                                          unsigned &pivoted, int &st) {
    using G = GjRoll2dGeo<T, N, CW>;
    constexpr int ROWS = G::ROWS, H = G::H;
    if (w == k / CW) {                                           // the owner warp of column k: search, multipliers, restart the column
        T v[ROWS];
        unsigned key[ROWS], mykey = 0u;
        #pragma unroll
        for (int q = 0; q < ROWS; ++q) {
            v[q] = B ? p64_hi(ap[q][0]) : p64_lo(ap[q][0]);
            const float av = fabsf(v[q]);
            key[q] = (!((pivoted >> q) & 1u) && av == av) ? __float_as_uint(av) : 0u;
            mykey = max(mykey, key[q]);
        }
        const unsigned mx = __reduce_max_sync(0xffffffffu, mykey);
#if GJ2D_TWO_REDUX
        // second reduction: the smallest candidate row (first maximum in row order, like isamax / the oracle) together with the
        // sign of its value -- (row << 1 | sign), the row dominates the order -- so neither ballots nor a shuffle of the pivot
        // value are needed: |pivot| is mx itself
        unsigned cand = 0xffffffffu;
        #pragma unroll
        for (int q = 0; q < ROWS; ++q) {
            const unsigned code = ((unsigned)(ROWS * lane + q) << 1) | (__float_as_uint(v[q]) >> 31);
            cand = (!((pivoted >> q) & 1u) && key[q] == mx) ? min(cand, code) : cand;
        }
        const unsigned best = __reduce_min_sync(0xffffffffu, cand);
        const int prow = (int)(best >> 1);
        const int pl = prow / ROWS, pq = prow % ROWS;
        const T r = fast_rcp<T>(__uint_as_float(mx | (best << 31)));
#else
        int prow = 1 << 30;                                        // first maximum in row order (row = ROWS lane + q)
        #pragma unroll
        for (int q = 0; q < ROWS; ++q) {
            const unsigned cq = __ballot_sync(0xffffffffu, !((pivoted >> q) & 1u) && key[q] == mx);
            if (cq) prow = min(prow, ROWS * (__ffs((int)cq) - 1) + q);
        }
        const int pl = prow / ROWS, pq = prow % ROWS;
        T mine = v[0];
        #pragma unroll
        for (int q = 1; q < ROWS; ++q) mine = (pq == q) ? v[q] : mine;
        const T r = fast_rcp<T>(__shfl_sync(0xffffffffu, mine, pl));
#endif
        T z[ROWS];
        #pragma unroll
        for (int q = 0; q < ROWS; ++q) {
            const bool isp = (lane == pl) && (pq == q);
            z[q] = isp ? T(0) : -v[q] * r;
            const T e = isp ? T(1) : T(0);
            ap[q][0] = B ? p64_set_hi(ap[q][0], e) : p64_set_lo(ap[q][0], e);
        }
        *reinterpret_cast<float4 *>(zline + ROWS * lane) = make_float4(z[0], z[1], z[2], z[3]);
        if (lane == 0) *reinterpret_cast<float4 *>(meta) = make_float4(__int_as_float(prow), r, mx == 0u ? 1.0f : 0.0f, 0.0f);
    }
    __syncthreads();
    const float4 zz = *reinterpret_cast<const float4 *>(zline + ROWS * lane);
    const float4 mt = *reinterpret_cast<const float4 *>(meta);
    const P64 z[ROWS] = {p64_pack(zz.x, zz.x), p64_pack(zz.y, zz.y), p64_pack(zz.z, zz.z), p64_pack(zz.w, zz.w)};
    const int prow = __float_as_int(mt.x);
    if (st == 0 && mt.z != 0.0f) st = k + 1;                       // uniform in the CTA
    const bool on_row = lane == prow / ROWS;
    pivoted |= on_row ? (1u << (prow % ROWS)) : 0u;               // bit q: row ROWS lane + q has been a pivot
    if (on_row) {                                                  // this warp's part of the pivot row: raw, one branch per row slot
        if (w == 0) { piv[k] = prow; piv[N + prow] = k; reinterpret_cast<T *>(piv)[2 * N + prow] = mt.y; }
        gj2d_publish<N, H, 0>(ap, rowline + G::CW * w, prow % ROWS);
    }
    __syncthreads();
    const ulonglong2 *pr = reinterpret_cast<const ulonglong2 *>(rowline + G::CW * w);   // two pairs per 128-bit broadcast load
    // Only the OWNER warp moves its window (it needs its pivot column at position 0; 32 pivots = one full turn, so the window is
    // in natural order whenever ownership changes hands); the other warps update in place.
    if (!B || (GJ2D_OWNER_ONLY_ROTATION && w != k / CW)) {
        #pragma unroll
        for (int i2 = 0; i2 < H; i2 += 2) {
            const ulonglong2 rr = pr[i2 / 2];
            #pragma unroll
            for (int q = 0; q < ROWS; ++q) {
                ap[q][i2] = p64_fma(z[q], rr.x, ap[q][i2]);
                ap[q][i2 + 1] = p64_fma(z[q], rr.y, ap[q][i2 + 1]);
            }
        }
    } else {
        P64 first[ROWS];
        {
            const ulonglong2 rr = pr[0];
            #pragma unroll
            for (int q = 0; q < ROWS; ++q) {
                first[q] = p64_fma(z[q], rr.x, ap[q][0]);           // position 0 re-enters at the end of the window
                ap[q][0] = p64_fma(z[q], rr.y, ap[q][1]);
            }
        }
        #pragma unroll
        for (int i2 = 2; i2 < H; i2 += 2) {
            const ulonglong2 rr = pr[i2 / 2];
            #pragma unroll
            for (int q = 0; q < ROWS; ++q) {
                ap[q][i2 - 1] = p64_fma(z[q], rr.x, ap[q][i2]);
                ap[q][i2] = p64_fma(z[q], rr.y, ap[q][i2 + 1]);
            }
        }
        #pragma unroll
        for (int q = 0; q < ROWS; ++q) ap[q][H - 1] = first[q];
    }
}

// EXACT: the runtime order is N; otherwise n <= N, embedded as blockdiag(A, I)
template <typename T, int N, int CW_, typename IO, int MINB, bool EXACT>
__global__ void __launch_bounds__((GjRoll2dGeo<T, N, CW_>::BLOCK), MINB)
gj_roll2d_kernel(IO io, int n_runtime, i64 batch, int *__restrict__ info) {
    using G = GjRoll2dGeo<T, N, CW_>;
    constexpr int ROWS = G::ROWS, H = G::H, CW = G::CW;
    const int n = EXACT ? N : n_runtime;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *zline = reinterpret_cast<T *>(smem_raw), *rowline = zline + N, *meta = rowline + N;
    int *piv = reinterpret_cast<int *>(meta + 4);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;

    #pragma unroll 1
    for (i64 m = blockIdx.x; m < batch; m += gridDim.x) {
        const T *__restrict__ src = io.src(m);
        // ap[q][i] = A(row ROWS lane + q, columns CW w + 2 i, + 1); identity padding outside n
        P64 ap[ROWS][H];
        if (EXACT && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
            #pragma unroll
            for (int i = 0; i < H; ++i) {
                const float4 c0 = __ldcs(reinterpret_cast<const float4 *>(src + (size_t)(CW * w + 2 * i) * N + ROWS * lane));
                const float4 c1 = __ldcs(reinterpret_cast<const float4 *>(src + (size_t)(CW * w + 2 * i + 1) * N + ROWS * lane));
                ap[0][i] = p64_pack(c0.x, c1.x); ap[1][i] = p64_pack(c0.y, c1.y);
                ap[2][i] = p64_pack(c0.z, c1.z); ap[3][i] = p64_pack(c0.w, c1.w);
            }
        } else {
            #pragma unroll
            for (int i = 0; i < H; ++i) {
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) {
                    const int row = ROWS * lane + q;
                    T e[2];
                    #pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int c = CW * w + 2 * i + h;
                        e[h] = (row == c) ? T(1) : T(0);
                        if (row < n && c < n) e[h] = __ldcs(src + (size_t)c * n + row);
                    }
                    ap[q][i] = p64_pack(e[0], e[1]);
                }
            }
        }

        int st = 0;
        unsigned pivoted = 0u;

        #pragma unroll 1
        for (int kk = 0; kk < N / 2; ++kk) {
            gj2d_step<T, N, CW, 0>(ap, zline, rowline, meta, piv, 2 * kk, w, lane, pivoted, st);
            gj2d_step<T, N, CW, 1>(ap, zline, rowline, meta, piv, 2 * kk + 1, w, lane, pivoted, st);
        }
        __syncthreads();                                           // piv[] complete
        int mystep[ROWS];
        T rscale[ROWS];
        #pragma unroll
        for (int q = 0; q < ROWS; ++q) { mystep[q] = piv[N + ROWS * lane + q]; rscale[q] = reinterpret_cast<const T *>(piv)[2 * N + ROWS * lane + q]; }

        if (threadIdx.x == 0 && info) info[m] = (st > n) ? 0 : st; // a "singular" padded column cannot happen; guard anyway
        T *__restrict__ dst = io.dst(m);
        const bool bad = st != 0 && st <= n;
        if (EXACT && !bad) {                                       // the common case: 32-bit offsets, no bounds predicates
            #pragma unroll
            for (int j = 0; j < CW; ++j) {
                const int ocol = piv[CW * w + j] * N;              // broadcast read
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) __stcs(dst + ocol + mystep[q], ((j & 1) ? p64_hi(ap[q][j >> 1]) : p64_lo(ap[q][j >> 1])) * rscale[q]);
            }
        } else {
            #pragma unroll
            for (int j = 0; j < CW; ++j) {
                const int c = CW * w + j;
                const int ocol = piv[c];
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) {
                    const int row = ROWS * lane + q, orow = mystep[q];
                    if (bad) { if (row < n && c < n) dst[(size_t)c * n + row] = dev_nan<T>(); }
                    else if (orow < n && ocol < n) __stcs(dst + (size_t)ocol * n + orow, ((j & 1) ? p64_hi(ap[q][j >> 1]) : p64_lo(ap[q][j >> 1])) * rscale[q]);
                }
            }
        }
        __syncthreads();                                           // piv / lines are reused by the next matrix
    }
}

// ==========================================================================================
// WARP-SPECIALISED form: the same tile, window and update on warps 0 .. 3, plus a FIFTH warp that does nothing but the
// pivot search.  In the kernel above the three other warps wait at the first barrier of every step while the owner warp
// runs its ~100-instruction search chain (23 % of all warp time, profiles/r2_gj_roll2d_128_summary.md) -- and giving the
// search to the next owner as a look-ahead only moves the wait to the next barrier, because that warp still executes
// search + pass back to back.  Here the owner of column k + 1 hands the column over as soon as its first chunk of pass k is
// done (four values per lane through shared memory, named barrier 1: arrive by the owner, sync by the pivot warp); the
// pivot warp searches and publishes multipliers / row index / reciprocal for step k + 1 WHILE the four FMA warps finish
// pass k.  The owner restarts its column as e_p after the barrier, from the published row index.
// MEASURED (B200, 16 384 x 128x128): 4.60 ms at two CTAs per SM against 4.21-4.26 ms of the four-warp kernel (one CTA per SM:
// 7.09 ms): the search is about as long as a pass, so the FMA warps still wait at the first barrier (15 % of the samples),
// and two CTAs per SM interleave worse than three.  Built by `make lab=1` only; INVGPU_GJR2_WS=1 selects it.
// ==========================================================================================
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <typename T, int ROWS>
__device__ __forceinline__ void gj2d_pivot_search(const T (&v)[ROWS], unsigned &pivoted, int lane, T *zline, T *meta) {
    unsigned key[ROWS], mykey = 0u;
    #pragma unroll
    for (int q = 0; q < ROWS; ++q) {
        const float av = fabsf(v[q]);
        key[q] = (!((pivoted >> q) & 1u) && av == av) ? __float_as_uint(av) : 0u;
        mykey = max(mykey, key[q]);
    }
    const unsigned mx = __reduce_max_sync(0xffffffffu, mykey);
    int prow = 1 << 30;                                            // first maximum in row order (row = ROWS lane + q)
    #pragma unroll
    for (int q = 0; q < ROWS; ++q) {
        const unsigned cq = __ballot_sync(0xffffffffu, !((pivoted >> q) & 1u) && key[q] == mx);
        if (cq) prow = min(prow, ROWS * (__ffs((int)cq) - 1) + q);
    }
    const int pl = prow / ROWS, pq = prow % ROWS;
    T mine = v[0];
    #pragma unroll
    for (int q = 1; q < ROWS; ++q) mine = (pq == q) ? v[q] : mine;
    const T r = fast_rcp<T>(__shfl_sync(0xffffffffu, mine, pl));
    T z[ROWS];
    #pragma unroll
    for (int q = 0; q < ROWS; ++q) z[q] = ((lane == pl) && (pq == q)) ? T(0) : -v[q] * r;
    *reinterpret_cast<float4 *>(zline + ROWS * lane) = make_float4(z[0], z[1], z[2], z[3]);
    if (lane == 0) *reinterpret_cast<float4 *>(meta) = make_float4(__int_as_float(prow), r, mx == 0u ? 1.0f : 0.0f, 0.0f);
    pivoted |= (lane == pl) ? (1u << pq) : 0u;
}

// pass of step k on an FMA warp; LOOK: this warp owns column k + 1 -- hand it to the pivot warp after the first chunk
template <typename T, int N, int B, bool LOOK>
__device__ __forceinline__ void gj2d_ws_pass(P64 (&ap)[N / 32][16], const P64 (&z)[N / 32], const ulonglong2 *pr, T *colbuf, int lane) {
    constexpr int ROWS = N / 32, H = 16;
    auto hand_over = [&]() {
        // B = 0: the next pivot column is pair 0 .y (window stays); B = 1: pair 0 .x of the moved window
        *reinterpret_cast<float4 *>(colbuf + ROWS * lane) =
            B ? make_float4(p64_lo(ap[0][0]), p64_lo(ap[1][0]), p64_lo(ap[2][0]), p64_lo(ap[3][0]))
              : make_float4(p64_hi(ap[0][0]), p64_hi(ap[1][0]), p64_hi(ap[2][0]), p64_hi(ap[3][0]));
        named_bar_arrive(1, 64);                                  // (the barrier orders the store above: PTX producer / consumer pattern)
    };
    if (!B) {
        #pragma unroll
        for (int i2 = 0; i2 < H; i2 += 2) {
            const ulonglong2 rr = pr[i2 / 2];
            #pragma unroll
            for (int q = 0; q < ROWS; ++q) {
                ap[q][i2] = p64_fma(z[q], rr.x, ap[q][i2]);
                ap[q][i2 + 1] = p64_fma(z[q], rr.y, ap[q][i2 + 1]);
            }
            if (LOOK && i2 == 0) hand_over();
        }
    } else {
        P64 first[ROWS];
        {
            const ulonglong2 rr = pr[0];
            #pragma unroll
            for (int q = 0; q < ROWS; ++q) {
                first[q] = p64_fma(z[q], rr.x, ap[q][0]);           // position 0 re-enters at the end of the window
                ap[q][0] = p64_fma(z[q], rr.y, ap[q][1]);
            }
            if (LOOK) hand_over();
        }
        #pragma unroll
        for (int i2 = 2; i2 < H; i2 += 2) {
            const ulonglong2 rr = pr[i2 / 2];
            #pragma unroll
            for (int q = 0; q < ROWS; ++q) {
                ap[q][i2 - 1] = p64_fma(z[q], rr.x, ap[q][i2]);
                ap[q][i2] = p64_fma(z[q], rr.y, ap[q][i2 + 1]);
            }
        }
        #pragma unroll
        for (int q = 0; q < ROWS; ++q) ap[q][H - 1] = first[q];
    }
}

template <typename T, int N, int B>
__device__ __forceinline__ void gj2d_ws_step(P64 (&ap)[N / 32][16], T *zline, T *rowline, T *meta, T *colbuf, int *piv, int k, int w, int lane, int &st) {
    constexpr int ROWS = N / 32, H = 16, CW = 32;
    named_bar_sync(2, 160);                                               // multipliers / meta of step k are visible
    const float4 zz = *reinterpret_cast<const float4 *>(zline + ROWS * lane);
    const float4 mt = *reinterpret_cast<const float4 *>(meta);
    const P64 z[ROWS] = {p64_pack(zz.x, zz.x), p64_pack(zz.y, zz.y), p64_pack(zz.z, zz.z), p64_pack(zz.w, zz.w)};
    const int prow = __float_as_int(mt.x);
    if (st == 0 && mt.z != 0.0f) st = k + 1;                       // uniform in the CTA
    const bool on_row = lane == prow / ROWS;
    if (w == k / CW) {                                             // owner: restart the pivot column as e_p
        #pragma unroll
        for (int q = 0; q < ROWS; ++q) {
            const T e = (on_row && (prow % ROWS) == q) ? T(1) : T(0);
            ap[q][0] = B ? p64_set_hi(ap[q][0], e) : p64_set_lo(ap[q][0], e);
        }
    }
    if (on_row) {                                                  // this warp's part of the pivot row: raw, one branch per row slot
        if (w == 0) { piv[k] = prow; piv[N + prow] = k; reinterpret_cast<T *>(piv)[2 * N + prow] = mt.y; }
        gj2d_publish<N, H, 0>(ap, rowline + CW * w, prow % ROWS);
    }
    named_bar_sync(2, 160);                                               // the pivot row is visible; everybody has read z / meta
    const ulonglong2 *pr = reinterpret_cast<const ulonglong2 *>(rowline + CW * w);
    if (k + 1 < N && w == (k + 1) / CW) gj2d_ws_pass<T, N, B, true>(ap, z, pr, colbuf, lane);
    else gj2d_ws_pass<T, N, B, false>(ap, z, pr, colbuf, lane);
}

template <typename T, int N, typename IO, int MINB, bool EXACT>
__global__ void __launch_bounds__(256, 2)
gj_roll2d_ws_kernel(IO io, int n_runtime, i64 batch, int *__restrict__ info) {
    using G = GjRoll2dGeo<T, N, 32>;
    constexpr int ROWS = G::ROWS, H = G::H, CW = 32;
    const int n = EXACT ? N : n_runtime;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *zline = reinterpret_cast<T *>(smem_raw), *rowline = zline + N, *meta = rowline + N;
    int *piv = reinterpret_cast<int *>(meta + 4);
    T *colbuf = reinterpret_cast<T *>(piv + 3 * N);                // the next pivot column on its way to the pivot warp
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // Registers are allocated per CTA in units of four warps: a five-warp CTA at 200 registers occupies 8 x 32 x 200 and only
    // one fits an SM.  So the CTA is two full warpgroups launched at 128 registers (two CTAs per SM); the second one (pivot
    // warp + three warps that exit at once) gives its registers back and the FMA warpgroup takes them (setmaxnreg).
    // CTA-wide synchronisation below is named barrier 2 over the 160 threads that stay.
    if (w >= 4) {                                                  // ---------------- the pivot warp (and three warps that only return registers)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (w > 4) return;
        #pragma unroll 1
        for (i64 m = blockIdx.x; m < batch; m += gridDim.x) {
            unsigned pivoted = 0u;
            #pragma unroll 1
            for (int k = 0; k < N; ++k) {
                named_bar_sync(1, 64);                             // column k has arrived
                const float4 c = *reinterpret_cast<const float4 *>(colbuf + ROWS * lane);
                const T v[ROWS] = {c.x, c.y, c.z, c.w};
                gj2d_pivot_search<T, ROWS>(v, pivoted, lane, zline, meta);
                named_bar_sync(2, 160);                            // first barrier of step k
                named_bar_sync(2, 160);                            // second barrier of step k
            }
            named_bar_sync(2, 160);
            named_bar_sync(2, 160);
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");          // ---------------- the four FMA warps (code dominated by the increase)

    #pragma unroll 1
    for (i64 m = blockIdx.x; m < batch; m += gridDim.x) {
        const T *__restrict__ src = io.src(m);
        P64 ap[ROWS][H];
        if (EXACT && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
            #pragma unroll
            for (int i = 0; i < H; ++i) {
                const float4 c0 = __ldcs(reinterpret_cast<const float4 *>(src + (size_t)(CW * w + 2 * i) * N + ROWS * lane));
                const float4 c1 = __ldcs(reinterpret_cast<const float4 *>(src + (size_t)(CW * w + 2 * i + 1) * N + ROWS * lane));
                ap[0][i] = p64_pack(c0.x, c1.x); ap[1][i] = p64_pack(c0.y, c1.y);
                ap[2][i] = p64_pack(c0.z, c1.z); ap[3][i] = p64_pack(c0.w, c1.w);
            }
        } else {
            #pragma unroll
            for (int i = 0; i < H; ++i) {
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) {
                    const int row = ROWS * lane + q;
                    T e[2];
                    #pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int c = CW * w + 2 * i + h;
                        e[h] = (row == c) ? T(1) : T(0);
                        if (row < n && c < n) e[h] = __ldcs(src + (size_t)c * n + row);
                    }
                    ap[q][i] = p64_pack(e[0], e[1]);
                }
            }
        }
        if (w == 0) {                                              // column 0 to the pivot warp
            *reinterpret_cast<float4 *>(colbuf + ROWS * lane) = make_float4(p64_lo(ap[0][0]), p64_lo(ap[1][0]), p64_lo(ap[2][0]), p64_lo(ap[3][0]));
            named_bar_arrive(1, 64);
        }
        int st = 0;
        #pragma unroll 1
        for (int kk = 0; kk < N / 2; ++kk) {
            gj2d_ws_step<T, N, 0>(ap, zline, rowline, meta, colbuf, piv, 2 * kk, w, lane, st);
            gj2d_ws_step<T, N, 1>(ap, zline, rowline, meta, colbuf, piv, 2 * kk + 1, w, lane, st);
        }
        named_bar_sync(2, 160);                                           // piv[] complete
        int mystep[ROWS];
        T rscale[ROWS];
        #pragma unroll
        for (int q = 0; q < ROWS; ++q) { mystep[q] = piv[N + ROWS * lane + q]; rscale[q] = reinterpret_cast<const T *>(piv)[2 * N + ROWS * lane + q]; }

        if (threadIdx.x == 0 && info) info[m] = (st > n) ? 0 : st; // a "singular" padded column cannot happen; guard anyway
        T *__restrict__ dst = io.dst(m);
        const bool bad = st != 0 && st <= n;
        if (EXACT && !bad) {                                       // the common case: 32-bit offsets, no bounds predicates
            #pragma unroll
            for (int j = 0; j < CW; ++j) {
                const int ocol = piv[CW * w + j] * N;              // broadcast read
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) __stcs(dst + ocol + mystep[q], ((j & 1) ? p64_hi(ap[q][j >> 1]) : p64_lo(ap[q][j >> 1])) * rscale[q]);
            }
        } else {
            #pragma unroll
            for (int j = 0; j < CW; ++j) {
                const int c = CW * w + j;
                const int ocol = piv[c];
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) {
                    const int row = ROWS * lane + q, orow = mystep[q];
                    if (bad) { if (row < n && c < n) dst[(size_t)c * n + row] = dev_nan<T>(); }
                    else if (orow < n && ocol < n) __stcs(dst + (size_t)ocol * n + orow, ((j & 1) ? p64_hi(ap[q][j >> 1]) : p64_lo(ap[q][j >> 1])) * rscale[q]);
                }
            }
        }
        named_bar_sync(2, 160);                                           // piv / lines are reused by the next matrix
    }
}

}  // namespace invgpu

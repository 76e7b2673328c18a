// sweep_kernels.cuh -- SPD inverse as ONE symmetric sweep, rolled form for square thread grids.
//
// Same mathematics as onesweep_kernels.cuh (potrf + trtri + lauum merged into n rank-1 updates of the
// lower triangle, trailing matrix kept negated: T = -A, and T -> A^-1 after the n-th pivot), with the
// instruction diet ncu asked for (profiles/r1_*: the one-sweep kernels are ISSUE-bound, 58-69 % issue
// slots busy with only ~45 % of the instructions being FMAs):
//
//   * no square roots, owners publish RAW values.  With z_i = T_ik (i != k), pivot d = -T_kk and
//     r = 1/d the update is  T_ic += (r z_i) z_c ;  T_ik <- -r z_i ;  T_kk <- r.  Publishing z_k := -1
//     makes all three cases the same FMA:  x_i = r z_i  (row operand, scaled),  y_c = z_c (column
//     operand, raw)  =>  x_k = -r, y_k = -1  and  x_i y_k = -r z_i,  x_k y_k = r.  No special cases, no
//     pivot needed before the barrier (the shuffle -> rsqrt -> scale -> store chain is gone), 8 instead
//     of 16 multiplies per pivot and thread.  The pivot itself travels in word N of the line.
//   * fp32: the tile is held as vertically adjacent PAIRS (64-bit registers) and updated with
//     fma.rn.f32x2 (FFMA2: two FMAs per issue slot, measured at full FMA rate and free of the operand
//     bank conflicts of the scalar rank-1 form, tools/microbench.cu); the scaling uses mul.rn.f32x2.
//   * pivot order: k = 4 (P Q + t) + w rolled over t (P owners take turns); the body for a given (Q, w)
//     is compiled once: N / P bodies of code for any N.  The order is a symmetric permutation of the
//     natural one, so A^-1 is unchanged (to rounding); `info` of a flagged matrix is recomputed in
//     natural order by one thread so that it is LAPACK's spotrf info (reference src/inverse.c:92-95).
//
// Layout: P x P threads per matrix, 4x4 sub-blocks dealt cyclically (TileGeo<N, P, P, false>); blocks
// strictly above the diagonal for every thread are neither stored nor updated.
#pragma once

#include "tile_kernels.cuh"

namespace invgpu {

// two vertically adjacent tile elements (rows 2i, 2i+1 of one column)
template <typename T> struct Pair2;
template <> struct Pair2<float> {
    float2 v;
    static __device__ __forceinline__ Pair2 make(float lo, float hi) { Pair2 p; p.v = make_float2(lo, hi); return p; }
    __device__ __forceinline__ float lo() const { return v.x; }
    __device__ __forceinline__ float hi() const { return v.y; }
    // this += x * (y, y)
    __device__ __forceinline__ void fma_bcast(const Pair2 &x, float y) { v = __ffma2_rn(x.v, make_float2(y, y), v); }
    __device__ __forceinline__ void scale(float s) { v = __fmul2_rn(v, make_float2(s, s)); }
};
template <> struct Pair2<double> {
    double l, h;
    static __device__ __forceinline__ Pair2 make(double lo, double hi) { Pair2 p; p.l = lo; p.h = hi; return p; }
    __device__ __forceinline__ double lo() const { return l; }
    __device__ __forceinline__ double hi() const { return h; }
    __device__ __forceinline__ void fma_bcast(const Pair2 &x, double y) { l = fma(x.l, y, l); h = fma(x.h, y, h); }
    __device__ __forceinline__ void scale(double s) { l *= s; h *= s; }
};

template <typename T> __device__ __forceinline__ T dev_min(T a, T b);
template <> __device__ __forceinline__ float dev_min<float>(float a, float b) { return fminf(a, b); }
template <> __device__ __forceinline__ double dev_min<double>(double a, double b) { return fmin(a, b); }
template <typename T> __device__ __forceinline__ T dev_rcp_fast(T x);
template <> __device__ __forceinline__ float dev_rcp_fast<float>(float x) {
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;   // one MUFU.RCP, <= 1 ulp
}
template <> __device__ __forceinline__ double dev_rcp_fast<double>(double x) { return 1.0 / x; }

template <int N, int P>
struct SweepGeo {
    using G = TileGeo<N, P, P, false>;
    static constexpr int LINE = N + 4;                                   // z line + the pivot word (16-byte aligned)
    static constexpr int WORDS = ((2 * LINE + 31) / 32) * 32 + (G::LANES < 32 ? 8 : 0);
};

template <typename T, int N, int P, typename IO, int MINB>
__global__ void __launch_bounds__((TileGeo<N, P, P, false>::BLOCK), MINB)
sweep_rolled_kernel(IO io, i64 batch, int *__restrict__ info) {
    using G = TileGeo<N, P, P, false>;
    using SG = SweepGeo<N, P>;
    using PR = Pair2<T>;
    constexpr int S = G::SR;                                       // tile side
    constexpr int NG = S / 4;                                      // 4-groups per thread
    constexpr int H = S / 2;                                       // row pairs per thread
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);

    const int grp = threadIdx.x / G::LANES;
    const int lane = threadIdx.x % G::LANES;
    const int ti = lane / P, tj = lane % P;
    T *sm = smem + grp * SG::WORDS;

    #pragma unroll 1
    for (i64 base = (i64)blockIdx.x * G::MPB; base < batch; base += (i64)gridDim.x * G::MPB) {
        const i64 m = base + grp;
        const bool valid = m < batch;
        const T *__restrict__ src = io.src(valid ? m : batch - 1);

        PR ap[H][S];                                               // ap[i][c] = T(rows 2i, 2i+1 ; column c), negated input
        {
            T a[S][S];
            tile_load_upper<T, N, P, P, false>(a, src, ti, tj);
            #pragma unroll
            for (int i = 0; i < H; ++i)
                #pragma unroll
                for (int c = 0; c < S; ++c) ap[i][c] = PR::make(-a[2 * i][c], -a[2 * i + 1][c]);
        }

        T dmin = T(1), d = T(1);                                   // smallest pivot so far / current pivot
        #pragma unroll
        for (int Q = 0; Q < NG; ++Q) {
            #pragma unroll
            for (int w = 0; w < 4; ++w) {
                constexpr int dummy = 0; (void)dummy;
                const int s = 4 * Q + w;                           // register slot of the pivots of this body
                #pragma unroll 1
                for (int t = 0; t < P; ++t) {
                    const int q = P * Q + t;
                    T *z = sm + (t & 1) * SG::LINE;                // consecutive pivots differ in t (P is even): double buffer
                    // ---- owners publish raw values and restart their slots
#ifndef SWEEP_DEBUG_NOPUB
                    if (tj == t) {                                 // column k: block rows below block q
                        #pragma unroll
                        for (int g = Q; g < NG; ++g) {
                            if (g > Q || ti > t) {
                                st4(z + 4 * (P * g + ti), ap[2 * g][s].lo(), ap[2 * g][s].hi(), ap[2 * g + 1][s].lo(), ap[2 * g + 1][s].hi());
                                ap[2 * g][s] = PR::make(T(0), T(0));
                                ap[2 * g + 1][s] = PR::make(T(0), T(0));
                            }
                        }
                    }
                    if (ti == t) {                                 // row k: block columns left of block q
                        #pragma unroll
                        for (int h = 0; h <= Q; ++h) {
                            if (h < Q || tj < t) {
                                T e[4];
                                #pragma unroll
                                for (int v = 0; v < 4; ++v) {
                                    PR &p = ap[s / 2][4 * h + v];
                                    if (s % 2 == 0) { e[v] = p.lo(); p = PR::make(T(0), p.hi()); }
                                    else { e[v] = p.hi(); p = PR::make(p.lo(), T(0)); }
                                }
                                st4(z + 4 * (P * h + tj), e[0], e[1], e[2], e[3]);
                            }
                        }
                        if (tj == t) {                             // diagonal thread: block q = [row part | -1 | column part]
                            T e[4];
                            #pragma unroll
                            for (int v = 0; v < 4; ++v) {
                                if (v <= w) {                      // element (row s, column 4Q+v); v == w is the pivot
                                    PR &p = ap[s / 2][4 * Q + v];
                                    if (s % 2 == 0) { e[v] = p.lo(); p = PR::make(T(0), p.hi()); }
                                    else { e[v] = p.hi(); p = PR::make(p.lo(), T(0)); }
                                } else {                           // element (row 4Q+v, column s)
                                    PR &p = ap[(4 * Q + v) / 2][s];
                                    if ((4 * Q + v) % 2 == 0) { e[v] = p.lo(); p = PR::make(T(0), p.hi()); }
                                    else { e[v] = p.hi(); p = PR::make(p.lo(), T(0)); }
                                }
                            }
                            z[N] = e[w];                           // -d
                            e[w] = T(-1);
                            st4(z + 4 * q, e[0], e[1], e[2], e[3]);
                        }
                    }
#endif
                    tile_sync<G::LANES>();
                    // ---- everybody: one rank-1 update of the lower triangle
                    d = -z[N];
                    dmin = dev_min(dmin, d);
                    const T r = dev_rcp_fast<T>(d);
                    PR x[H];
                    T y[S];
                    #pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        T x0, x1, x2, x3;
                        ld4(z + 4 * (P * g + ti), x0, x1, x2, x3);
                        x[2 * g] = PR::make(x0, x1); x[2 * g + 1] = PR::make(x2, x3);
                        x[2 * g].scale(r); x[2 * g + 1].scale(r);
                        ld4(z + 4 * (P * g + tj), y[4 * g], y[4 * g + 1], y[4 * g + 2], y[4 * g + 3]);
                    }
                    #pragma unroll
                    for (int i = 0; i < H; ++i)
                        #pragma unroll
                        for (int c = 0; c < S; ++c) {
                            if (c / 4 > i / 2) continue;           // strictly upper for every thread
                            ap[i][c].fma_bcast(x[i], y[c]);
                        }
                }
            }
        }
        tile_sync<G::LANES>();                                       // the lines are reused by the next matrix

        if (!valid) continue;
        T *__restrict__ dst = io.dst(m);
        // a NaN pivot turns every later pivot into NaN, so the last one tells; otherwise the minimum does
        const bool bad = !(dmin > T(0)) || !(d == d);
        int st = 0;
        if (bad) {                                                    // rare: LAPACK's natural-order index, dst as scratch
            if (lane == 0) { st = exact_potrf_info<T>(src, dst, N); if (st == 0) st = N; }
            if (G::LANES <= 32) st = __shfl_sync(__activemask(), st, (threadIdx.x & 31 & ~(G::LANES - 1)));
            else { if (lane == 0) sm[0] = (T)st; __syncthreads(); st = (int)sm[0]; __syncthreads(); }
        }
        if (lane == 0 && info) info[m] = st;
        #pragma unroll
        for (int g = 0; g < NG; ++g) {
            #pragma unroll
            for (int h = 0; h < NG; ++h) {
                const int br = P * g + ti, bc = P * h + tj;
                if (bad) {                                            // flagged: this thread's natural blocks, all NaN
                    #pragma unroll
                    for (int v = 0; v < 4; ++v)
                        stg4(dst + (size_t)(4 * bc + v) * N + 4 * br, dev_nan<T>(), dev_nan<T>(), dev_nan<T>(), dev_nan<T>());
                    continue;
                }
                if (h > g) continue;                                  // strictly upper for every thread
                T b[4][4];                                            // b[row][col] of this block
                #pragma unroll
                for (int v = 0; v < 4; ++v) {
                    b[0][v] = ap[2 * g][4 * h + v].lo(); b[1][v] = ap[2 * g][4 * h + v].hi();
                    b[2][v] = ap[2 * g + 1][4 * h + v].lo(); b[3][v] = ap[2 * g + 1][4 * h + v].hi();
                }
                if (br == bc) {                                       // diagonal block: symmetrise in registers
                    #pragma unroll
                    for (int v = 0; v < 4; ++v)
                        stg4(dst + (size_t)(4 * bc + v) * N + 4 * br, v <= 0 ? b[0][v] : b[v][0], v <= 1 ? b[1][v] : b[v][1],
                             v <= 2 ? b[2][v] : b[v][2], b[3][v]);
                } else if (br > bc) {
                    #pragma unroll
                    for (int v = 0; v < 4; ++v)                       // natural position
                        stg4(dst + (size_t)(4 * bc + v) * N + 4 * br, b[0][v], b[1][v], b[2][v], b[3][v]);
                    #pragma unroll
                    for (int ww = 0; ww < 4; ++ww)                    // mirror image
                        stg4(dst + (size_t)(4 * br + ww) * N + 4 * bc, b[ww][0], b[ww][1], b[ww][2], b[ww][3]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Fully UNROLLED form for warp-sized groups (TR x TC <= 32 threads per matrix, n <= 32): pivots in
// natural order, every owner / slot index static.  Same update rule as above (raw publish, z_k = -1,
// x scaled by 1/d, packed FFMA2).  `info`: the smallest pivot is tracked with one FMNMX per step; a
// flagged matrix gets LAPACK's index from the single-thread recomputation.
// ------------------------------------------------------------------------------------------
template <int N, int TR, int TC>
struct SweepWarpGeo {
    using G = TileGeo<N, TR, TC, false>;
    static constexpr int LINE = N + 4;
    static constexpr int WORDS = ((2 * LINE + 31) / 32) * 32 + (G::LANES < 32 ? 8 : 0);
};

// element `odd` of a pair: read it and zero it
template <typename T, typename PR>
__device__ __forceinline__ T pair_take(PR &p, bool odd) {
    T e;
    if (!odd) { e = p.lo(); p = PR::make(T(0), p.hi()); }
    else { e = p.hi(); p = PR::make(p.lo(), T(0)); }
    return e;
}

template <typename T, int N, int TR, int TC, typename IO, int MINB>
__global__ void __launch_bounds__((TileGeo<N, TR, TC, false>::BLOCK), MINB)
sweep_unrolled_kernel(IO io, i64 batch, int *__restrict__ info) {
    using G = TileGeo<N, TR, TC, false>;
    using SG = SweepWarpGeo<N, TR, TC>;
    using PR = Pair2<T>;
    constexpr int SR = G::SR, SC = G::SC, H = SR / 2;
    static_assert(G::LANES <= 32, "warp-sized groups only");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);

    const int grp = threadIdx.x / G::LANES;
    const int lane = threadIdx.x % G::LANES;
    const int ti = lane / TC, tj = lane % TC;
    T *sm = smem + grp * SG::WORDS;

    #pragma unroll 1
    for (i64 base = (i64)blockIdx.x * G::MPB; base < batch; base += (i64)gridDim.x * G::MPB) {
        const i64 m = base + grp;
        const bool valid = m < batch;
        const T *__restrict__ src = io.src(valid ? m : batch - 1);

        PR ap[H][SC];
        {
            T a[SR][SC];
            tile_load_upper<T, N, TR, TC, false>(a, src, ti, tj);
            #pragma unroll
            for (int i = 0; i < H; ++i)
                #pragma unroll
                for (int c = 0; c < SC; ++c) ap[i][c] = PR::make(-a[2 * i][c], -a[2 * i + 1][c]);
        }

        T dmin = T(1), d = T(1);
        #pragma unroll
        for (int k = 0; k < N; ++k) {
            tile_lockstep<G::LANES>(k);
            T *z = sm + (k & 1) * SG::LINE;
            const int q = k / 4, w = k % 4;
            const int rk = G::rowner(k), ck = G::cowner(k), srk = G::rslot(k), sk = G::cslot(k);
            const int gk = srk / 4, hk = sk / 4;
            // ---- owners of column k: rows in blocks below block q, raw
            if (tj == ck) {
                #pragma unroll
                for (int g = 0; g < SR / 4; ++g) {
                    if (G::rblock(g, TR - 1) < q) continue;          // no thread has rows > k here
                    const int blk = G::rblock(g, 0) + ti;
                    if (blk > q) {
                        st4(z + 4 * blk, ap[2 * g][sk].lo(), ap[2 * g][sk].hi(), ap[2 * g + 1][sk].lo(), ap[2 * g + 1][sk].hi());
                        ap[2 * g][sk] = PR::make(T(0), T(0));
                        ap[2 * g + 1][sk] = PR::make(T(0), T(0));
                    }
                }
            }
            // ---- owners of row k: columns in blocks left of block q, raw
            if (ti == rk) {
                #pragma unroll
                for (int h = 0; h < SC / 4; ++h) {
                    if (G::cblock(h, 0) > q) continue;               // no thread has cols < k here
                    const int blk = G::cblock(h, 0) + tj;
                    if (blk < q) {
                        T e[4];
                        #pragma unroll
                        for (int v = 0; v < 4; ++v) e[v] = pair_take<T>(ap[srk / 2][4 * h + v], srk % 2);
                        st4(z + 4 * blk, e[0], e[1], e[2], e[3]);
                    }
                }
                // ---- the diagonal thread assembles block q = [ row part | -1 | column part ], pivot in word N
                if (tj == ck) {
                    T e[4];
                    #pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        if (v <= w) e[v] = pair_take<T>(ap[srk / 2][4 * hk + v], srk % 2);
                        else e[v] = pair_take<T>(ap[(4 * gk + v) / 2][sk], (4 * gk + v) % 2);
                    }
                    z[N] = e[w];
                    e[w] = T(-1);
                    st4(z + 4 * q, e[0], e[1], e[2], e[3]);
                }
            }
            tile_sync<G::LANES>();
            // ---- one rank-1 update of the whole lower triangle
            d = -z[N];
            dmin = dev_min(dmin, d);
            const T r = dev_rcp_fast<T>(d);
            PR x[H];
            T y[SC];
            #pragma unroll
            for (int g = 0; g < SR / 4; ++g) {
                T x0, x1, x2, x3;
                ld4(z + 4 * (G::rblock(g, 0) + ti), x0, x1, x2, x3);
                x[2 * g] = PR::make(x0, x1); x[2 * g + 1] = PR::make(x2, x3);
                x[2 * g].scale(r); x[2 * g + 1].scale(r);
            }
            #pragma unroll
            for (int h = 0; h < SC / 4; ++h)
                ld4(z + 4 * (G::cblock(h, 0) + tj), y[4 * h], y[4 * h + 1], y[4 * h + 2], y[4 * h + 3]);
            #pragma unroll
            for (int i = 0; i < H; ++i)
                #pragma unroll
                for (int c = 0; c < SC; ++c) {
                    if (G::cmin(c) > G::rmax(2 * i + 1)) continue;   // strictly upper for every thread
                    ap[i][c].fma_bcast(x[i], y[c]);
                }
        }
        tile_sync<G::LANES>();                                       // the lines are reused by the next matrix

        if (!valid) continue;
        T *__restrict__ dst = io.dst(m);
        const bool bad = !(dmin > T(0)) || !(d == d);
        int st = 0;
        if (bad) {                                                    // rare: LAPACK's natural-order index, dst as scratch
            if (lane == 0) { st = exact_potrf_info<T>(src, dst, N); if (st == 0) st = N; }
            st = __shfl_sync(__activemask(), st, (threadIdx.x & 31 & ~(G::LANES - 1)));
        }
        if (lane == 0 && info) info[m] = st;
        #pragma unroll
        for (int g = 0; g < SR / 4; ++g) {
            #pragma unroll
            for (int h = 0; h < SC / 4; ++h) {
                const int br = G::rblock(g, 0) + ti, bc = G::cblock(h, 0) + tj;
                if (bad) {                                            // flagged: this thread's natural blocks, all NaN
                    #pragma unroll
                    for (int v = 0; v < 4; ++v)
                        stg4(dst + (size_t)(4 * bc + v) * N + 4 * br, dev_nan<T>(), dev_nan<T>(), dev_nan<T>(), dev_nan<T>());
                    continue;
                }
                if (G::cblock(h, 0) > G::rblock(g, TR - 1)) continue; // strictly upper for every thread
                T b[4][4];                                            // b[row][col] of this block
                #pragma unroll
                for (int v = 0; v < 4; ++v) {
                    b[0][v] = ap[2 * g][4 * h + v].lo(); b[1][v] = ap[2 * g][4 * h + v].hi();
                    b[2][v] = ap[2 * g + 1][4 * h + v].lo(); b[3][v] = ap[2 * g + 1][4 * h + v].hi();
                }
                if (br == bc) {                                       // diagonal block: symmetrise in registers
                    #pragma unroll
                    for (int v = 0; v < 4; ++v)
                        stg4(dst + (size_t)(4 * bc + v) * N + 4 * br, v <= 0 ? b[0][v] : b[v][0], v <= 1 ? b[1][v] : b[v][1],
                             v <= 2 ? b[2][v] : b[v][2], b[3][v]);
                } else if (br > bc) {
                    #pragma unroll
                    for (int v = 0; v < 4; ++v)                       // natural position
                        stg4(dst + (size_t)(4 * bc + v) * N + 4 * br, b[0][v], b[1][v], b[2][v], b[3][v]);
                    #pragma unroll
                    for (int ww = 0; ww < 4; ++ww)                    // mirror image
                        stg4(dst + (size_t)(4 * br + ww) * N + 4 * bc, b[ww][0], b[ww][1], b[ww][2], b[ww][3]);
                }
            }
        }
    }
}

}  // namespace invgpu

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
//   * LOOK-AHEAD: right after the operands of pivot k are loaded, the column / row of the NEXT pivot
//     is brought up to date and published into the other line; only then comes the barrier and, behind
//     it, the bulk of the rank-1 update.  The publish -> barrier -> load latency of pivot k+1 hides
//     behind the FMAs of pivot k (BAR.SYNC is deferred-blocking: a warp runs on until it touches the
//     line), which is what the stall samples of the non-pipelined kernels asked for (barrier + LDS
//     + MUFU chain ~ 40 % of all samples).
//   * pivot order: blocks q are taken in ranges of PMIN = min(TR, TC) consecutive 4-blocks; within a
//     range the order is (w, t) -> pivot 4 (R PMIN + t) + w, so that consecutive pivots differ only in
//     WHO owns them and the loop over t can stay rolled: 8 NB / PMIN bodies of code for any N.  The
//     order is a symmetric permutation of the natural one, so A^-1 is unchanged (to rounding); `info`
//     of a flagged matrix is recomputed in natural order by one thread so that it is LAPACK's spotrf
//     info (reference src/inverse.c:92-95).
//
// Layout: TR x TC threads per matrix (any power-of-two grid: 4x2 lanes for n = 32, one warp as 8x4 for
// n = 64, a 128-thread CTA as 8x16 for n = 128, square grids for fp64), 4x4 sub-blocks dealt cyclically
// (row block b -> thread row b % TR, register group b / TR; likewise columns); blocks strictly above
// the diagonal for every thread are neither stored nor updated.
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
    // one FMUL2 by zero instead of two moves (a NaN survives, which only happens in flagged matrices)
    __device__ __forceinline__ void clear() { v = __fmul2_rn(v, make_float2(0.f, 0.f)); }
};
template <> struct Pair2<double> {
    double l, h;
    static __device__ __forceinline__ Pair2 make(double lo, double hi) { Pair2 p; p.l = lo; p.h = hi; return p; }
    __device__ __forceinline__ double lo() const { return l; }
    __device__ __forceinline__ double hi() const { return h; }
    __device__ __forceinline__ void fma_bcast(const Pair2 &x, double y) { l = fma(x.l, y, l); h = fma(x.h, y, h); }
    __device__ __forceinline__ void scale(double s) { l *= s; h *= s; }
    __device__ __forceinline__ void clear() { l = 0.0; h = 0.0; }
};

template <typename T> __device__ __forceinline__ T dev_min(T a, T b);
template <> __device__ __forceinline__ float dev_min<float>(float a, float b) { return fminf(a, b); }
template <> __device__ __forceinline__ double dev_min<double>(double a, double b) { return fmin(a, b); }
template <typename T> __device__ __forceinline__ T dev_rcp_fast(T x);
template <> __device__ __forceinline__ float dev_rcp_fast<float>(float x) {
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;   // one MUFU.RCP, <= 1 ulp
}
template <> __device__ __forceinline__ double dev_rcp_fast<double>(double x) { return 1.0 / x; }

// Shared-memory stores that ptxas must not merge: a 128-bit store wants its four registers adjacent
// and aligned, and the accumulators cannot be laid out to satisfy that for every column AND row slot
// -- merged stores cost four staging moves each (seen in the SASS).  A pair is stored as one 64-bit
// (fp64: 128-bit) access straight from its register pair, row elements as single words.
__device__ __forceinline__ void sts_pair(float *p, const Pair2<float> &v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "f"(v.lo()), "f"(v.hi()) : "memory");
}
__device__ __forceinline__ void sts_pair(double *p, const Pair2<double> &v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "d"(v.lo()), "d"(v.hi()) : "memory");
}
__device__ __forceinline__ void sts_one(float *p, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_one(double *p, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "d"(v) : "memory");
}

template <int N, int TR, int TC>
struct SweepGeo {
    static constexpr int SR = N / TR, SC = N / TC;                 // register tile
    static constexpr int H = SR / 2;                               // row pairs
    static constexpr int NGR = SR / 4, NGC = SC / 4;               // 4-groups per thread
    static constexpr int LANES = TR * TC;
    static constexpr int PMIN = TR < TC ? TR : TC;
    static constexpr int NB = N / 4;                               // 4-blocks per side
    static constexpr int NR = NB / PMIN;                           // ranges of PMIN consecutive blocks
    static_assert(N % (4 * TR) == 0 && N % (4 * TC) == 0, "N must be a multiple of 4 TR and 4 TC");
    static_assert(TR % PMIN == 0 && TC % PMIN == 0, "TR and TC must divide each other");
    static constexpr int LINE = N + 4;                             // z line + the pivot word (16-byte aligned)
    static constexpr int WORDS = ((2 * LINE + 31) / 32) * 32 + (LANES < 32 ? 8 : 0);
    static constexpr int BLOCK = LANES >= 64 ? LANES : INVGPU_WARP_TIER_BLOCK;
    static constexpr int MPB = BLOCK / LANES;
    // no thread holds a lower-triangle element in (row pair i, column c)
    __host__ __device__ static constexpr bool upper(int i, int c) { return TR * ((2 * i) / 4) + TR - 1 < TC * (c / 4); }
};

// element `odd` of a pair: read it and zero it
template <typename T, typename PR>
__device__ __forceinline__ T pair_take(PR &p, bool odd) {
    T e;
    if (!odd) { e = p.lo(); p = PR::make(T(0), p.hi()); }
    else { e = p.hi(); p = PR::make(p.lo(), T(0)); }
    return e;
}

// Publish pivot k' = 4 qn + WN (block qn = TR GRN + rkn = TC HCN + ckn) into line zn, raw, and restart
// the published slots:  zn[i] = T_ik' (i != k'),  zn[k'] = -1,  zn[N] = T_k'k' = -d.
template <typename T, int N, int TR, int TC, int GRN, int HCN, int WN>
__device__ __forceinline__ void sweep_publish(Pair2<T> (&ap)[N / TR / 2][N / TC], T *zn, int ti, int tj, int qn) {
    using SG = SweepGeo<N, TR, TC>;
    using PR = Pair2<T>;
    constexpr int srn = 4 * GRN + WN, scn = 4 * HCN + WN, ipn = srn / 2;
    const int rkn = qn % TR, ckn = qn % TC;
    if (tj == ckn) {                                               // column k': block rows below block qn
        #pragma unroll
        for (int g = GRN; g < SG::NGR; ++g) {
            if (g > GRN || ti > rkn) {
                sts_pair(zn + 4 * (TR * g + ti), ap[2 * g][scn]);
                sts_pair(zn + 4 * (TR * g + ti) + 2, ap[2 * g + 1][scn]);
                ap[2 * g][scn].clear();
                ap[2 * g + 1][scn].clear();
            }
        }
    }
    if (ti == rkn) {                                               // row k': block columns left of block qn
        #pragma unroll
        for (int h = 0; h <= HCN; ++h) {
            if (h < HCN || tj < ckn) {
                #pragma unroll
                for (int v = 0; v < 4; ++v)                        // scalar stores: no register-quad staging
                    sts_one(zn + 4 * (TC * h + tj) + v, pair_take<T>(ap[ipn][4 * h + v], srn % 2));
            }
        }
        if (tj == ckn) {                                           // diagonal thread: block qn = [row part | -1 | column part]
            T e[4];
            #pragma unroll
            for (int v = 0; v < 4; ++v) {
                if (v <= WN) e[v] = pair_take<T>(ap[ipn][4 * HCN + v], srn % 2);
                else e[v] = pair_take<T>(ap[(4 * GRN + v) / 2][scn], (4 * GRN + v) % 2);
            }
            sts_one(zn + N, e[WN]);
            e[WN] = T(-1);
            #pragma unroll
            for (int v = 0; v < 4; ++v) sts_one(zn + 4 * qn + v, e[v]);
        }
    }
}

// One pivot whose line zc is published and visible: load its operands, bring the slots of the NEXT
// pivot up to date and publish them into zn, barrier, then the bulk of the rank-1 update.
template <typename T, int N, int TR, int TC, int GRN, int HCN, int WN, bool HAS_NEXT>
__device__ __forceinline__ void sweep_step(Pair2<T> (&ap)[N / TR / 2][N / TC], const T *zc, T *zn, int ti, int tj, int qn,
                                           T &dmin, T &d) {
    using SG = SweepGeo<N, TR, TC>;
    using PR = Pair2<T>;
    constexpr int srn = 4 * GRN + WN, scn = 4 * HCN + WN, ipn = srn / 2;
    d = -zc[N];
    dmin = dev_min(dmin, d);
    const T r = dev_rcp_fast<T>(d);
    PR x[SG::H];
    T y[SG::SC];
    #pragma unroll
    for (int g = 0; g < SG::NGR; ++g) {
        T x0, x1, x2, x3;
        ld4(zc + 4 * (TR * g + ti), x0, x1, x2, x3);
        x[2 * g] = PR::make(x0, x1); x[2 * g + 1] = PR::make(x2, x3);
        x[2 * g].scale(r); x[2 * g + 1].scale(r);
    }
    #pragma unroll
    for (int h = 0; h < SG::NGC; ++h)
        ld4(zc + 4 * (TC * h + tj), y[4 * h], y[4 * h + 1], y[4 * h + 2], y[4 * h + 3]);
    if (HAS_NEXT) {
        #pragma unroll
        for (int i = 0; i < SG::H; ++i)                            // column slot of the next pivot
            if (!SG::upper(i, scn)) ap[i][scn].fma_bcast(x[i], y[scn]);
        #pragma unroll
        for (int c = 0; c < SG::SC; ++c)                           // row pair of the next pivot
            if (c != scn && !SG::upper(ipn, c)) ap[ipn][c].fma_bcast(x[ipn], y[c]);
        sweep_publish<T, N, TR, TC, GRN, HCN, WN>(ap, zn, ti, tj, qn);
        tile_sync<SG::LANES>();
    }
    #pragma unroll
    for (int i = 0; i < SG::H; ++i)
        #pragma unroll
        for (int c = 0; c < SG::SC; ++c) {
            if (SG::upper(i, c)) continue;
            if (HAS_NEXT && (c == scn || i == ipn)) continue;      // done before the publish
            ap[i][c].fma_bcast(x[i], y[c]);
        }
}

// all pivots of range R, sub-index W
template <typename T, int N, int TR, int TC, int R, int W, bool UNROLL>
__device__ __forceinline__ void sweep_range(Pair2<T> (&ap)[N / TR / 2][N / TC], T *sm, int ti, int tj, T &dmin, T &d) {
    using SG = SweepGeo<N, TR, TC>;
    constexpr int PM = SG::PMIN;
    constexpr int GR = (R * PM) / TR, HC = (R * PM) / TC;
    constexpr int J0 = (4 * R + W) * PM;                           // sequence number of the first pivot of this body
    // t = 0 .. PM-2: the next pivot has the same static slots
    if (UNROLL) {
        #pragma unroll
        for (int t = 0; t < PM - 1; ++t) {
            const int par = (J0 + t) & 1;
            sweep_step<T, N, TR, TC, GR, HC, W, true>(ap, sm + par * SG::LINE, sm + (par ^ 1) * SG::LINE, ti, tj, R * PM + t + 1, dmin, d);
        }
    } else {
        #pragma unroll 1
        for (int t = 0; t < PM - 1; ++t) {
            const int par = (J0 + t) & 1;
            sweep_step<T, N, TR, TC, GR, HC, W, true>(ap, sm + par * SG::LINE, sm + (par ^ 1) * SG::LINE, ti, tj, R * PM + t + 1, dmin, d);
        }
    }
    // t = PM-1: the next pivot opens the next body
    constexpr int par = (J0 + PM - 1) & 1;
    constexpr bool last = (R == SG::NR - 1) && (W == 3);
    constexpr int RN = (W == 3) ? R + 1 : R, WN = (W == 3) ? 0 : W + 1;
    constexpr int GRN = last ? 0 : (RN * PM) / TR, HCN = last ? 0 : (RN * PM) / TC;
    sweep_step<T, N, TR, TC, GRN, HCN, WN, !last>(ap, sm + par * SG::LINE, sm + (par ^ 1) * SG::LINE, ti, tj, RN * PM, dmin, d);
}

template <typename T, int N, int TR, int TC, bool UNROLL, int R>
struct SweepRanges {
    static __device__ __forceinline__ void run(Pair2<T> (&ap)[N / TR / 2][N / TC], T *sm, int ti, int tj, T &dmin, T &d) {
        sweep_range<T, N, TR, TC, R, 0, UNROLL>(ap, sm, ti, tj, dmin, d);
        sweep_range<T, N, TR, TC, R, 1, UNROLL>(ap, sm, ti, tj, dmin, d);
        sweep_range<T, N, TR, TC, R, 2, UNROLL>(ap, sm, ti, tj, dmin, d);
        sweep_range<T, N, TR, TC, R, 3, UNROLL>(ap, sm, ti, tj, dmin, d);
        if constexpr (R + 1 < SweepGeo<N, TR, TC>::NR) SweepRanges<T, N, TR, TC, UNROLL, R + 1>::run(ap, sm, ti, tj, dmin, d);
    }
};

template <typename T, int N, int TR, int TC, bool UNROLL, typename IO, int MINB>
__global__ void __launch_bounds__((SweepGeo<N, TR, TC>::BLOCK), MINB)
sweep_spd_kernel(IO io, i64 batch, int *__restrict__ info) {
    using SG = SweepGeo<N, TR, TC>;
    using PR = Pair2<T>;
    constexpr int SR = SG::SR, SC = SG::SC, H = SG::H;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);

    const int grp = threadIdx.x / SG::LANES;
    const int lane = threadIdx.x % SG::LANES;
    const int ti = lane / TC, tj = lane % TC;
    T *sm = smem + grp * SG::WORDS;

    #pragma unroll 1
    for (i64 base = (i64)blockIdx.x * SG::MPB; base < batch; base += (i64)gridDim.x * SG::MPB) {
        const i64 m = base + grp;
        const bool valid = m < batch;
        const T *__restrict__ src = io.src(valid ? m : batch - 1);

        PR ap[H][SC];                                              // ap[i][c] = T(rows 2i, 2i+1 ; column c), T = -A
        {
            T a[SR][SC];
            tile_load_upper<T, N, TR, TC, false>(a, src, ti, tj);
            #pragma unroll
            for (int i = 0; i < H; ++i)
                #pragma unroll
                for (int c = 0; c < SC; ++c) ap[i][c] = PR::make(-a[2 * i][c], -a[2 * i + 1][c]);
        }

        T dmin = T(1), d = T(1);                                   // smallest pivot so far / current pivot
        sweep_publish<T, N, TR, TC, 0, 0, 0>(ap, sm, ti, tj, 0);
        tile_sync<SG::LANES>();
        SweepRanges<T, N, TR, TC, UNROLL, 0>::run(ap, sm, ti, tj, dmin, d);

        if (!valid) continue;
        T *__restrict__ dst = io.dst(m);
        // a NaN pivot turns every later pivot into NaN, so the last one tells; otherwise the minimum does
        const bool bad = !(dmin > T(0)) || !(d == d);
        int st = 0;
        if (bad) {                                                 // rare: LAPACK's natural-order index, dst as scratch
            if (lane == 0) { st = exact_potrf_info<T>(src, dst, N); if (st == 0) st = N; }
            if (SG::LANES <= 32) st = __shfl_sync(__activemask(), st, (threadIdx.x & 31 & ~(SG::LANES - 1)));
            else { __syncthreads(); if (lane == 0) sm[0] = (T)st; __syncthreads(); st = (int)sm[0]; __syncthreads(); }
        }
        if (lane == 0 && info) info[m] = st;
        #pragma unroll
        for (int g = 0; g < SG::NGR; ++g) {
            #pragma unroll
            for (int h = 0; h < SG::NGC; ++h) {
                const int br = TR * g + ti, bc = TC * h + tj;
                if (bad) {                                         // flagged: this thread's natural blocks, all NaN
                    #pragma unroll
                    for (int v = 0; v < 4; ++v)
                        stg4(dst + (size_t)(4 * bc + v) * N + 4 * br, dev_nan<T>(), dev_nan<T>(), dev_nan<T>(), dev_nan<T>());
                    continue;
                }
                if (SG::upper(2 * g, 4 * h)) continue;             // strictly upper for every thread
                T b[4][4];                                         // b[row][col] of this block
                #pragma unroll
                for (int v = 0; v < 4; ++v) {
                    b[0][v] = ap[2 * g][4 * h + v].lo(); b[1][v] = ap[2 * g][4 * h + v].hi();
                    b[2][v] = ap[2 * g + 1][4 * h + v].lo(); b[3][v] = ap[2 * g + 1][4 * h + v].hi();
                }
                if (br == bc) {                                    // diagonal block: symmetrise in registers
                    #pragma unroll
                    for (int v = 0; v < 4; ++v)
                        stg4(dst + (size_t)(4 * bc + v) * N + 4 * br, v <= 0 ? b[0][v] : b[v][0], v <= 1 ? b[1][v] : b[v][1],
                             v <= 2 ? b[2][v] : b[v][2], b[3][v]);
                } else if (br > bc) {
                    #pragma unroll
                    for (int v = 0; v < 4; ++v)                    // natural position
                        stg4(dst + (size_t)(4 * bc + v) * N + 4 * br, b[0][v], b[1][v], b[2][v], b[3][v]);
                    #pragma unroll
                    for (int ww = 0; ww < 4; ++ww)                 // mirror image
                        stg4(dst + (size_t)(4 * br + ww) * N + 4 * bc, b[ww][0], b[ww][1], b[ww][2], b[ww][3]);
                }
            }
        }
    }
}

}  // namespace invgpu

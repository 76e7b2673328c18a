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

// Measured on B200 (tools/variant.sh A/B/C, n = 16 / 32 fp32): interleaving the lanes of the matrices that
// share a warp (column owners in one quarter-warp: fewer shared-memory store wavefronts) costs 15-17 %
// because the global loads / stores lose their 32-byte runs; re-aligning the warps of a CTA for
// instruction-cache sharing is neutral (n = 32) to negative (n = 16).  Both stay off.
#ifndef INVGPU_SWEEP_INTERLEAVE
#define INVGPU_SWEEP_INTERLEAVE 0
#endif
#ifndef INVGPU_SWEEP_LOCKSTEP
#define INVGPU_SWEEP_LOCKSTEP 0
#endif

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

enum { SWEEP_INVERSE = 0, SWEEP_GP = 1 };

template <int N, int TR, int TC>
struct SweepGeo {
    static constexpr int SR = N / TR, SC = N / TC;                 // register tile
    static constexpr int H = SR / 2;                               // row pairs
    static constexpr int NGR = SR / 4, NGC = SC / 4;               // 4-groups per thread
    static constexpr int LANES = TR * TC;
    static constexpr int PMIN = TR < TC ? TR : TC;
    static constexpr int PMAX = TR < TC ? TC : TR;
    static constexpr int NB = N / 4;                               // 4-blocks per side
    static constexpr int NS = NB / PMAX;                           // super-ranges of PMAX consecutive blocks
    static constexpr int SUBS = PMAX / PMIN;                       // ranges of PMIN blocks per super-range
    static constexpr int RG = PMAX / TR, CG = PMAX / TC;           // register groups per super-range
    static_assert(N % (4 * TR) == 0 && N % (4 * TC) == 0, "N must be a multiple of 4 TR and 4 TC");
    static_assert(PMAX % PMIN == 0, "TR and TC must divide each other");
    static constexpr int LINE = N + 4;                             // z line + pivot word + two right-hand-side words
    static constexpr int WORDS = ((2 * LINE + 31) / 32) * 32 + (LANES < 32 ? 8 : 0);
    static constexpr int BLOCK = LANES >= 64 ? LANES : INVGPU_WARP_TIER_BLOCK;
    static constexpr int MPB = BLOCK / LANES;
    // no thread holds a lower-triangle element in (row pair i, column c)
    __host__ __device__ static constexpr bool upper(int i, int c) { return TR * ((2 * i) / 4) + TR - 1 < TC * (c / 4); }
    // Pivot order: for S (super-range), for W (index within the 4-block), for SUB, for t:
    //   block q = S PMAX + SUB PMIN + t,  pivot 4 q + W.
    // After level (S, W) every thread has eliminated its row / column slots (g, W), g in super-range S:
    // for the factor-only (GP) mode rows and columns die evenly over the threads and are pruned statically.
    __host__ __device__ static constexpr bool row_dead(int slot, int S, int W) {
        return (slot / 4) / RG < S || ((slot / 4) / RG == S && slot % 4 < W);
    }
    __host__ __device__ static constexpr bool col_dead(int slot, int S, int W) {
        return (slot / 4) / CG < S || ((slot / 4) / CG == S && slot % 4 < W);
    }
    template <int MODE>
    __host__ __device__ static constexpr bool skip(int i, int c, int S, int W) {
        return upper(i, c) || (MODE == SWEEP_GP && (row_dead(2 * i + 1, S, W) || col_dead(c, S, W)));
    }
    // right-hand sides of the GP mode: combination cb = e H + ip (vector e, row pair ip) lives in lane cb % TC
    static constexpr int NC = (2 * H + TC - 1) / TC;
};

// element `odd` of a pair: read it and (inverse mode) zero it
template <typename T, int MODE, typename PR>
__device__ __forceinline__ T pair_take(PR &p, bool odd) {
    T e;
    if (!odd) { e = p.lo(); if (MODE == SWEEP_INVERSE) p = PR::make(T(0), p.hi()); }
    else { e = p.hi(); if (MODE == SWEEP_INVERSE) p = PR::make(p.lo(), T(0)); }
    return e;
}

// Publish pivot k' = 4 qn + WN (block qn = TR GRN + rkn = TC HCN + ckn) into line zn, raw, and (inverse
// mode) restart the published slots:  zn[i] = T_ik' (i != k'),  zn[k'] = -1,  zn[N] = T_k'k' = -d;
// GP mode: zn[N+1], zn[N+2] = the two right-hand sides at row k'.
template <typename T, int N, int TR, int TC, int MODE, int GRN, int HCN, int WN>
__device__ __forceinline__ void sweep_publish(Pair2<T> (&ap)[N / TR / 2][N / TC], Pair2<T> (&rhs)[SweepGeo<N, TR, TC>::NC],
                                              T *zn, int ti, int tj, int qn) {
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
                if (MODE == SWEEP_INVERSE) { ap[2 * g][scn].clear(); ap[2 * g + 1][scn].clear(); }
            }
        }
    }
    if (ti == rkn) {                                               // row k': block columns left of block qn
        #pragma unroll
        for (int h = 0; h <= HCN; ++h) {
            if (h < HCN || tj < ckn) {
                #pragma unroll
                for (int v = 0; v < 4; ++v)
                    sts_one(zn + 4 * (TC * h + tj) + v, pair_take<T, MODE>(ap[ipn][4 * h + v], srn % 2));
            }
        }
        if (tj == ckn) {                                           // diagonal thread: block qn = [row part | -1 | column part]
            T e[4];
            #pragma unroll
            for (int v = 0; v < 4; ++v) {
                if (v <= WN) e[v] = pair_take<T, MODE>(ap[ipn][4 * HCN + v], srn % 2);
                else e[v] = pair_take<T, MODE>(ap[(4 * GRN + v) / 2][scn], (4 * GRN + v) % 2);
            }
            sts_one(zn + N, e[WN]);
            e[WN] = T(-1);
            #pragma unroll
            for (int v = 0; v < 4; ++v) sts_one(zn + 4 * qn + v, e[v]);
        }
        if (MODE == SWEEP_GP) {                                    // the lanes holding row k' of the right-hand sides
            #pragma unroll
            for (int e = 0; e < 2; ++e) {
                constexpr int dummy = 0; (void)dummy;
                const int cb = e * SG::H + ipn;                    // static
                if (tj == cb % TC) sts_one(zn + N + 1 + e, (srn % 2) ? rhs[cb / TC].hi() : rhs[cb / TC].lo());
            }
        }
    }
}

// One pivot whose line zc is published and visible: load its operands, bring the slots of the NEXT
// pivot up to date and publish them into zn, barrier, then the bulk of the rank-1 update.
// (S, W): elimination level of THIS pivot (static pruning in GP mode).
template <typename T, int N, int TR, int TC, int MODE, int S, int W, int GRN, int HCN, int WN, bool HAS_NEXT>
__device__ __forceinline__ void sweep_step(Pair2<T> (&ap)[N / TR / 2][N / TC], Pair2<T> (&rhs)[SweepGeo<N, TR, TC>::NC],
                                           const T *zc, T *zn, int ti, int tj, int qn, T &dmin, T &d, T &acc_m, T &acc_q) {
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
        if (MODE == SWEEP_GP && SG::row_dead(4 * g + 3, S, W)) continue;     // the whole group is dead
        T x0, x1, x2, x3;
        ld4(zc + 4 * (TR * g + ti), x0, x1, x2, x3);
        x[2 * g] = PR::make(x0, x1); x[2 * g + 1] = PR::make(x2, x3);
        x[2 * g].scale(r); x[2 * g + 1].scale(r);
    }
    #pragma unroll
    for (int h = 0; h < SG::NGC; ++h) {
        if (MODE == SWEEP_GP && SG::col_dead(4 * h + 3, S, W)) continue;
        ld4(zc + 4 * (TC * h + tj), y[4 * h], y[4 * h + 1], y[4 * h + 2], y[4 * h + 3]);
    }
    if (MODE == SWEEP_GP) {                                        // right-hand sides: u_i += x_i u_k, and the two scalars
        const T uk = zc[N + 1], vk = zc[N + 2];
        const T ur = uk * r;
        acc_m = fma(ur, vk, acc_m);
        acc_q = fma(ur, uk, acc_q);
        #pragma unroll
        for (int j = 0; j < SG::NC; ++j) {
            const int cb = tj + TC * j;                            // vector cb / H, row pair cb % H of this thread's rows
            const int ip = cb % SG::H;
            const T *zp = zc + 4 * (TR * (ip / 2) + ti) + 2 * (ip % 2);
            PR xs = PR::make(zp[0], zp[1]);
            xs.scale(r);
            rhs[j].fma_bcast(xs, cb / SG::H == 0 ? uk : vk);
        }
    }
    if (HAS_NEXT) {
        #pragma unroll
        for (int i = 0; i < SG::H; ++i)                            // column slot of the next pivot
            if (!SG::template skip<MODE>(i, scn, S, W)) ap[i][scn].fma_bcast(x[i], y[scn]);
        #pragma unroll
        for (int c = 0; c < SG::SC; ++c)                           // row pair of the next pivot
            if (c != scn && !SG::template skip<MODE>(ipn, c, S, W)) ap[ipn][c].fma_bcast(x[ipn], y[c]);
        sweep_publish<T, N, TR, TC, MODE, GRN, HCN, WN>(ap, rhs, zn, ti, tj, qn);
        tile_sync<SG::LANES>();
    }
    #pragma unroll
    for (int i = 0; i < SG::H; ++i)
        #pragma unroll
        for (int c = 0; c < SG::SC; ++c) {
            if (SG::template skip<MODE>(i, c, S, W)) continue;
            if (HAS_NEXT && (c == scn || i == ipn)) continue;      // done before the publish
            ap[i][c].fma_bcast(x[i], y[c]);
        }
}

// all pivots of level (S, W), range SUB
template <typename T, int N, int TR, int TC, int MODE, bool UNROLL, int S, int W, int SUB>
__device__ __forceinline__ void sweep_range(Pair2<T> (&ap)[N / TR / 2][N / TC], Pair2<T> (&rhs)[SweepGeo<N, TR, TC>::NC],
                                            T *sm, int ti, int tj, T &dmin, T &d, T &acc_m, T &acc_q) {
    using SG = SweepGeo<N, TR, TC>;
    constexpr int PM = SG::PMIN;
    constexpr int Q0 = S * SG::PMAX + SUB * PM;                    // first block of this range
    constexpr int GR = Q0 / TR, HC = Q0 / TC;
    constexpr int J0 = ((4 * S + W) * SG::SUBS + SUB) * PM;        // sequence number of the first pivot of this body
    // warp tiers: optionally re-align the warps of the CTA so that they share instruction fetches
    if (SG::LANES <= 32 && INVGPU_SWEEP_LOCKSTEP > 0 && J0 % (INVGPU_SWEEP_LOCKSTEP > 0 ? INVGPU_SWEEP_LOCKSTEP : 1) == 0) __syncthreads();
    // t = 0 .. PM-2: the next pivot has the same static slots
    if (UNROLL) {
        #pragma unroll
        for (int t = 0; t < PM - 1; ++t) {
            const int par = (J0 + t) & 1;
            sweep_step<T, N, TR, TC, MODE, S, W, GR, HC, W, true>(ap, rhs, sm + par * SG::LINE, sm + (par ^ 1) * SG::LINE, ti, tj,
                                                                  Q0 + t + 1, dmin, d, acc_m, acc_q);
        }
    } else {
        #pragma unroll 1
        for (int t = 0; t < PM - 1; ++t) {
            const int par = (J0 + t) & 1;
            sweep_step<T, N, TR, TC, MODE, S, W, GR, HC, W, true>(ap, rhs, sm + par * SG::LINE, sm + (par ^ 1) * SG::LINE, ti, tj,
                                                                  Q0 + t + 1, dmin, d, acc_m, acc_q);
        }
    }
    // t = PM-1: the next pivot opens the next body
    constexpr int par = (J0 + PM - 1) & 1;
    constexpr bool last = (S == SG::NS - 1) && (W == 3) && (SUB == SG::SUBS - 1);
    constexpr int SUBN = (SUB + 1 < SG::SUBS) ? SUB + 1 : 0;
    constexpr int WN = (SUBN != 0) ? W : (W == 3 ? 0 : W + 1);
    constexpr int SN = (SUBN != 0 || W != 3) ? S : S + 1;
    constexpr int QN = last ? 0 : SN * SG::PMAX + SUBN * PM;
    sweep_step<T, N, TR, TC, MODE, S, W, QN / TR, QN / TC, WN, !last>(ap, rhs, sm + par * SG::LINE, sm + (par ^ 1) * SG::LINE, ti, tj,
                                                                     QN, dmin, d, acc_m, acc_q);
}

template <typename T, int N, int TR, int TC, int MODE, bool UNROLL, int S, int W, int SUB>
struct SweepRanges {
    static __device__ __forceinline__ void run(Pair2<T> (&ap)[N / TR / 2][N / TC], Pair2<T> (&rhs)[SweepGeo<N, TR, TC>::NC],
                                               T *sm, int ti, int tj, T &dmin, T &d, T &acc_m, T &acc_q) {
        using SG = SweepGeo<N, TR, TC>;
        sweep_range<T, N, TR, TC, MODE, UNROLL, S, W, SUB>(ap, rhs, sm, ti, tj, dmin, d, acc_m, acc_q);
        constexpr int SUBN = (SUB + 1 < SG::SUBS) ? SUB + 1 : 0;
        constexpr int WN = (SUBN != 0) ? W : (W == 3 ? 0 : W + 1);
        constexpr int SN = (SUBN != 0 || W != 3) ? S : S + 1;
        if constexpr (SN < SG::NS) SweepRanges<T, N, TR, TC, MODE, UNROLL, SN, WN, SUBN>::run(ap, rhs, sm, ti, tj, dmin, d, acc_m, acc_q);
    }
};

// Padded tiers (PadIO): the logical-lower positions of the tile from the UPPER triangle of a column-major
// n x n matrix (lda = n, n <= N), identity outside; and the store of the leading n x n part of the result.
template <typename T, int N, int TR, int TC>
__device__ __forceinline__ void tile_load_upper_padded(T (&a)[N / TR][N / TC], const T *__restrict__ src, int n, int ti, int tj) {
    using SG = SweepGeo<N, TR, TC>;
    #pragma unroll
    for (int g = 0; g < SG::NGR; ++g)
        #pragma unroll
        for (int h = 0; h < SG::NGC; ++h) {
            const int br = TR * g + ti, bc = TC * h + tj;
            #pragma unroll
            for (int w = 0; w < 4; ++w)
                #pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int r = 4 * br + w, c = 4 * bc + v;
                    T e = T(0);
                    if (!SG::upper(2 * g, 4 * h) && r >= c) {
                        if (r < n) e = src[(size_t)r * n + c];      // element (row c, column r) of the upper triangle
                        else if (r == c) e = T(1);
                    }
                    a[4 * g + w][4 * h + v] = e;
                }
        }
}

template <typename T, int N, int TR, int TC, bool UNROLL, typename IO, int MINB>
__global__ void __launch_bounds__((SweepGeo<N, TR, TC>::BLOCK), MINB)
sweep_spd_kernel(IO io, i64 batch, int *__restrict__ info) {
    using SG = SweepGeo<N, TR, TC>;
    using PR = Pair2<T>;
    constexpr int SR = SG::SR, SC = SG::SC, H = SG::H;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);

    // Lane map.  Groups smaller than a warp are INTERLEAVED: lane = ti + TR (mat + MPW tj), so that the
    // owners of a pivot column (tj == const) of all the matrices of a warp sit in one quarter-warp and
    // their 128-bit stores cost one shared-memory wavefront instead of four.
    int grp, ti, tj;
    if (SG::LANES < 32 && INVGPU_SWEEP_INTERLEAVE) {
        constexpr int MPW = 32 / SG::LANES;
        const int wl = threadIdx.x & 31;
        ti = wl % TR; tj = wl / (TR * MPW);
        grp = (threadIdx.x >> 5) * MPW + (wl / TR) % MPW;
    } else {
        grp = threadIdx.x / SG::LANES;
        ti = (threadIdx.x % SG::LANES) / TC; tj = threadIdx.x % TC;
    }
    const bool lead = (ti == 0 && tj == 0);
    T *sm = smem + grp * SG::WORDS;

    #pragma unroll 1
    for (i64 base = (i64)blockIdx.x * SG::MPB; base < batch; base += (i64)gridDim.x * SG::MPB) {
        const i64 m = base + grp;
        const bool valid = m < batch;
        const T *__restrict__ src;
        T *__restrict__ dst;
        int nn = N;                                                // order of the matrix inside the N x N tile
        i64 iidx = m;                                              // where its info goes
        if constexpr (IoTraits<IO>::PADDED) {
            const MixedItem it = io.items[valid ? m : batch - 1];
            src = static_cast<const T *>(it.in); dst = static_cast<T *>(it.out); nn = it.n; iidx = it.index;
        } else {
            src = io.src(valid ? m : batch - 1); dst = io.dst(valid ? m : batch - 1);
        }

        PR ap[H][SC];                                              // ap[i][c] = T(rows 2i, 2i+1 ; column c), T = -A
        {
            T a[SR][SC];
            if constexpr (IoTraits<IO>::PADDED) tile_load_upper_padded<T, N, TR, TC>(a, src, nn, ti, tj);
            else tile_load_upper<T, N, TR, TC, false>(a, src, ti, tj);
            #pragma unroll
            for (int i = 0; i < H; ++i)
                #pragma unroll
                for (int c = 0; c < SC; ++c) ap[i][c] = PR::make(-a[2 * i][c], -a[2 * i + 1][c]);
        }

        T dmin = T(1), d = T(1);                                   // smallest pivot so far / current pivot
        T acc_m = T(0), acc_q = T(0);
        PR rhs[SG::NC];                                            // unused in this mode
        sweep_publish<T, N, TR, TC, SWEEP_INVERSE, 0, 0, 0>(ap, rhs, sm, ti, tj, 0);
        tile_sync<SG::LANES>();
        SweepRanges<T, N, TR, TC, SWEEP_INVERSE, UNROLL, 0, 0, 0>::run(ap, rhs, sm, ti, tj, dmin, d, acc_m, acc_q);

        // a NaN pivot turns every later pivot into NaN, so the last one tells; otherwise the minimum does
        const bool bad = valid && (!(dmin > T(0)) || !(d == d));
        int st = 0;
        if (SG::LANES <= 32) {                                     // rare: LAPACK's natural-order index, dst as scratch
            if (__any_sync(0xffffffffu, bad)) {
                if (bad && lead) { st = exact_potrf_info<T>(src, dst, nn); if (st == 0) st = nn; }
                __syncwarp();                                      // scratch use ends before the NaN fill
            }
        } else if (bad) {                                          // one matrix per CTA: uniform
            if (lead) { st = exact_potrf_info<T>(src, dst, nn); if (st == 0) st = nn; }
            __syncthreads();
        }
        if (!valid) continue;
        if (lead && info) info[iidx] = st;
        if constexpr (IoTraits<IO>::PADDED) {                      // bounds-checked scalar stores, lda = nn
            #pragma unroll
            for (int g = 0; g < SG::NGR; ++g) {
                #pragma unroll
                for (int h = 0; h < SG::NGC; ++h) {
                    const int br = TR * g + ti, bc = TC * h + tj;
                    if (bad) {
                        #pragma unroll
                        for (int w = 0; w < 4; ++w)
                            #pragma unroll
                            for (int v = 0; v < 4; ++v)
                                if (4 * br + w < nn && 4 * bc + v < nn) dst[(size_t)(4 * bc + v) * nn + 4 * br + w] = dev_nan<T>();
                        continue;
                    }
                    if (SG::upper(2 * g, 4 * h) || br < bc) continue;
                    #pragma unroll
                    for (int w = 0; w < 4; ++w)
                        #pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const int r = 4 * br + w, c = 4 * bc + v;
                            if (r >= nn || c >= nn) continue;
                            // static register indices only: element (w, v) and, on a diagonal block, its mirror (v, w)
                            const T e_wv = ((4 * g + w) % 2) ? ap[(4 * g + w) / 2][4 * h + v].hi() : ap[(4 * g + w) / 2][4 * h + v].lo();
                            const T e_vw = ((4 * g + v) % 2) ? ap[(4 * g + v) / 2][4 * h + w].hi() : ap[(4 * g + v) / 2][4 * h + w].lo();
                            const T e = (br == bc && v > w) ? e_vw : e_wv;
                            dst[(size_t)c * nn + r] = e;
                            if (br > bc) dst[(size_t)r * nn + c] = e;
                        }
                }
            }
            continue;
        }
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

// ------------------------------------------------------------------------------------------
// Fused GP mean / variance (reference src/gauss_bench.cu:127-265, 275-409 in ONE launch) on the same
// machinery: B is loaded with diag(C) added, only the factor part of the sweep is executed (dead rows
// and columns pruned statically), the two right-hand sides ride along distributed over the thread
// columns, and with A = L D L^T the scalars are  sum_k u_k v_k / d_k  and  E - sum_k u_k^2 / d_k.
// Nothing but the scalars is written.  scratch: N*N words per matrix slot, used only to recompute
// LAPACK's natural-order info for a flagged (non-SPD) matrix.
// ------------------------------------------------------------------------------------------
template <typename T, int N, int TR, int TC, bool UNROLL, int MINB>
__global__ void __launch_bounds__((SweepGeo<N, TR, TC>::BLOCK), MINB)
sweep_gp_kernel(GpIO<T> io, i64 batch, int *__restrict__ info, T *__restrict__ scratch) {
    using SG = SweepGeo<N, TR, TC>;
    using PR = Pair2<T>;
    constexpr int SR = SG::SR, SC = SG::SC, H = SG::H;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);

    const int grp = threadIdx.x / SG::LANES;
    const int ti = (threadIdx.x % SG::LANES) / TC, tj = threadIdx.x % TC;
    const bool lead = (ti == 0 && tj == 0);
    T *sm = smem + grp * SG::WORDS;

    #pragma unroll 1
    for (i64 base = (i64)blockIdx.x * SG::MPB; base < batch; base += (i64)gridDim.x * SG::MPB) {
        const i64 m = base + grp;
        const bool valid = m < batch;
        const i64 mm = valid ? m : batch - 1;
        const T *__restrict__ src = io.b + mm * (i64)(N * N);
        const T *__restrict__ cv = io.c + mm * N;

        PR ap[H][SC];
        {
            T a[SR][SC];
            tile_load_upper<T, N, TR, TC, false>(a, src, ti, tj);
            // + diag(C) on load: only threads holding a diagonal sub-block see diagonal elements
            #pragma unroll
            for (int g = 0; g < SG::NGR; ++g)
                #pragma unroll
                for (int h = 0; h < SG::NGC; ++h) {
                    if (TR * g > TC * h + TC - 1 || TR * g + TR - 1 < TC * h) continue;
                    if (TR * g + ti == TC * h + tj) {
                        T c0, c1, c2, c3;
                        ldg4(cv + 4 * (TR * g + ti), c0, c1, c2, c3);
                        a[4 * g][4 * h] += c0; a[4 * g + 1][4 * h + 1] += c1;
                        a[4 * g + 2][4 * h + 2] += c2; a[4 * g + 3][4 * h + 3] += c3;
                    }
                }
            #pragma unroll
            for (int i = 0; i < H; ++i)
                #pragma unroll
                for (int c = 0; c < SC; ++c) ap[i][c] = PR::make(-a[2 * i][c], -a[2 * i + 1][c]);
        }
        PR rhs[SG::NC];
        {
            const T *__restrict__ av = io.a + mm * N;
            const T *__restrict__ dv = (io.d ? io.d : io.a) + mm * N;
            #pragma unroll
            for (int j = 0; j < SG::NC; ++j) {
                const int cb = tj + TC * j, ip = cb % H;
                const T *p = (cb / H == 0 ? av : dv) + 4 * (TR * (ip / 2) + ti) + 2 * (ip % 2);
                rhs[j] = PR::make(p[0], p[1]);                     // lanes beyond 2 H combinations hold copies nobody reads
            }
        }

        T dmin = T(1), d = T(1), acc_m = T(0), acc_q = T(0);
        sweep_publish<T, N, TR, TC, SWEEP_GP, 0, 0, 0>(ap, rhs, sm, ti, tj, 0);
        tile_sync<SG::LANES>();
        SweepRanges<T, N, TR, TC, SWEEP_GP, UNROLL, 0, 0, 0>::run(ap, rhs, sm, ti, tj, dmin, d, acc_m, acc_q);

        const bool bad = valid && (!(dmin > T(0)) || !(d == d));
        if (!valid || !lead) continue;
        if (bad) {                                    // rare: report LAPACK's natural-order index
            T *w = scratch + ((i64)blockIdx.x * SG::MPB + grp) * (i64)(N * N);
            for (int j = 0; j < N; ++j) {
                for (int i = 0; i < j; ++i) w[(size_t)j * N + i] = src[(size_t)j * N + i];
                w[(size_t)j * N + j] = src[(size_t)j * N + j] + cv[j];
            }
            int st = exact_potrf_info_inplace<T>(w, N);
            if (st == 0) st = N;
            if (io.means) io.means[m] = dev_nan<T>();
            if (io.variances) io.variances[m] = dev_nan<T>();
            if (info) info[m] = st;
            continue;
        }
        if (io.means) io.means[m] = acc_m;
        if (io.variances) io.variances[m] = io.e[m] - acc_q;
        if (info) info[m] = 0;
    }
}

}  // namespace invgpu

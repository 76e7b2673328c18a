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
    // ---- 2x2 block pivots (sweep2_*): two z lines, three pivot words (+1 pad), four right-hand-side words
    static constexpr int LINE2 = 2 * N + 8;
    static constexpr int WORDS2 = ((2 * LINE2 + 31) / 32) * 32 + (LANES < 32 ? 8 : 0);
    // level (S, WP): row pairs / columns whose pivots are done for every thread (factor-only pruning)
    __host__ __device__ static constexpr bool pair_dead2(int i, int S, int WP) {
        return ((2 * i) / 4) / RG < S || (((2 * i) / 4) / RG == S && ((2 * i) % 4) / 2 < WP);
    }
    __host__ __device__ static constexpr bool col_dead2(int c, int S, int WP) {
        return (c / 4) / CG < S || ((c / 4) / CG == S && (c % 4) / 2 < WP);
    }
    template <int MODE>
    __host__ __device__ static constexpr bool skip2(int i, int c, int S, int WP) {
        return upper(i, c) || (MODE == SWEEP_GP && (pair_dead2(i, S, WP) || col_dead2(c, S, WP)));
    }
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

// ==========================================================================================
// 2x2 BLOCK PIVOTS.  The pivots 4q+2wp, 4q+2wp+1 (one Pair2 row of the diagonal thread) are taken
// together: with Z = [z1 z2] = T(:, K) raw, D = -T(K, K) and W = D^-1 the sweep of both is
//     T_ic += (Z W)_i . Z_c ,   Z_K := -I   (the block form of the z_k = -1 convention),
// i.e. one publish, one barrier, one reciprocal and one dependency chain per TWO pivots; the operands
// X = Z W cost four packed multiplies per row pair instead of two.  The n = 64 / 128 tiers and the GP
// kernel are bound by exactly that chain (profiles/r1_sweep64_summary.md), not by issue slots or pipes.
// Line layout: z1[N] | z2[N] | p11 p21 p22 - | u1 v1 u2 v2   (p = raw T entries, u / v = right-hand sides).
// Levels are (S, WP), WP = 0, 1: whole register pairs die together in the factor-only mode.
// ==========================================================================================
template <typename T, int N, int TR, int TC, int MODE, int GRN, int HCN, int WPN>
__device__ __forceinline__ void sweep2_publish(Pair2<T> (&ap)[N / TR / 2][N / TC], Pair2<T> (&rhs)[SweepGeo<N, TR, TC>::NC],
                                               T *zn, int ti, int tj, int qn) {
    using SG = SweepGeo<N, TR, TC>;
    using PR = Pair2<T>;
    constexpr int sc1 = 4 * HCN + 2 * WPN, sc2 = sc1 + 1, ipn = (4 * GRN + 2 * WPN) / 2;
    T *z1 = zn, *z2 = zn + N, *pw = zn + 2 * N;
    const int rkn = qn % TR, ckn = qn % TC;
    if (tj == ckn) {                                               // columns k1, k2: block rows below block qn
        #pragma unroll
        for (int g = GRN; g < SG::NGR; ++g) {
            if (g > GRN || ti > rkn) {
                const int o = 4 * (TR * g + ti);
                sts_pair(z1 + o, ap[2 * g][sc1]); sts_pair(z1 + o + 2, ap[2 * g + 1][sc1]);
                sts_pair(z2 + o, ap[2 * g][sc2]); sts_pair(z2 + o + 2, ap[2 * g + 1][sc2]);
                if (MODE == SWEEP_INVERSE) { ap[2 * g][sc1].clear(); ap[2 * g + 1][sc1].clear(); ap[2 * g][sc2].clear(); ap[2 * g + 1][sc2].clear(); }
            }
        }
    }
    if (ti == rkn) {                                               // rows k1, k2 (one register pair): block columns left of qn
        #pragma unroll
        for (int h = 0; h <= HCN; ++h) {
            if (h < HCN || tj < ckn) {
                #pragma unroll
                for (int v = 0; v < 4; ++v) {
                    PR &p = ap[ipn][4 * h + v];
                    sts_one(z1 + 4 * (TC * h + tj) + v, p.lo());
                    sts_one(z2 + 4 * (TC * h + tj) + v, p.hi());
                    if (MODE == SWEEP_INVERSE) p.clear();
                }
            }
        }
        if (tj == ckn) {                                           // diagonal thread: block qn of both lines, Z_K = -I
            constexpr int a = 2 * WPN, b = a + 1;
            T e1[4], e2[4];
            #pragma unroll
            for (int v = 0; v < 4; ++v) {
                if (v < a) {                                       // (rows a, b ; column v)
                    PR &p = ap[ipn][4 * HCN + v];
                    e1[v] = p.lo(); e2[v] = p.hi();
                    if (MODE == SWEEP_INVERSE) p.clear();
                } else if (v == a) {                               // T_aa, T_ba
                    PR &p = ap[ipn][sc1];
                    sts_one(pw + 0, p.lo()); sts_one(pw + 1, p.hi());
                    e1[v] = T(-1); e2[v] = T(0);
                    if (MODE == SWEEP_INVERSE) p.clear();
                } else if (v == b) {                               // T_bb (the lo half is above the diagonal: unused)
                    PR &p = ap[ipn][sc2];
                    sts_one(pw + 2, p.hi());
                    e1[v] = T(0); e2[v] = T(-1);
                    if (MODE == SWEEP_INVERSE) p.clear();
                } else {                                           // WPN == 0, v = 2, 3: rows of the next pair, columns a, b
                    const PR &p1 = ap[2 * GRN + 1][sc1], &p2 = ap[2 * GRN + 1][sc2];
                    e1[v] = (v == 2) ? p1.lo() : p1.hi();
                    e2[v] = (v == 2) ? p2.lo() : p2.hi();
                }
            }
            if (WPN == 0 && MODE == SWEEP_INVERSE) { ap[2 * GRN + 1][sc1].clear(); ap[2 * GRN + 1][sc2].clear(); }
            #pragma unroll
            for (int v = 0; v < 4; ++v) { sts_one(z1 + 4 * qn + v, e1[v]); sts_one(z2 + 4 * qn + v, e2[v]); }
        }
        if (MODE == SWEEP_GP) {                                    // the lanes holding rows k1, k2 of the right-hand sides
            #pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int cb = e * SG::H + ipn;                    // static
                if (tj == cb % TC) { sts_one(pw + 4 + e, rhs[cb / TC].lo()); sts_one(pw + 6 + e, rhs[cb / TC].hi()); }
            }
        }
    }
}

template <typename T, int N, int TR, int TC, int MODE, int S, int WP, int GRN, int HCN, int WPN, bool HAS_NEXT>
__device__ __forceinline__ void sweep2_step(Pair2<T> (&ap)[N / TR / 2][N / TC], Pair2<T> (&rhs)[SweepGeo<N, TR, TC>::NC],
                                            const T *zc, T *zn, int ti, int tj, int qn, T &dmin, T &d, T &acc_m, T &acc_q) {
    using SG = SweepGeo<N, TR, TC>;
    using PR = Pair2<T>;
    constexpr int sc1n = 4 * HCN + 2 * WPN, sc2n = sc1n + 1, ipn = (4 * GRN + 2 * WPN) / 2;
    const T *z1 = zc, *z2 = zc + N, *pw = zc + 2 * N;
    const T d1 = -pw[0], o = -pw[1], d2 = -pw[2];
    const T det = fma(d1, d2, -o * o);                             // d1 > 0 and det > 0  <=>  both pivots positive
    d = det;
    dmin = dev_min(dmin, dev_min(d1, det));
    const T rdet = dev_rcp_fast<T>(det);
    const T w11 = d2 * rdet, w12 = -o * rdet, w22 = d1 * rdet;     // W = D^-1
    PR x1[SG::H], x2[SG::H];
    T y1[SG::SC], y2[SG::SC];
    #pragma unroll
    for (int g = 0; g < SG::NGR; ++g) {
        if (MODE == SWEEP_GP && SG::pair_dead2(2 * g + 1, S, WP)) continue;   // the whole group is dead
        T a0, a1, a2, a3, b0, b1, b2, b3;
        ld4(z1 + 4 * (TR * g + ti), a0, a1, a2, a3);
        ld4(z2 + 4 * (TR * g + ti), b0, b1, b2, b3);
        #pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const PR za = hh ? PR::make(a2, a3) : PR::make(a0, a1);
            const PR zb = hh ? PR::make(b2, b3) : PR::make(b0, b1);
            PR u = za; u.scale(w11); u.fma_bcast(zb, w12); x1[2 * g + hh] = u;
            PR v = za; v.scale(w12); v.fma_bcast(zb, w22); x2[2 * g + hh] = v;
        }
    }
    #pragma unroll
    for (int h = 0; h < SG::NGC; ++h) {
        if (MODE == SWEEP_GP && SG::col_dead2(4 * h + 3, S, WP)) continue;
        ld4(z1 + 4 * (TC * h + tj), y1[4 * h], y1[4 * h + 1], y1[4 * h + 2], y1[4 * h + 3]);
        ld4(z2 + 4 * (TC * h + tj), y2[4 * h], y2[4 * h + 1], y2[4 * h + 2], y2[4 * h + 3]);
    }
    if (MODE == SWEEP_GP) {                                        // u += X u_K ; scalars: u_K^T W v_K
        const T u1 = pw[4], v1 = pw[5], u2 = pw[6], v2 = pw[7];
        const T wu1 = fma(w11, u1, w12 * u2), wu2 = fma(w12, u1, w22 * u2);   // W u_K
        acc_m = fma(wu1, v1, fma(wu2, v2, acc_m));
        acc_q = fma(wu1, u1, fma(wu2, u2, acc_q));
        #pragma unroll
        for (int j = 0; j < SG::NC; ++j) {
            const int cb = tj + TC * j;                            // vector cb / H, row pair cb % H of this thread's rows
            const int ip = cb % SG::H;
            const int off = 4 * (TR * (ip / 2) + ti) + 2 * (ip % 2);
            const PR za = PR::make(z1[off], z1[off + 1]), zb = PR::make(z2[off], z2[off + 1]);
            PR xa = za; xa.scale(w11); xa.fma_bcast(zb, w12);
            PR xb = za; xb.scale(w12); xb.fma_bcast(zb, w22);
            const bool first = cb / SG::H == 0;
            rhs[j].fma_bcast(xa, first ? u1 : v1);
            rhs[j].fma_bcast(xb, first ? u2 : v2);
        }
    }
    if (HAS_NEXT) {
        #pragma unroll
        for (int i = 0; i < SG::H; ++i) {                          // the two column slots of the next pair
            if (!SG::template skip2<MODE>(i, sc1n, S, WP)) { ap[i][sc1n].fma_bcast(x1[i], y1[sc1n]); ap[i][sc1n].fma_bcast(x2[i], y2[sc1n]); }
            if (!SG::template skip2<MODE>(i, sc2n, S, WP)) { ap[i][sc2n].fma_bcast(x1[i], y1[sc2n]); ap[i][sc2n].fma_bcast(x2[i], y2[sc2n]); }
        }
        #pragma unroll
        for (int c = 0; c < SG::SC; ++c)                           // the row pair of the next pair
            if (c != sc1n && c != sc2n && !SG::template skip2<MODE>(ipn, c, S, WP)) {
                ap[ipn][c].fma_bcast(x1[ipn], y1[c]); ap[ipn][c].fma_bcast(x2[ipn], y2[c]);
            }
        sweep2_publish<T, N, TR, TC, MODE, GRN, HCN, WPN>(ap, rhs, zn, ti, tj, qn);
        tile_sync<SG::LANES>();
    }
    #pragma unroll
    for (int i = 0; i < SG::H; ++i)
        #pragma unroll
        for (int c = 0; c < SG::SC; ++c) {
            if (SG::template skip2<MODE>(i, c, S, WP)) continue;
            if (HAS_NEXT && (c == sc1n || c == sc2n || i == ipn)) continue;   // done before the publish
            ap[i][c].fma_bcast(x1[i], y1[c]);
            ap[i][c].fma_bcast(x2[i], y2[c]);
        }
}

// all pivot pairs of level (S, WP), range SUB
template <typename T, int N, int TR, int TC, int MODE, bool UNROLL, int S, int WP, int SUB>
__device__ __forceinline__ void sweep2_range(Pair2<T> (&ap)[N / TR / 2][N / TC], Pair2<T> (&rhs)[SweepGeo<N, TR, TC>::NC],
                                             T *sm, int ti, int tj, T &dmin, T &d, T &acc_m, T &acc_q) {
    using SG = SweepGeo<N, TR, TC>;
    constexpr int PM = SG::PMIN;
    constexpr int Q0 = S * SG::PMAX + SUB * PM;                    // first block of this range
    constexpr int GR = Q0 / TR, HC = Q0 / TC;
    constexpr int J0 = ((2 * S + WP) * SG::SUBS + SUB) * PM;       // sequence number of the first pair of this body
    if (UNROLL) {
        #pragma unroll
        for (int t = 0; t < PM - 1; ++t) {
            const int par = (J0 + t) & 1;
            sweep2_step<T, N, TR, TC, MODE, S, WP, GR, HC, WP, true>(ap, rhs, sm + par * SG::LINE2, sm + (par ^ 1) * SG::LINE2, ti, tj,
                                                                     Q0 + t + 1, dmin, d, acc_m, acc_q);
        }
    } else {
        #pragma unroll 1
        for (int t = 0; t < PM - 1; ++t) {
            const int par = (J0 + t) & 1;
            sweep2_step<T, N, TR, TC, MODE, S, WP, GR, HC, WP, true>(ap, rhs, sm + par * SG::LINE2, sm + (par ^ 1) * SG::LINE2, ti, tj,
                                                                     Q0 + t + 1, dmin, d, acc_m, acc_q);
        }
    }
    constexpr int par = (J0 + PM - 1) & 1;
    constexpr bool last = (S == SG::NS - 1) && (WP == 1) && (SUB == SG::SUBS - 1);
    constexpr int SUBN = (SUB + 1 < SG::SUBS) ? SUB + 1 : 0;
    constexpr int WPN = (SUBN != 0) ? WP : (WP == 1 ? 0 : 1);
    constexpr int SN = (SUBN != 0 || WP != 1) ? S : S + 1;
    constexpr int QN = last ? 0 : SN * SG::PMAX + SUBN * PM;
    sweep2_step<T, N, TR, TC, MODE, S, WP, QN / TR, QN / TC, WPN, !last>(ap, rhs, sm + par * SG::LINE2, sm + (par ^ 1) * SG::LINE2, ti, tj,
                                                                        QN, dmin, d, acc_m, acc_q);
}

template <typename T, int N, int TR, int TC, int MODE, bool UNROLL, int S, int WP, int SUB>
struct Sweep2Ranges {
    static __device__ __forceinline__ void run(Pair2<T> (&ap)[N / TR / 2][N / TC], Pair2<T> (&rhs)[SweepGeo<N, TR, TC>::NC],
                                               T *sm, int ti, int tj, T &dmin, T &d, T &acc_m, T &acc_q) {
        using SG = SweepGeo<N, TR, TC>;
        sweep2_range<T, N, TR, TC, MODE, UNROLL, S, WP, SUB>(ap, rhs, sm, ti, tj, dmin, d, acc_m, acc_q);
        constexpr int SUBN = (SUB + 1 < SG::SUBS) ? SUB + 1 : 0;
        constexpr int WPN = (SUBN != 0) ? WP : (WP == 1 ? 0 : 1);
        constexpr int SN = (SUBN != 0 || WP != 1) ? S : S + 1;
        if constexpr (SN < SG::NS) Sweep2Ranges<T, N, TR, TC, MODE, UNROLL, SN, WPN, SUBN>::run(ap, rhs, sm, ti, tj, dmin, d, acc_m, acc_q);
    }
};

// the whole elimination of one matrix whose tile is loaded: prologue publish + all steps
// BLK == 0: the same pivots WITHOUT look-ahead: publish -> barrier -> update, every pivot of a range through one
// rolled body (half the code of the look-ahead form, whose last pivot of a range is peeled because the slots
// of ITS successor differ).  An experiment on instruction-cache pressure (variant 8).
template <typename T, int N, int TR, int TC, int MODE, int S, int W, int SUB>
struct SweepRangesPlain {
    static __device__ __forceinline__ void run(Pair2<T> (&ap)[N / TR / 2][N / TC], Pair2<T> (&rhs)[SweepGeo<N, TR, TC>::NC],
                                               T *sm, int ti, int tj, T &dmin, T &d, T &acc_m, T &acc_q) {
        using SG = SweepGeo<N, TR, TC>;
        constexpr int PM = SG::PMIN;
        constexpr int Q0 = S * SG::PMAX + SUB * PM;
        constexpr int GR = Q0 / TR, HC = Q0 / TC;
        constexpr int J0 = ((4 * S + W) * SG::SUBS + SUB) * PM;
        #pragma unroll 1
        for (int t = 0; t < PM; ++t) {
            T *z = sm + ((J0 + t) & 1) * SG::LINE;
            sweep_publish<T, N, TR, TC, MODE, GR, HC, W>(ap, rhs, z, ti, tj, Q0 + t);
            tile_sync<SG::LANES>();
            sweep_step<T, N, TR, TC, MODE, S, W, 0, 0, 0, false>(ap, rhs, z, z, ti, tj, 0, dmin, d, acc_m, acc_q);
        }
        constexpr int SUBN = (SUB + 1 < SG::SUBS) ? SUB + 1 : 0;
        constexpr int WN = (SUBN != 0) ? W : (W == 3 ? 0 : W + 1);
        constexpr int SN = (SUBN != 0 || W != 3) ? S : S + 1;
        if constexpr (SN < SG::NS) SweepRangesPlain<T, N, TR, TC, MODE, SN, WN, SUBN>::run(ap, rhs, sm, ti, tj, dmin, d, acc_m, acc_q);
    }
};

template <typename T, int N, int TR, int TC, int MODE, bool UNROLL, int BLK>
__device__ __forceinline__ void sweep_eliminate(Pair2<T> (&ap)[N / TR / 2][N / TC], Pair2<T> (&rhs)[SweepGeo<N, TR, TC>::NC],
                                                T *sm, int ti, int tj, T &dmin, T &d, T &acc_m, T &acc_q) {
    using SG = SweepGeo<N, TR, TC>;
    if constexpr (BLK == 0) {
        SweepRangesPlain<T, N, TR, TC, MODE, 0, 0, 0>::run(ap, rhs, sm, ti, tj, dmin, d, acc_m, acc_q);
    } else if constexpr (BLK == 2) {
        sweep2_publish<T, N, TR, TC, MODE, 0, 0, 0>(ap, rhs, sm, ti, tj, 0);
        tile_sync<SG::LANES>();
        Sweep2Ranges<T, N, TR, TC, MODE, UNROLL, 0, 0, 0>::run(ap, rhs, sm, ti, tj, dmin, d, acc_m, acc_q);
    } else {
        sweep_publish<T, N, TR, TC, MODE, 0, 0, 0>(ap, rhs, sm, ti, tj, 0);
        tile_sync<SG::LANES>();
        SweepRanges<T, N, TR, TC, MODE, UNROLL, 0, 0, 0>::run(ap, rhs, sm, ti, tj, dmin, d, acc_m, acc_q);
    }
}

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

template <typename T, int N, int TR, int TC, bool UNROLL, typename IO, int MINB, int BLK = 1>
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
    T *sm = smem + grp * (BLK == 2 ? SG::WORDS2 : SG::WORDS);

    #pragma unroll 1
    for (i64 base = (i64)blockIdx.x * SG::MPB; base < batch; base += (i64)gridDim.x * SG::MPB) {
        const i64 m = base + grp;
        const bool valid = m < batch;
        const T *__restrict__ src;
        T *__restrict__ dst;
        int nn = N;                                                // order of the matrix inside the N x N tile
        i64 iidx = m;                                              // where its info goes
        if constexpr (IoTraits<IO>::PADDED) {
            const T *s0; T *d0;
            io.get(valid ? m : batch - 1, s0, d0, nn, iidx);
            src = s0; dst = d0;
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
        sweep_eliminate<T, N, TR, TC, SWEEP_INVERSE, UNROLL, BLK>(ap, rhs, sm, ti, tj, dmin, d, acc_m, acc_q);

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
template <typename T, int N, int TR, int TC, bool UNROLL, int MINB, int BLK = 1>
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
    T *sm = smem + grp * (BLK == 2 ? SG::WORDS2 : SG::WORDS);

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
        sweep_eliminate<T, N, TR, TC, SWEEP_GP, UNROLL, BLK>(ap, rhs, sm, ti, tj, dmin, d, acc_m, acc_q);

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

// ==========================================================================================
// TMA tile I/O for the warp tiers whose matrix columns are exactly one 128-byte line (n = 32 fp32,
// n = 16 fp64): the global side of the kernel leaves the LSU pipe.
//
// ncu on the direct-access kernel (profiles/r1_sweep32_summary.md): the LSU data pipe is the busiest unit
// (67 %), and 40 % of its wavefronts are the 16-byte global loads / stores of the four matrices that share
// a warp (16 distinct lines per instruction).  Here one lane issues ONE `cp.async.bulk.tensor.2d` per
// warp tile: the batch is described to the TMA unit as a 2-D tensor [N rows x (N batch) columns], the box
// is the warp's MPW consecutive matrices, and the 128-byte swizzle mode (16-byte chunk index XOR column
// mod 8) makes the tile reads (transposed twins of the upper triangle) and the mirrored stores of the
// 2 x 4 lane grid bank-conflict free; the natural-position stores are 2-way.  Results go back with one
// `cp.async.bulk.tensor` store per warp tile (columns beyond the batch are clipped by the tensor map,
// loads of them are zero-filled), so the tail needs no special code.  Sequence per warp and tile:
//   wait(mbarrier) -> registers <- smem -> eliminate -> smem <- registers -> fence.proxy.async ->
//   bulk store -> wait_group.read -> arm mbarrier + bulk load of the next tile.
// All mbarrier waits are bounded and trap instead of hanging.
// ==========================================================================================
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 22)) __trap();                        // a lost transaction must not hang the GPU
}
__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, int c0, int c1, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void *tmap, int c0, int c1, const void *src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

struct TmaMaps {                                                   // 128-byte CUtensorMap objects, opaque here
    alignas(64) unsigned char in[128];                             // box = the MPW matrices of a warp tile
    alignas(64) unsigned char out[128];
    alignas(64) unsigned char in1[128];                            // box = one matrix (interleaved-lane variant)
    alignas(64) unsigned char out1[128];
};

template <typename T, int N, int TR, int TC>
struct SweepTmaGeo {
    using SG = SweepGeo<N, TR, TC>;
    static_assert(N * sizeof(T) == 128, "one matrix column must be one 128-byte swizzle line");
    static_assert(SG::LANES < 32, "warp tiers with several matrices per warp");
    static constexpr int MPW = 32 / SG::LANES;                     // matrices per warp = per TMA box
    static constexpr int WARPS = SG::BLOCK / 32;
    static constexpr int MAT_BYTES = N * N * (int)sizeof(T);
    static constexpr int BOX_BYTES = MPW * MAT_BYTES;
    // INTERLEAVE: every matrix of the warp gets its own box, 128 bytes further than a plain 4 KB stride, so that the
    // swizzle phase (address bits 7-9) differs from matrix to matrix
#ifndef INVGPU_TMA_STAGGER
#define INVGPU_TMA_STAGGER 1       // swizzle lines between the boxes of neighbouring matrices (interleaved layout)
#endif
    template <bool INTERLEAVE> __host__ __device__ static constexpr int mstride() { return MAT_BYTES + (INTERLEAVE ? 128 * INVGPU_TMA_STAGGER : 0); }
    template <bool INTERLEAVE> __host__ __device__ static constexpr size_t smem() {
        return (size_t)WARPS * (INTERLEAVE ? (MPW * mstride<INTERLEAVE>() + 1023) / 1024 * 1024 : MPW * mstride<INTERLEAVE>()) +
               (size_t)SG::MPB * SG::WORDS * sizeof(T) + WARPS * 8;
    }
    static constexpr size_t SMEM = (size_t)WARPS * BOX_BYTES + (size_t)SG::MPB * SG::WORDS * sizeof(T) + WARPS * 8;
    // byte offset of the 16-byte chunk `chunk` of column `col` inside the swizzled buffer of matrix j of the warp
    // (the warp's first buffer is 1024-byte aligned; j shifts the swizzle phase only in the interleaved layout)
    template <bool INTERLEAVE>
    static __device__ __forceinline__ int off(int chunk, int col, int j) {
        return col * 128 + ((chunk ^ ((col + (INTERLEAVE ? INVGPU_TMA_STAGGER * j : 0)) & 7)) << 4);
    }
};

// DIRECT_OUT: results leave through ordinary 16-byte global stores instead of the bulk store; the box is then free as
// soon as the tile sits in registers, so the bulk load of the NEXT warp tile is issued before the elimination and
// lands behind it (prefetch for free) -- at the price of the ~140 store wavefronts per matrix on the LSU pipe.
// INTERLEAVE: lane = ti + TR (mat + MPW tj): the owners of a pivot column of all the matrices of the warp sit in one
// quarter-warp, so publishing a column costs one shared-memory wavefront per 16-byte store instead of four.  With
// direct global access that lane map loses the 32-byte runs of the loads / stores (measured -15 %); here the TMA
// unit does the global side, and per-matrix boxes 128 bytes apart keep the tile accesses conflict-free.
template <typename T, int N, int TR, int TC, bool UNROLL, int MINB, bool DIRECT_OUT = false, bool INTERLEAVE = false>
__global__ void __launch_bounds__((SweepGeo<N, TR, TC>::BLOCK), MINB)
sweep_spd_tma_kernel(const __grid_constant__ TmaMaps maps, const T *__restrict__ in, i64 in_stride, T *__restrict__ out, i64 out_stride,
                     i64 batch, int *__restrict__ info) {
    using SG = SweepGeo<N, TR, TC>;
    using TG = SweepTmaGeo<T, N, TR, TC>;
    using PR = Pair2<T>;
    constexpr int SR = SG::SR, SC = SG::SC, H = SG::H;
    constexpr int CH = 16 / (int)sizeof(T);                        // elements per 16-byte chunk (= 4 for fp32)
    static_assert(CH == 4, "the tile code assumes 4-element chunks");
    extern __shared__ __align__(1024) unsigned char smem_raw_tma[];   // 128-byte swizzle repeats every 1024 bytes
    unsigned char *base = smem_raw_tma;
    const int warp = threadIdx.x >> 5, wl = threadIdx.x & 31;
    int mat, ti, tj;
    if (INTERLEAVE) { ti = wl % TR; mat = (wl / TR) % TG::MPW; tj = wl / (TR * TG::MPW); }
    else { mat = wl / SG::LANES; ti = (wl % SG::LANES) / TC; tj = wl % TC; }
    const bool lead = (ti == 0 && tj == 0);
    constexpr int MS = TG::template mstride<INTERLEAVE>();
    constexpr int WBOX = (TG::MPW * MS + (INTERLEAVE ? 1023 : 0)) / (INTERLEAVE ? 1024 : 1) * (INTERLEAVE ? 1024 : 1);   // per-warp bytes, 1 KB multiple
    unsigned char *box = base + (size_t)warp * WBOX;               // this warp's MPW matrices
    unsigned char *buf = box + (size_t)mat * MS;                   // this thread's matrix
    T *lines = reinterpret_cast<T *>(base + (size_t)TG::WARPS * WBOX);
    T *sm = lines + (size_t)(warp * TG::MPW + mat) * SG::WORDS;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(lines + (size_t)SG::MPB * SG::WORDS) + warp;

    if ((smem_u32(base) & 1023u) != 0) __trap();                   // the swizzle pattern is defined on 1024-byte aligned addresses
    if (wl == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const i64 ntiles = (batch + TG::MPW - 1) / TG::MPW;
    const i64 tstride = (i64)gridDim.x * TG::WARPS;
    i64 tile = (i64)blockIdx.x * TG::WARPS + warp;
    unsigned phase = 0;
    auto issue_load = [&](i64 t) {                                 // lane 0 only
        mbar_expect_tx(bar, TG::BOX_BYTES);
        if (INTERLEAVE) {
            #pragma unroll
            for (int j = 0; j < TG::MPW; ++j) tma_load_2d(box + j * MS, &maps.in1, 0, (int)((t * TG::MPW + j) * N), bar);
        } else {
            tma_load_2d(box, &maps.in, 0, (int)(t * TG::MPW * N), bar);
        }
    };
    if (tile < ntiles && wl == 0) issue_load(tile);
    #pragma unroll 1
    for (; tile < ntiles; tile += tstride) {
        const i64 m = tile * TG::MPW + mat;
        const bool valid = m < batch;
        mbar_wait(bar, phase);
        phase ^= 1;

        PR ap[H][SC];
        {
            T a[SR][SC];
            #pragma unroll
            for (int g = 0; g < SG::NGR; ++g)
                #pragma unroll
                for (int h = 0; h < SG::NGC; ++h) {
                    const int br = TR * g + ti, bc = TC * h + tj;
                    #pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        // lower block (br >= bc): its transposed twin from the UPPER triangle: column 4 br + p, rows 4 bc ..
                        if (!SG::upper(2 * g, 4 * h) && br >= bc)
                            ld4(reinterpret_cast<const T *>(buf + TG::template off<INTERLEAVE>(bc, 4 * br + p, mat)), a[4 * g + p][4 * h], a[4 * g + p][4 * h + 1],
                                a[4 * g + p][4 * h + 2], a[4 * g + p][4 * h + 3]);
                        else { a[4 * g + p][4 * h] = T(0); a[4 * g + p][4 * h + 1] = T(0); a[4 * g + p][4 * h + 2] = T(0); a[4 * g + p][4 * h + 3] = T(0); }
                    }
                }
            #pragma unroll
            for (int i = 0; i < H; ++i)
                #pragma unroll
                for (int c = 0; c < SC; ++c) ap[i][c] = PR::make(-a[2 * i][c], -a[2 * i + 1][c]);
        }
        __syncwarp();                                              // everybody has its tile: the box may be overwritten
        if (DIRECT_OUT && wl == 0 && tile + tstride < ntiles) issue_load(tile + tstride);   // prefetch under the elimination

        T dmin = T(1), d = T(1), acc_m = T(0), acc_q = T(0);
        PR rhs[SG::NC];
        sweep_eliminate<T, N, TR, TC, SWEEP_INVERSE, UNROLL, 1>(ap, rhs, sm, ti, tj, dmin, d, acc_m, acc_q);

        const bool bad = valid && (!(dmin > T(0)) || !(d == d));
        int st = 0;
        if (__any_sync(0xffffffffu, bad)) {                        // rare: LAPACK's natural-order index; global output as scratch
            if (bad && lead) {
                st = exact_potrf_info<T>(in + m * in_stride, out + m * out_stride, N);
                if (st == 0) st = N;
            }
            fence_proxy_async();                                   // scratch writes are ordered before the bulk store below
            __syncwarp();
        }
        if (valid && lead && info) info[m] = st;
        if constexpr (DIRECT_OUT) {
            if (valid) {
                T *__restrict__ dst = out + m * out_stride;
                #pragma unroll
                for (int g = 0; g < SG::NGR; ++g) {
                    #pragma unroll
                    for (int h = 0; h < SG::NGC; ++h) {
                        const int br = TR * g + ti, bc = TC * h + tj;
                        if (bad) {
                            #pragma unroll
                            for (int v = 0; v < 4; ++v)
                                stg4(dst + (size_t)(4 * bc + v) * N + 4 * br, dev_nan<T>(), dev_nan<T>(), dev_nan<T>(), dev_nan<T>());
                            continue;
                        }
                        if (SG::upper(2 * g, 4 * h)) continue;
                        T b[4][4];
                        #pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            b[0][v] = ap[2 * g][4 * h + v].lo(); b[1][v] = ap[2 * g][4 * h + v].hi();
                            b[2][v] = ap[2 * g + 1][4 * h + v].lo(); b[3][v] = ap[2 * g + 1][4 * h + v].hi();
                        }
                        if (br == bc) {
                            #pragma unroll
                            for (int v = 0; v < 4; ++v)
                                stg4(dst + (size_t)(4 * bc + v) * N + 4 * br, v <= 0 ? b[0][v] : b[v][0], v <= 1 ? b[1][v] : b[v][1],
                                     v <= 2 ? b[2][v] : b[v][2], b[3][v]);
                        } else if (br > bc) {
                            #pragma unroll
                            for (int v = 0; v < 4; ++v)
                                stg4(dst + (size_t)(4 * bc + v) * N + 4 * br, b[0][v], b[1][v], b[2][v], b[3][v]);
                            #pragma unroll
                            for (int ww = 0; ww < 4; ++ww)
                                stg4(dst + (size_t)(4 * br + ww) * N + 4 * bc, b[ww][0], b[ww][1], b[ww][2], b[ww][3]);
                        }
                    }
                }
            }
            continue;
        }
        // ---- result -> swizzled box (both triangles; NaN for a flagged matrix)
        #pragma unroll
        for (int g = 0; g < SG::NGR; ++g) {
            #pragma unroll
            for (int h = 0; h < SG::NGC; ++h) {
                const int br = TR * g + ti, bc = TC * h + tj;
                if (bad) {
                    #pragma unroll
                    for (int v = 0; v < 4; ++v)
                        st4(reinterpret_cast<T *>(buf + TG::template off<INTERLEAVE>(br, 4 * bc + v, mat)), dev_nan<T>(), dev_nan<T>(), dev_nan<T>(), dev_nan<T>());
                    continue;
                }
                if (SG::upper(2 * g, 4 * h)) continue;
                T b[4][4];
                #pragma unroll
                for (int v = 0; v < 4; ++v) {
                    b[0][v] = ap[2 * g][4 * h + v].lo(); b[1][v] = ap[2 * g][4 * h + v].hi();
                    b[2][v] = ap[2 * g + 1][4 * h + v].lo(); b[3][v] = ap[2 * g + 1][4 * h + v].hi();
                }
                if (br == bc) {
                    #pragma unroll
                    for (int v = 0; v < 4; ++v)
                        st4(reinterpret_cast<T *>(buf + TG::template off<INTERLEAVE>(br, 4 * bc + v, mat)), v <= 0 ? b[0][v] : b[v][0], v <= 1 ? b[1][v] : b[v][1],
                            v <= 2 ? b[2][v] : b[v][2], b[3][v]);
                } else if (br > bc) {
                    #pragma unroll
                    for (int v = 0; v < 4; ++v)
                        st4(reinterpret_cast<T *>(buf + TG::template off<INTERLEAVE>(br, 4 * bc + v, mat)), b[0][v], b[1][v], b[2][v], b[3][v]);
                    #pragma unroll
                    for (int ww = 0; ww < 4; ++ww)
                        st4(reinterpret_cast<T *>(buf + TG::template off<INTERLEAVE>(bc, 4 * br + ww, mat)), b[ww][0], b[ww][1], b[ww][2], b[ww][3]);
                }
            }
        }
        fence_proxy_async();                                       // generic-proxy writes -> visible to the TMA unit
        __syncwarp();
        if (wl == 0) {
            if (INTERLEAVE) {
                #pragma unroll
                for (int j = 0; j < TG::MPW; ++j) tma_store_2d(&maps.out1, 0, (int)((tile * TG::MPW + j) * N), box + j * MS);
            } else {
                tma_store_2d(&maps.out, 0, (int)(tile * TG::MPW * N), box);
            }
            tma_store_commit_and_wait_read();                      // the box has been read: it can take the next tile
            if (tile + tstride < ntiles) issue_load(tile + tstride);
        }
        __syncwarp();
    }
    if (wl == 0) tma_store_wait_all();                             // global writes complete before the CTA retires
}

// ==========================================================================================
// n = 8: ONE THREAD PER MATRIX, the whole sweep in registers, TMA tile I/O.
//
// An 8x8 matrix is 36 lower-triangle values: a thread sweeps it with static register indices and no
// shared-memory traffic at all, so the kernel is pure data movement -- which is the hard part when every
// lane owns a 256-byte matrix: direct 16-byte accesses touch 32 different lines per instruction.  Here a
// warp tile (32 consecutive matrices, 8 KB fp32 / 16 KB fp64) is one `cp.async.bulk.tensor.2d` into a
// 128-byte-swizzled box (2-way conflicts for the per-lane reads instead of 8-way), results go back through
// the same box with one bulk store, and two boxes per warp alternate so that the load of tile i+1 and the
// store of tile i-1 overlap the arithmetic of tile i.  Pivots in natural order: `info` is spotrf's directly.
// ==========================================================================================
template <typename T, int NBUF>
struct Small8Geo {
    static constexpr int N = 8;
    static constexpr int WARPS = 4, BLOCK = 32 * WARPS, MPB = 32 * WARPS;
    static constexpr int MAT_BYTES = 64 * (int)sizeof(T);
    static constexpr int BOX_BYTES = 32 * MAT_BYTES;               // 32 matrices
    static constexpr int LINES_PER_MAT = MAT_BYTES / 128;          // 2 (fp32) or 4 (fp64)
    static constexpr int EPC = 16 / (int)sizeof(T);                // elements per 16-byte chunk
    static constexpr size_t SMEM = (size_t)WARPS * NBUF * BOX_BYTES + WARPS * 2 * 8;
    // byte offset inside the (1024-byte aligned) box of the chunk holding rows r .. r+EPC-1 of column c of lane l's matrix
    static __device__ __forceinline__ int off(int l, int c, int r) {
        const int byte = l * MAT_BYTES + (c * 8 + r) * (int)sizeof(T);
        const int line = byte >> 7, chunk = (byte >> 4) & 7;
        return (line << 7) + ((chunk ^ (line & 7)) << 4);
    }
};

// per-thread arithmetic on the matrix of lane l inside the swizzled box: read it, transform it in registers, write
// the result back into the box; returns LAPACK's info
template <typename T>
struct Spd8Math {                                                  // SPD inverse: symmetric sweep on the lower triangle
    template <typename G>
    static __device__ __forceinline__ int apply(unsigned char *box, int l) {
        constexpr int N = 8, EPC = G::EPC;
        // ---- upper triangle -> registers (t[i][c], i >= c, = -A(c, i))
        T t[N][N];
        #pragma unroll
        for (int c = 0; c < N; ++c)
            #pragma unroll
            for (int r = 0; r <= c; r += EPC) {
                const T *p = reinterpret_cast<const T *>(box + G::off(l, c, r));
                #pragma unroll
                for (int e = 0; e < EPC; ++e)
                    if (r + e <= c) t[c][r + e] = -p[e];
            }
        // ---- the sweep: T = -A -> A^-1, natural pivot order, everything in registers
        int st = 0;
        #pragma unroll
        for (int k = 0; k < N; ++k) {
            const T d = -t[k][k];
            if (st == 0 && !(d > T(0))) st = k + 1;
            const T r = dev_rcp_fast<T>(d);
            T z[N], x[N];
            #pragma unroll
            for (int i = 0; i < N; ++i) { z[i] = (i > k) ? t[i][k] : (i < k ? t[k][i] : T(-1)); x[i] = r * z[i]; }
            #pragma unroll
            for (int i = 0; i < N; ++i)
                #pragma unroll
                for (int c = 0; c <= i; ++c) {
                    if (i == k || c == k) t[i][c] = x[i] * z[c];   // restarted slots: column / row k and the pivot
                    else t[i][c] = fma(x[i], z[c], t[i][c]);
                }
        }
        // ---- both triangles -> box (NaN for a flagged matrix)
        #pragma unroll
        for (int c = 0; c < N; ++c)
            #pragma unroll
            for (int r = 0; r < N; r += EPC) {
                T *p = reinterpret_cast<T *>(box + G::off(l, c, r));
                #pragma unroll
                for (int e = 0; e < EPC; ++e) {
                    const int rr = r + e;
                    p[e] = st ? dev_nan<T>() : (rr >= c ? t[rr][c] : t[c][rr]);
                }
            }
        return st;
    }
};

// General inverse of an 8x8 matrix: in-place Gauss-Jordan with partial pivoting, rows swapped explicitly by
// predicated selects (the pivot row index is the only run-time index), columns un-permuted at the end
// (inv(P A) = A^-1 P^T).  Same arithmetic per element as the oracle / the lane = row kernel (scale the pivot
// row, then a -= f * row); replaces src/gauss/batched_invert.cu:17-95 for dense batches of order exactly 8.
template <typename T>
struct Gj8Math {
    template <typename G>
    static __device__ __forceinline__ int apply(unsigned char *box, int l) {
        constexpr int N = 8, EPC = G::EPC;
        T a[N][N];                                                 // a[r][c]
        #pragma unroll
        for (int c = 0; c < N; ++c)
            #pragma unroll
            for (int r = 0; r < N; r += EPC) {
                const T *p = reinterpret_cast<const T *>(box + G::off(l, c, r));
                #pragma unroll
                for (int e = 0; e < EPC; ++e) a[r + e][c] = p[e];
            }
        int st = 0;
        int perm[N];
        #pragma unroll
        for (int k = 0; k < N; ++k) {
            T best = dev_abs(a[k][k]);
            int p = k;
            if (!(best >= T(0))) best = T(-1);                     // NaN never wins
            #pragma unroll
            for (int r = k + 1; r < N; ++r) {
                const T v = dev_abs(a[r][k]);
                if (v > best) { best = v; p = r; }
            }
            if (st == 0 && !(best > T(0))) st = k + 1;
            perm[k] = p;
            #pragma unroll
            for (int r = k + 1; r < N; ++r) {                      // swap rows k and p
                const bool sw = (p == r);
                #pragma unroll
                for (int c = 0; c < N; ++c) {
                    const T u = a[k][c], v = a[r][c];
                    a[k][c] = sw ? v : u;
                    a[r][c] = sw ? u : v;
                }
            }
            const T pv = T(1) / a[k][k];
            #pragma unroll
            for (int c = 0; c < N; ++c) a[k][c] = (c == k) ? pv : a[k][c] * pv;
            #pragma unroll
            for (int r = 0; r < N; ++r) {
                if (r == k) continue;
                const T f = a[r][k];
                #pragma unroll
                for (int c = 0; c < N; ++c) a[r][c] = (c == k) ? -f * pv : fma(-f, a[k][c], a[r][c]);
            }
        }
        #pragma unroll
        for (int k = N - 2; k >= 0; --k) {                         // undo the row swaps on the columns, last first
            #pragma unroll
            for (int c = k + 1; c < N; ++c) {
                const bool sw = (perm[k] == c);
                #pragma unroll
                for (int r = 0; r < N; ++r) {
                    const T u = a[r][k], v = a[r][c];
                    a[r][k] = sw ? v : u;
                    a[r][c] = sw ? u : v;
                }
            }
        }
        #pragma unroll
        for (int c = 0; c < N; ++c)
            #pragma unroll
            for (int r = 0; r < N; r += EPC) {
                T *p = reinterpret_cast<T *>(box + G::off(l, c, r));
                #pragma unroll
                for (int e = 0; e < EPC; ++e) p[e] = st ? dev_nan<T>() : a[r + e][c];
            }
        return st;
    }
};

template <typename T, int NBUF, int MINB, typename MATH = Spd8Math<T>>
__global__ void __launch_bounds__((Small8Geo<T, NBUF>::BLOCK), MINB)
spd8_tma_kernel(const __grid_constant__ TmaMaps maps, const T *__restrict__ in, T *__restrict__ out, i64 batch, int *__restrict__ info) {
    using G = Small8Geo<T, NBUF>;
    constexpr int N = 8, EPC = G::EPC;
    extern __shared__ __align__(1024) unsigned char smem_raw_s8[];
    const int warp = threadIdx.x >> 5, l = threadIdx.x & 31;
    unsigned char *boxes = smem_raw_s8 + (size_t)warp * NBUF * G::BOX_BYTES;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem_raw_s8 + (size_t)G::WARPS * NBUF * G::BOX_BYTES) + 2 * warp;
    if ((smem_u32(smem_raw_s8) & 1023u) != 0) __trap();
    if (l == 0) {
        mbar_init(bars, 1); mbar_init(bars + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const i64 ntiles = (batch + 31) / 32;
    const i64 tstride = (i64)gridDim.x * G::WARPS;
    i64 tile = (i64)blockIdx.x * G::WARPS + warp;
    constexpr int LPT = 32 * G::LINES_PER_MAT;                     // 128-byte lines per warp tile
    if (tile < ntiles && l == 0) {
        mbar_expect_tx(bars, G::BOX_BYTES);
        tma_load_2d(boxes, &maps.in, 0, (int)(tile * LPT), bars);
    }
    unsigned phase0 = 0, phase1 = 0;
    int cur = 0;
    #pragma unroll 1
    for (; tile < ntiles; tile += tstride, cur ^= (NBUF - 1)) {
        unsigned char *box = boxes + cur * G::BOX_BYTES;
        // two boxes: the other one still feeds the bulk store of the previous tile: wait for that read, then prefetch into it
        if (NBUF == 2 && l == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (tile + tstride < ntiles) {
                mbar_expect_tx(bars + (cur ^ 1), G::BOX_BYTES);
                tma_load_2d(boxes + (cur ^ 1) * G::BOX_BYTES, &maps.in, 0, (int)((tile + tstride) * LPT), bars + (cur ^ 1));
            }
        }
        if (cur == 0) { mbar_wait(bars, phase0); phase0 ^= 1; } else { mbar_wait(bars + 1, phase1); phase1 ^= 1; }
        const i64 m = tile * 32 + l;
        const bool valid = m < batch;

        const int st = MATH::template apply<G>(box, l);
        if (valid && info) info[m] = st;
        fence_proxy_async();
        __syncwarp();
        if (l == 0) {
            tma_store_2d(&maps.out, 0, (int)(tile * LPT), box);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (NBUF == 1) {                                       // one box: reload it as soon as the store has read it
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                if (tile + tstride < ntiles) {
                    mbar_expect_tx(bars, G::BOX_BYTES);
                    tma_load_2d(box, &maps.in, 0, (int)((tile + tstride) * LPT), bars);
                }
            }
        }
        if (NBUF == 1) __syncwarp();
    }
    if (l == 0) tma_store_wait_all();
}

// ==========================================================================================
// n = 16 fp32: one thread per matrix as above (136 lower-triangle values in registers), but the matrices
// of a warp are 1 KB apart, which defeats the 128-byte swizzle (all lanes would hit the same banks).
// Every lane therefore moves ITS matrix with a 1-D bulk copy (`cp.async.bulk`, no tensor map) into a slot
// that is 16 bytes longer than the matrix: consecutive lanes start one bank group further and the per-lane
// 16-byte accesses are conflict-free.  One mbarrier per warp collects the 32 copies; every lane stores its
// own result with a bulk copy of its own group.
// ==========================================================================================
template <typename T, int N, int WARPS>
struct ThreadBulkGeo {
    static constexpr int BLOCK = 32 * WARPS, MPB = 32 * WARPS;
    static constexpr int MAT_BYTES = N * N * (int)sizeof(T);
    static constexpr int SLOT = MAT_BYTES + 16;
    static constexpr int EPC = 16 / (int)sizeof(T);
    static constexpr size_t SMEM = (size_t)WARPS * 32 * SLOT + WARPS * 8 + 16;
};

__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void *dst, const void *src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

template <typename T, int N, int WARPS, int MINB>
__global__ void __launch_bounds__((ThreadBulkGeo<T, N, WARPS>::BLOCK), MINB)
spd_thread_bulk_kernel(const T *__restrict__ in, i64 in_stride, T *__restrict__ out, i64 out_stride, i64 batch, int *__restrict__ info) {
    using G = ThreadBulkGeo<T, N, WARPS>;
    constexpr int EPC = G::EPC;
    extern __shared__ __align__(16) unsigned char smem_raw_tb[];
    const int warp = threadIdx.x >> 5, l = threadIdx.x & 31;
    unsigned char *slot = smem_raw_tb + ((size_t)warp * 32 + l) * G::SLOT;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_raw_tb + (((size_t)WARPS * 32 * G::SLOT + 15) & ~(size_t)15)) + warp;
    if (l == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const i64 ntiles = (batch + 31) / 32;
    const i64 tstride = (i64)gridDim.x * WARPS;
    i64 tile = (i64)blockIdx.x * WARPS + warp;
    auto issue = [&](i64 t) {                                      // all lanes: lane 0 arms, every valid lane copies its matrix
        const i64 m = t * 32 + l;
        const i64 left = batch - t * 32;
        if (l == 0) mbar_expect_tx(bar, (unsigned)((left < 32 ? left : 32) * G::MAT_BYTES));
        __syncwarp();
        if (m < batch) bulk_load_1d(slot, in + m * in_stride, G::MAT_BYTES, bar);
    };
    if (tile < ntiles) issue(tile);
    unsigned phase = 0;
    #pragma unroll 1
    for (; tile < ntiles; tile += tstride) {
        const i64 m = tile * 32 + l;
        const bool valid = m < batch;
        mbar_wait(bar, phase);
        phase ^= 1;

        T t[N][N];                                                 // t[i][c], i >= c, = -A(c, i) (upper triangle of the input)
        #pragma unroll
        for (int c = 0; c < N; ++c)
            #pragma unroll
            for (int r = 0; r <= c; r += EPC) {
                const T *p = reinterpret_cast<const T *>(slot + (c * N + r) * (int)sizeof(T));
                #pragma unroll
                for (int e = 0; e < EPC; ++e)
                    if (r + e <= c) t[c][r + e] = -p[e];
            }
        int st = 0;
        #pragma unroll
        for (int k = 0; k < N; ++k) {
            const T d = -t[k][k];
            if (st == 0 && !(d > T(0))) st = k + 1;
            const T r = dev_rcp_fast<T>(d);
            T z[N], x[N];
            #pragma unroll
            for (int i = 0; i < N; ++i) { z[i] = (i > k) ? t[i][k] : (i < k ? t[k][i] : T(-1)); x[i] = r * z[i]; }
            #pragma unroll
            for (int i = 0; i < N; ++i)
                #pragma unroll
                for (int c = 0; c <= i; ++c) {
                    if (i == k || c == k) t[i][c] = x[i] * z[c];
                    else t[i][c] = fma(x[i], z[c], t[i][c]);
                }
        }
        if (valid && info) info[m] = st;
        #pragma unroll
        for (int c = 0; c < N; ++c)
            #pragma unroll
            for (int r = 0; r < N; r += EPC) {
                T *p = reinterpret_cast<T *>(slot + (c * N + r) * (int)sizeof(T));
                #pragma unroll
                for (int e = 0; e < EPC; ++e) {
                    const int rr = r + e;
                    p[e] = st ? dev_nan<T>() : (rr >= c ? t[rr][c] : t[c][rr]);
                }
            }
        fence_proxy_async();
        if (valid) bulk_store_1d(out + m * out_stride, slot, G::MAT_BYTES);
        asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 0;" ::: "memory");   // own slot read: reusable
        __syncwarp();
        if (tile + tstride < ntiles) issue(tile + tstride);
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ==========================================================================================
// Fused GP mean / variance for tiny matrices (n = 8; n = 16 fp32): ONE THREAD PER EVALUATION.
// B moves with per-lane 1-D bulk copies into padded slots (as in spd_thread_bulk_kernel); the slot is free as
// soon as the upper triangle sits in registers, so the copy of the next evaluation is issued before the
// arithmetic and lands behind it.  diag(C) is added on the way into the registers, the factorisation is
// A = L D L^T with the two right-hand sides riding along, the scalars are sum u_k v_k / d_k and
// E - sum u_k^2 / d_k (reference src/gauss_bench.cu:127-265, 275-409 in one launch; nothing but the scalars is
// written).  Natural pivot order: `info` is spotrf's directly.
// ==========================================================================================
template <typename T, int N, int WARPS, int MINB>
__global__ void __launch_bounds__((ThreadBulkGeo<T, N, WARPS>::BLOCK), MINB)
gp_thread_kernel(GpIO<T> io, i64 batch, int *__restrict__ info) {
    using G = ThreadBulkGeo<T, N, WARPS>;
    constexpr int EPC = G::EPC;
    extern __shared__ __align__(16) unsigned char smem_raw_gt[];
    const int warp = threadIdx.x >> 5, l = threadIdx.x & 31;
    unsigned char *slot = smem_raw_gt + ((size_t)warp * 32 + l) * G::SLOT;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_raw_gt + (((size_t)WARPS * 32 * G::SLOT + 15) & ~(size_t)15)) + warp;
    if (l == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const i64 ntiles = (batch + 31) / 32;
    const i64 tstride = (i64)gridDim.x * WARPS;
    i64 tile = (i64)blockIdx.x * WARPS + warp;
    auto issue = [&](i64 t) {
        const i64 m = t * 32 + l;
        const i64 left = batch - t * 32;
        if (l == 0) mbar_expect_tx(bar, (unsigned)((left < 32 ? left : 32) * G::MAT_BYTES));
        __syncwarp();
        if (m < batch) bulk_load_1d(slot, io.b + m * (i64)(N * N), G::MAT_BYTES, bar);
    };
    if (tile < ntiles) issue(tile);
    unsigned phase = 0;
    #pragma unroll 1
    for (; tile < ntiles; tile += tstride) {
        const i64 m = tile * 32 + l;
        const bool valid = m < batch;
        const i64 mm = valid ? m : batch - 1;
        // right-hand sides and the diagonal: N contiguous values per lane, straight from global memory
        T u[N], v[N], cd[N];
        {
            const T *__restrict__ av = io.a + mm * N;
            const T *__restrict__ dv = (io.d ? io.d : io.a) + mm * N;
            const T *__restrict__ cv = io.c + mm * N;
            #pragma unroll
            for (int i = 0; i < N; i += 4) {
                ldg4(av + i, u[i], u[i + 1], u[i + 2], u[i + 3]);
                ldg4(dv + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
                ldg4(cv + i, cd[i], cd[i + 1], cd[i + 2], cd[i + 3]);
            }
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        T t[N][N];                                                 // t[i][c], i >= c, = (B + diag C)(c, i): upper triangle of the input
        #pragma unroll
        for (int c = 0; c < N; ++c)
            #pragma unroll
            for (int r = 0; r <= c; r += EPC) {
                const T *p = reinterpret_cast<const T *>(slot + (c * N + r) * (int)sizeof(T));
                #pragma unroll
                for (int e = 0; e < EPC; ++e)
                    if (r + e <= c) t[c][r + e] = p[e] + ((r + e == c) ? cd[c] : T(0));
            }
        __syncwarp();                                              // every lane has read its slot
        if (tile + tstride < ntiles) issue(tile + tstride);        // prefetch under the arithmetic

        int st = 0;
        T acc_m = T(0), acc_q = T(0);
        #pragma unroll
        for (int k = 0; k < N; ++k) {
            const T d = t[k][k];
            if (st == 0 && !(d > T(0))) st = k + 1;
            const T r = dev_rcp_fast<T>(d);
            const T ur = u[k] * r;
            acc_m = fma(ur, v[k], acc_m);
            acc_q = fma(ur, u[k], acc_q);
            #pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const T li = t[i][k] * r;                          // L(i, k)
                u[i] = fma(-li, u[k], u[i]);
                v[i] = fma(-li, v[k], v[i]);
                #pragma unroll
                for (int c = k + 1; c <= i; ++c) t[i][c] = fma(-li, t[c][k], t[i][c]);
            }
        }
        if (valid) {
            if (io.means) io.means[m] = st ? dev_nan<T>() : acc_m;
            if (io.variances) io.variances[m] = st ? dev_nan<T>() : io.e[m] - acc_q;
            if (info) info[m] = st;
        }
    }
}

}  // namespace invgpu

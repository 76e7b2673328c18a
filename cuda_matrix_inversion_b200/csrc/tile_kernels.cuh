// tile_kernels.cuh -- register-tiled fast tiers for the SPD (Cholesky) inverse / factor.
//
// One matrix of order N is owned by a TR x TC grid of threads; thread (ti, tj) keeps an
// (N/TR) x (N/TC) register tile made of 4x4 sub-blocks.  Memory 4-blocks (rows 4b..4b+3) are dealt
// to the thread rows cyclically (b mod TR), likewise columns (b mod TC), so every sub-block is four
// 128-bit global accesses and every shared-memory broadcast line is read 16 bytes at a time.
//
// Elimination order.  For the INVERSE (and anything else invariant under a symmetric permutation,
// e.g. the GP mean) the pivots are taken in a permuted "logical" order in which consecutive pivots
// belong to consecutive threads (pure cyclic): (P A P^T)^-1 = P A^-1 P^T, so the result lands in
// the right memory positions while the trailing matrix shrinks evenly over the threads and about
// half of the FMA slots of a block-cyclic order disappear at compile time.  The FACTOR-only entry
// points must return the natural-order L and use the identity order (PERM = false).
//
//   N   TR x TC  threads/matrix  matrices/warp        tile (fp32 registers)
//   8   1 x 1         1              32                8 x 8   (64)
//   16  2 x 2         4               8                8 x 8   (64)
//   32  4 x 2         8               4                8 x 16  (128)      warp tiers: __syncwarp only
//   64  8 x 4        32               1                8 x 16  (128)
//   128 16 x 16     256           one per CTA          8 x 8   (64)       CTA tier: __syncthreads
//
// Every step k of the three phases is a rank-1 update whose two vectors are broadcast through a
// tiny double-buffered shared-memory line (one barrier per step); the k loop is fully unrolled so
// every register index is static and slots that cannot be active at step k vanish at compile time.
//
//   potrf : A = L L^T, right-looking.  Step k: the owners of column k scale it by 1/sqrt(a_kk)
//           (pivot fetched with one shuffle) and publish it;  a_ic -= l_i l_c  on the lower part.
//   trtri : M = L^-1 in place, outer-product form.  Step k: row k of M is final; publish it and
//           column k of L;  acc_ic += L_ik M_kc  (i > k, c <= k).
//   lauum : P = M^T M in place on the FULL tile.  Step k: publish row k of M;  P_ic += M_ki M_kc
//           for all i, c <= k, so each thread ends up with its complete part of A^-1 and simply
//           stores it (no mirroring, no exchange).
//
// Only the upper triangle of the column-major input is loaded (spotrf_("U") convention of the
// reference CPU path, src/inverse.c:92): sub-blocks below the diagonal are fetched from their
// transposed twins.  info: first k (1-based, natural order) with a non-positive / NaN pivot, like
// spotrf; with the permuted order a flagged matrix is re-examined in natural order by one thread
// (rare path) so that the reported index is LAPACK's.  Flagged outputs are NaN.
#pragma once

#include "common.cuh"

#ifndef INVGPU_WARP_TIER_BLOCK
#define INVGPU_WARP_TIER_BLOCK 128
#endif

namespace invgpu {

// ------------------------------------------------------------------------------------------
// 4-element vector access (16 B for float, 2 x 16 B for double)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld4(const float *p, float &a, float &b, float &c, float &d) {
    const float4 v = *reinterpret_cast<const float4 *>(p);
    a = v.x; b = v.y; c = v.z; d = v.w;
}
__device__ __forceinline__ void ld4(const double *p, double &a, double &b, double &c, double &d) {
    const double2 u = *reinterpret_cast<const double2 *>(p);
    const double2 v = *reinterpret_cast<const double2 *>(p + 2);
    a = u.x; b = u.y; c = v.x; d = v.y;
}
__device__ __forceinline__ void st4(float *p, float a, float b, float c, float d) {
    *reinterpret_cast<float4 *>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void st4(double *p, double a, double b, double c, double d) {
    *reinterpret_cast<double2 *>(p) = make_double2(a, b);
    *reinterpret_cast<double2 *>(p + 2) = make_double2(c, d);
}
// streaming global access: every byte is touched exactly once
__device__ __forceinline__ void ldg4(const float *p, float &a, float &b, float &c, float &d) {
    const float4 v = __ldcs(reinterpret_cast<const float4 *>(p));
    a = v.x; b = v.y; c = v.z; d = v.w;
}
__device__ __forceinline__ void ldg4(const double *p, double &a, double &b, double &c, double &d) {
    const double2 u = __ldcs(reinterpret_cast<const double2 *>(p));
    const double2 v = __ldcs(reinterpret_cast<const double2 *>(p + 2));
    a = u.x; b = u.y; c = v.x; d = v.y;
}
__device__ __forceinline__ void stg4(float *p, float a, float b, float c, float d) {
    __stcs(reinterpret_cast<float4 *>(p), make_float4(a, b, c, d));
}
__device__ __forceinline__ void stg4(double *p, double a, double b, double c, double d) {
    __stcs(reinterpret_cast<double2 *>(p), make_double2(a, b));
    __stcs(reinterpret_cast<double2 *>(p + 2), make_double2(c, d));
}

// Two FMAs in one instruction slot: (x0, x1) += (l0, l1) * c.  On sm_100a this is one FFMA2 with the
// scalar operand broadcast (R.F32) and an optional negation folded into the packed operand; the two
// accumulators are vertically adjacent tile elements, which is also what the 128-bit stores want.
#ifndef INVGPU_FFMA2
#define INVGPU_FFMA2 0   // measured on B200: the kernels are not issue-bound, FFMA2 buys nothing and its
                         // register-pair constraints add moves (n = 128: 8.5 ms vs 7.7 ms); kept for experiments
#endif
__device__ __forceinline__ void fma2_rows(float &x0, float &x1, float l0, float l1, float c) {
#if !INVGPU_FFMA2
    x0 = fmaf(l0, c, x0);
    x1 = fmaf(l1, c, x1);
    return;
#endif
    unsigned long long acc, ll, cc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ll) : "f"(l0), "f"(l1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c), "f"(c));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(ll), "l"(cc));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(acc));
}
__device__ __forceinline__ void fma2_rows(double &x0, double &x1, double l0, double l1, double c) {
    x0 = fma(l0, c, x0);
    x1 = fma(l1, c, x1);
}

template <typename T> __device__ __forceinline__ T dev_rcp(T x);
template <> __device__ __forceinline__ float dev_rcp<float>(float x) { return __frcp_rn(x); }
template <> __device__ __forceinline__ double dev_rcp<double>(double x) { return 1.0 / x; }
template <typename T> __device__ __forceinline__ T dev_rsqrt(T x);
// one MUFU.RSQ: denormal pivots flush to zero and are flagged (1/sqrt(0) = inf), which is the right
// answer for a matrix whose pivot underflows fp32 anyway
template <> __device__ __forceinline__ float dev_rsqrt<float>(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
template <> __device__ __forceinline__ double dev_rsqrt<double>(double x) { return rsqrt(x); }

// ------------------------------------------------------------------------------------------
// compile-time geometry: TR x TC threads per matrix, register tile SR x SC
// ------------------------------------------------------------------------------------------
template <int N, int TR, int TC, bool PERM>
struct TileGeo {
    static constexpr int SR = N / TR;         // rows per thread
    static constexpr int SC = N / TC;         // columns per thread
    static constexpr int LANES = TR * TC;     // threads per matrix
    static constexpr int P = TR > TC ? TR : TC;
    static constexpr int QR = P / TR, QC = P / TC;
    static_assert(N % (4 * P) == 0, "N must be a multiple of 4 * max(TR, TC)");
    static_assert(P % TR == 0 && P % TC == 0, "TR and TC must divide each other");
    // memory 4-block owned by thread-row ti as its row group gi (likewise columns)
    __host__ __device__ static constexpr int rblock(int gi, int ti) { return (gi / QR) * P + (gi % QR) * TR + ti; }
    __host__ __device__ static constexpr int cblock(int hi, int tj) { return (hi / QC) * P + (hi % QC) * TC + tj; }
    // position of memory index i in the elimination order, and back
    __host__ __device__ static constexpr int logical(int i) {
        return PERM ? P * (4 * ((i / 4) / P) + i % 4) + (i / 4) % P : i;
    }
    __host__ __device__ static constexpr int memory(int l) {
        return PERM ? 4 * (((l / P) / 4) * P + l % P) + (l / P) % 4 : l;
    }
    __host__ __device__ static constexpr int rlog(int s, int ti) { return logical(4 * rblock(s / 4, ti) + s % 4); }
    __host__ __device__ static constexpr int clog(int s, int tj) { return logical(4 * cblock(s / 4, tj) + s % 4); }
    __host__ __device__ static constexpr int rmin(int s) { return rlog(s, 0); }
    __host__ __device__ static constexpr int rmax(int s) { return rlog(s, TR - 1); }
    __host__ __device__ static constexpr int cmin(int s) { return clog(s, 0); }
    __host__ __device__ static constexpr int cmax(int s) { return clog(s, TC - 1); }
    // owner thread coordinate / slot of the pivot with logical index k
    __host__ __device__ static constexpr int rowner(int k) { return ((memory(k) / 4) % P) % TR; }
    __host__ __device__ static constexpr int cowner(int k) { return ((memory(k) / 4) % P) % TC; }
    __host__ __device__ static constexpr int rslot(int k) {
        return 4 * (((memory(k) / 4) / P) * QR + ((memory(k) / 4) % P) / TR) + memory(k) % 4;
    }
    __host__ __device__ static constexpr int cslot(int k) {
        return 4 * (((memory(k) / 4) / P) * QC + ((memory(k) / 4) % P) / TC) + memory(k) % 4;
    }
    // "super-block" B = 4*P consecutive logical indices = one 4-block of every thread row / column.
    // All shared-line traffic is predicated on super-blocks so that what is read was written, even
    // when rows and columns are partitioned differently (TR != TC).
    __host__ __device__ static constexpr bool sb_hi(int B, int k) { return 4 * P * (B + 1) - 1 > k; }   // has an index > k
    __host__ __device__ static constexpr bool sb_lo(int B, int k) { return 4 * P * B <= k; }            // has an index <= k
    static __device__ __forceinline__ bool sb_lo_rt(int B, int k) { return 4 * P * B <= k; }             // same, k known at run time
    // shared-memory words per matrix: 2 column lines (double buffer), the line of 1/L_kk, all N rows
    // of M = L^-1; the stride is 8 mod 32 so that the matrices sharing a warp start 8 banks apart
    static constexpr int LINES = 3 * N;                           // offset of the M rows
    static constexpr int SMEM_WORDS = ((3 * N + N * N + 31) / 32) * 32 + (LANES < 32 ? 8 : 0);
    static constexpr int BLOCK = LANES >= 64 ? LANES : INVGPU_WARP_TIER_BLOCK;       // threads per CTA
    static constexpr int MPB = BLOCK / LANES;                     // matrices per CTA
};

template <int LANES>
__device__ __forceinline__ void tile_sync() {
    if (LANES <= 32) __syncwarp(); else __syncthreads();
}

// The fully unrolled phases are ~100 KB of straight-line code per matrix; warps that drift apart
// each stream it through the instruction cache on their own.  Re-aligning the warps of a CTA every
// few pivots lets them share the fetches (warp tiers only; CTA tiers barrier every step anyway).
#ifndef INVGPU_LOCKSTEP
#define INVGPU_LOCKSTEP 8
#endif
// look-ahead variants of potrf / trtri (next pivot column published before the bulk update): measured
// neutral to slightly negative on B200 (longer live ranges -> spills), kept for experiments
#ifndef INVGPU_LOOKAHEAD
#define INVGPU_LOOKAHEAD 0
#endif
// rolled pivot loops for square thread grids (see TileSpd::potrf_rolled)
#ifndef INVGPU_ROLLED
#define INVGPU_ROLLED 1
#endif
#ifndef INVGPU_LAUUM_SEG
#define INVGPU_LAUUM_SEG 8
#endif
template <int LANES>
__device__ __forceinline__ void tile_lockstep(int k) {
    if (LANES <= 32 && INVGPU_LOCKSTEP > 0 && (k % (INVGPU_LOCKSTEP > 0 ? INVGPU_LOCKSTEP : 1)) == 0) __syncthreads();
}

// ------------------------------------------------------------------------------------------
// the three phases, operating on the register tile a[SR][SC]; all comparisons are in logical order
// ------------------------------------------------------------------------------------------
template <typename T, int N, int TR, int TC, bool PERM>
struct TileSpd {
    using G = TileGeo<N, TR, TC, PERM>;
    static constexpr int SR = G::SR, SC = G::SC;

    static __device__ __forceinline__ int rlog(int s, int ti) { return G::rmin(s) + (PERM ? ti : 4 * ti); }
    static __device__ __forceinline__ int clog(int s, int tj) { return G::cmin(s) + (PERM ? tj : 4 * tj); }
    // shared line offsets (memory order) of row group g / column group h of this thread
    static __device__ __forceinline__ int roff(int g, int ti) { return 4 * G::rblock(g, 0) + 4 * ti; }
    static __device__ __forceinline__ int coff(int h, int tj) { return 4 * G::cblock(h, 0) + 4 * tj; }

    // =====================================================================================
    // Warp-sized groups: phases with LOOK-AHEAD.  The next pivot column (row) is brought up to date,
    // scaled and published before the bulk of the current rank-1 update, so the shuffle -> rsqrt ->
    // scale -> store -> barrier -> load chain of step k+1 overlaps the FMAs of step k.
    // =====================================================================================

    // update of one column slot c by pivot k (rows with a logical index > k somewhere)
    static __device__ __forceinline__ void potrf_update_col(T (&a)[SR][SC], const T (&lr)[SR], const T (&lc)[SC],
                                                            int k, int c, int sk, int tj) {
        if (G::cmax(c) <= k) return;
        #pragma unroll
        for (int r = 0; r < SR; ++r) {
            if (G::rmax(r) <= k) continue;
            if (G::cmin(c) > G::rmax(r)) continue;
            if (c == sk) { if (clog(c, tj) > k) a[r][c] = fma(-lr[r], lc[c], a[r][c]); }
            else a[r][c] = fma(-lr[r], lc[c], a[r][c]);
        }
    }

    // pivot k: column k is up to date.  Fetch d, scale the column into L(:,k), publish it (+ y_k).
    template <bool GP>
    static __device__ __forceinline__ void potrf_publish(T (&a)[SR][SC], T (*z)[2], T *sm, int line_words, int k,
                                                         int ti, int tj, int &info) {
        T *cb = sm + (k & 1) * line_words;
        const int ck = G::cowner(k), sk = G::cslot(k), rk = G::rowner(k), srk = G::rslot(k);
        const int src = (threadIdx.x & 31 & ~(G::LANES - 1)) + rk * TC + ck;
        const T d = __shfl_sync(0xffffffffu, a[srk][sk], src);
        const T rs = dev_rsqrt<T>(d);
        if (info == 0 && !(d > T(0))) info = k + 1;
        if (tj == ck) {
            #pragma unroll
            for (int r = 0; r < SR; ++r)
                if (G::rmax(r) >= k) a[r][sk] *= rs;               // row k itself: a_kk * rs = L_kk
            if (k + 1 < N) {
                #pragma unroll
                for (int g = 0; g < SR / 4; ++g)
                    if (G::sb_hi(g / G::QR, k))
                        st4(cb + roff(g, ti), a[4 * g][sk], a[4 * g + 1][sk], a[4 * g + 2][sk], a[4 * g + 3][sk]);
            }
            if (ti == rk) {
                if (GP) { cb[N] = z[srk][0] * rs; cb[N + 1] = z[srk][1] * rs; }
                else sm[2 * N + G::memory(k)] = rs;                // 1 / L_kk for trtri
            }
        }
    }

    template <bool GP>
    static __device__ __forceinline__ void potrf_la(T (&a)[SR][SC], T (*z)[2], T *sm, int ti, int tj, int &info,
                                                    T &acc_m, T &acc_q) {
        constexpr int LINE = GP ? N + 4 : N;
        potrf_publish<GP>(a, z, sm, LINE, 0, ti, tj, info);
        tile_sync<G::LANES>();
        #pragma unroll
        for (int k = 0; k < N; ++k) {
            tile_lockstep<G::LANES>(k);
            const T *cb = sm + (k & 1) * LINE;
            const int sk = G::cslot(k);
            T ya = T(0), yd = T(0);
            if (GP) {
                ya = cb[N]; yd = cb[N + 1];
                acc_m = fma(ya, yd, acc_m);
                acc_q = fma(ya, ya, acc_q);
            }
            if (k + 1 == N) break;
            T lr[SR], lc[SC];
            #pragma unroll
            for (int g = 0; g < SR / 4; ++g)
                if (G::sb_hi(g / G::QR, k))
                    ld4(cb + roff(g, ti), lr[4 * g], lr[4 * g + 1], lr[4 * g + 2], lr[4 * g + 3]);
            #pragma unroll
            for (int g = 0; g < SC / 4; ++g)
                if (G::sb_hi(g / G::QC, k))
                    ld4(cb + coff(g, tj), lc[4 * g], lc[4 * g + 1], lc[4 * g + 2], lc[4 * g + 3]);
            if (GP) {
                #pragma unroll
                for (int r = 0; r < SR; ++r) {
                    if (G::rmax(r) <= k) continue;
                    z[r][0] = fma(-lr[r], ya, z[r][0]);
                    z[r][1] = fma(-lr[r], yd, z[r][1]);
                }
            }
            const int sk1 = G::cslot(k + 1);
            potrf_update_col(a, lr, lc, k, sk1, sk, tj);          // look-ahead: next pivot column first
            potrf_publish<GP>(a, z, sm, LINE, k + 1, ti, tj, info);
            #pragma unroll
            for (int c = 0; c < SC; ++c)
                if (c != sk1) potrf_update_col(a, lr, lc, k, c, sk, tj);
            tile_sync<G::LANES>();
        }
    }

    // row k of M is complete: finish it and park it in shared memory
    static __device__ __forceinline__ void trtri_publish_row(T (&a)[SR][SC], const T *sm, T *mbuf, int k, int ti, int tj) {
        const int rk = G::rowner(k), srk = G::rslot(k);
        if (ti == rk) {
            T *rb = mbuf + k * N;
            const T rdk = sm[2 * N + G::memory(k)];
            #pragma unroll
            for (int c = 0; c < SC; ++c) {
                if (G::cmin(c) > k) continue;
                if (G::cmax(c) < k) a[srk][c] *= -rdk;
                else {
                    const int col = clog(c, tj);
                    a[srk][c] = col < k ? a[srk][c] * -rdk : (col == k ? rdk : a[srk][c]);
                }
            }
            #pragma unroll
            for (int h = 0; h < SC / 4; ++h)
                if (G::sb_lo(h / G::QC, k))
                    st4(rb + coff(h, tj), a[srk][4 * h], a[srk][4 * h + 1], a[srk][4 * h + 2], a[srk][4 * h + 3]);
        }
    }
    // column k of L (rows > k): publish, then clear it for the accumulation of M
    static __device__ __forceinline__ void trtri_publish_col(T (&a)[SR][SC], T *sm, int k, int ti, int tj) {
        const int ck = G::cowner(k), sk = G::cslot(k);
        if (tj == ck) {
            T *cb = sm + (k & 1) * N;
            #pragma unroll
            for (int g = 0; g < SR / 4; ++g)
                if (G::sb_hi(g / G::QR, k))
                    st4(cb + roff(g, ti), a[4 * g][sk], a[4 * g + 1][sk], a[4 * g + 2][sk], a[4 * g + 3][sk]);
            #pragma unroll
            for (int r = 0; r < SR; ++r) {
                if (G::rmax(r) <= k) continue;
                if (G::rmin(r) > k) a[r][sk] = T(0);
                else a[r][sk] = rlog(r, ti) > k ? T(0) : a[r][sk];
            }
        }
    }
    static __device__ __forceinline__ void trtri_update_row(T (&a)[SR][SC], const T (&lr)[SR], const T (&mr)[SC], int k, int r) {
        if (G::rmax(r) <= k) return;
        #pragma unroll
        for (int c = 0; c < SC; ++c) {
            if (G::cmin(c) > k) continue;
            a[r][c] = fma(lr[r], mr[c], a[r][c]);
        }
    }
    static __device__ __forceinline__ void trtri_la(T (&a)[SR][SC], T *sm, T *mbuf, int ti, int tj) {
        trtri_publish_row(a, sm, mbuf, 0, ti, tj);
        if (N > 1) trtri_publish_col(a, sm, 0, ti, tj);
        tile_sync<G::LANES>();
        #pragma unroll
        for (int k = 0; k + 1 < N; ++k) {
            tile_lockstep<G::LANES>(k);
            const T *cb = sm + (k & 1) * N;
            const T *rb = mbuf + k * N;
            const int rk = G::rowner(k), srk = G::rslot(k);
            T lr[SR], mr[SC];
            #pragma unroll
            for (int g = 0; g < SR / 4; ++g)
                if (G::sb_hi(g / G::QR, k))
                    ld4(cb + roff(g, ti), lr[4 * g], lr[4 * g + 1], lr[4 * g + 2], lr[4 * g + 3]);
            if (ti == rk) lr[srk] = T(0);                          // row k is finished: leave it alone
            #pragma unroll
            for (int h = 0; h < SC / 4; ++h)
                if (G::sb_lo(h / G::QC, k))
                    ld4(rb + coff(h, tj), mr[4 * h], mr[4 * h + 1], mr[4 * h + 2], mr[4 * h + 3]);
            const int srk1 = G::rslot(k + 1);
            trtri_update_row(a, lr, mr, k, srk1);                  // look-ahead: next row of M first
            trtri_publish_row(a, sm, mbuf, k + 1, ti, tj);
            if (k + 2 < N) trtri_publish_col(a, sm, k + 1, ti, tj);
            #pragma unroll
            for (int r = 0; r < SR; ++r)
                if (r != srk1) trtri_update_row(a, lr, mr, k, r);
            tile_sync<G::LANES>();
        }
    }

    // =====================================================================================
    // Square thread grids (TR == TC == P, permuted order): the P consecutive pivots k = P*s + t,
    // t = 0..P-1, all live in register slot (s, s) of the diagonal threads (t, t), and the set of
    // active slots is the same for all of them.  So the pivot loop is ROLLED over t (run-time owner
    // tests, static register indices): N/P small bodies instead of N -- for n = 128 the difference
    // between ~25 thousand and ~1 thousand instructions of code.
    // =====================================================================================
    template <bool GP>
    static __device__ __forceinline__ void potrf_rolled(T (&a)[SR][SC], T (*z)[2], T *sm, int ti, int tj, int &info,
                                                        T &acc_m, T &acc_q) {
        static_assert(TR == TC && PERM, "rolled phases need a square thread grid and the cyclic order");
        constexpr int LINE = GP ? N + 4 : N;
        #pragma unroll
        for (int s = 0; s < SR; ++s) {
            const int kbase = G::P * s;                          // logical index of slot s in thread 0
            const int mbase = 4 * (G::P * (s / 4)) + (s % 4);     // its memory index; thread t adds 4 t
            #pragma unroll 1
            for (int t = 0; t < G::P; ++t) {
                const int k = kbase + t;
                T *cb = sm + (k & 1) * LINE;
                T d, rs;
                if (G::LANES <= 32) {
                    const int src = (threadIdx.x & 31 & ~(G::LANES - 1)) + t * TC + t;
                    d = __shfl_sync(0xffffffffu, a[s][s], src);
                    rs = dev_rsqrt<T>(d);
                    if (tj == t) {
                        #pragma unroll
                        for (int r = s; r < SR; ++r) a[r][s] *= rs;
                        #pragma unroll
                        for (int g = s / 4; g < SR / 4; ++g)
                            st4(cb + roff(g, ti), a[4 * g][s], a[4 * g + 1][s], a[4 * g + 2][s], a[4 * g + 3][s]);
                        if (ti == t) {
                            if (GP) { cb[N] = z[s][0] * rs; cb[N + 1] = z[s][1] * rs; }
                            else sm[2 * N + mbase + 4 * t] = rs;
                        }
                    }
                    tile_sync<G::LANES>();
                } else {
                    if (tj == t) {
                        #pragma unroll
                        for (int g = s / 4; g < SR / 4; ++g)
                            st4(cb + roff(g, ti), a[4 * g][s], a[4 * g + 1][s], a[4 * g + 2][s], a[4 * g + 3][s]);
                        if (ti == t) {                             // the diagonal thread owns the pivot itself
                            const T r0 = dev_rsqrt<T>(a[s][s]);
                            if (GP) { cb[N] = z[s][0] * r0; cb[N + 1] = z[s][1] * r0; }
                            else sm[2 * N + mbase + 4 * t] = r0;
                        }
                    }
                    tile_sync<G::LANES>();
                    d = cb[mbase + 4 * t];
                    rs = dev_rsqrt<T>(d);
                    if (tj == t) {
                        #pragma unroll
                        for (int r = s; r < SR; ++r) a[r][s] *= rs;
                    }
                }
                if (info == 0 && !(d > T(0))) info = k + 1;
                T ya = T(0), yd = T(0);
                if (GP) {
                    ya = cb[N]; yd = cb[N + 1];
                    acc_m = fma(ya, yd, acc_m);
                    acc_q = fma(ya, ya, acc_q);
                }
                if (k + 1 == N) break;
                T lr[SR], lc[SC];
                #pragma unroll
                for (int g = s / 4; g < SR / 4; ++g) {
                    ld4(cb + roff(g, ti), lr[4 * g], lr[4 * g + 1], lr[4 * g + 2], lr[4 * g + 3]);
                    ld4(cb + coff(g, tj), lc[4 * g], lc[4 * g + 1], lc[4 * g + 2], lc[4 * g + 3]);
                    if (G::LANES > 32) {
                        #pragma unroll
                        for (int w = 0; w < 4; ++w) { lr[4 * g + w] *= rs; lc[4 * g + w] *= rs; }
                    }
                }
                #pragma unroll
                for (int r = s & ~1; r < SR; r += 2) {             // row pairs (slot s may be the odd one of its pair)
                    if (GP) {
                        fma2_rows(z[r][0], z[r + 1][0], -lr[r], -lr[r + 1], ya);
                        fma2_rows(z[r][1], z[r + 1][1], -lr[r], -lr[r + 1], yd);
                    }
                    #pragma unroll
                    for (int c = s; c <= r + 1; ++c) {             // lower part: column slot <= row slot
                        if (c == s) { if (tj > t) fma2_rows(a[r][c], a[r + 1][c], -lr[r], -lr[r + 1], lc[c]); }
                        else fma2_rows(a[r][c], a[r + 1][c], -lr[r], -lr[r + 1], lc[c]);
                    }
                }
            }
        }
    }

    static __device__ __forceinline__ void trtri_rolled(T (&a)[SR][SC], T *sm, T *mbuf, int ti, int tj) {
        static_assert(TR == TC && PERM, "rolled phases need a square thread grid and the cyclic order");
        #pragma unroll
        for (int s = 0; s < SR; ++s) {
            const int kbase = G::P * s;
            const int mbase = 4 * (G::P * (s / 4)) + (s % 4);
            #pragma unroll 1
            for (int t = 0; t < G::P; ++t) {
                const int k = kbase + t;
                T *cb = sm + (k & 1) * N;
                T *rb = mbuf + k * N;
                if (ti == t) {                                     // owners of row k: finish and park it
                    const T rdk = sm[2 * N + mbase + 4 * t];
                    #pragma unroll
                    for (int c = 0; c < s; ++c) a[s][c] *= -rdk;   // columns < kbase <= k
                    a[s][s] = tj < t ? a[s][s] * -rdk : (tj == t ? rdk : a[s][s]);
                    #pragma unroll
                    for (int h = 0; h <= s / 4; ++h)
                        st4(rb + coff(h, tj), a[s][4 * h], a[s][4 * h + 1], a[s][4 * h + 2], a[s][4 * h + 3]);
                }
                if (k + 1 == N) break;
                if (tj == t) {                                     // column k of L, rows > k: publish, then clear
                    #pragma unroll
                    for (int g = s / 4; g < SR / 4; ++g)
                        st4(cb + roff(g, ti), a[4 * g][s], a[4 * g + 1][s], a[4 * g + 2][s], a[4 * g + 3][s]);
                    a[s][s] = ti > t ? T(0) : a[s][s];
                    #pragma unroll
                    for (int r = s + 1; r < SR; ++r) a[r][s] = T(0);
                }
                tile_sync<G::LANES>();
                T lr[SR], mr[SC];
                #pragma unroll
                for (int g = s / 4; g < SR / 4; ++g)
                    ld4(cb + roff(g, ti), lr[4 * g], lr[4 * g + 1], lr[4 * g + 2], lr[4 * g + 3]);
                lr[s] = ti > t ? lr[s] : T(0);                     // rows <= k of this slot are finished
                #pragma unroll
                for (int h = 0; h <= s / 4; ++h)
                    ld4(rb + coff(h, tj), mr[4 * h], mr[4 * h + 1], mr[4 * h + 2], mr[4 * h + 3]);
                #pragma unroll
                for (int r = s & ~1; r < SR; r += 2) {
                    const T l0 = r < s ? T(0) : lr[r];             // slot below s: rows already finished
                    #pragma unroll
                    for (int c = 0; c <= s; ++c) fma2_rows(a[r][c], a[r + 1][c], l0, lr[r + 1], mr[c]);
                }
            }
        }
    }

    // ---- potrf ---------------------------------------------------------------------------
    // the line sm[2N + i] receives 1 / L_ii (memory index i) for trtri
    static __device__ __forceinline__ void potrf(T (&a)[SR][SC], T *sm, int ti, int tj, int &info) {
        #pragma unroll
        for (int k = 0; k < N; ++k) {
            tile_lockstep<G::LANES>(k);
            T *cb = sm + (k & 1) * N;                     // double-buffered column line
            const int ck = G::cowner(k), sk = G::cslot(k), rk = G::rowner(k), srk = G::rslot(k);
            T d, rs;
            if (G::LANES <= 32) {
                // warp-sized group: fetch the pivot from the diagonal thread, owners publish L(:,k)
                const int src = (threadIdx.x & 31 & ~(G::LANES - 1)) + rk * TC + ck;
                d = __shfl_sync(0xffffffffu, a[srk][sk], src);
                rs = dev_rsqrt<T>(d);
                if (tj == ck) {
                    #pragma unroll
                    for (int r = 0; r < SR; ++r)
                        if (G::rmax(r) >= k) a[r][sk] *= rs;       // row k itself: a_kk * rs = L_kk
                    if (k + 1 < N) {
                        #pragma unroll
                        for (int g = 0; g < SR / 4; ++g)
                            if (G::sb_hi(g / G::QR, k))
                                st4(cb + roff(g, ti), a[4 * g][sk], a[4 * g + 1][sk], a[4 * g + 2][sk], a[4 * g + 3][sk]);
                    }
                }
                if (k + 1 < N) tile_sync<G::LANES>();
            } else {
                // CTA-sized group: owners publish the raw column, everybody derives the scale
                if (tj == ck) {
                    #pragma unroll
                    for (int g = 0; g < SR / 4; ++g)
                        if (G::sb_hi(g / G::QR, k - 1))
                            st4(cb + roff(g, ti), a[4 * g][sk], a[4 * g + 1][sk], a[4 * g + 2][sk], a[4 * g + 3][sk]);
                }
                tile_sync<G::LANES>();
                d = cb[G::memory(k)];
                rs = dev_rsqrt<T>(d);
                if (tj == ck) {
                    #pragma unroll
                    for (int r = 0; r < SR; ++r)
                        if (G::rmax(r) >= k) a[r][sk] *= rs;
                }
            }
            if (info == 0 && !(d > T(0))) info = k + 1;   // uniform over the matrix' threads
            if (ti == rk && tj == ck) sm[2 * N + G::memory(k)] = rs;
            if (k + 1 == N) break;
            T lr[SR], lc[SC];
            #pragma unroll
            for (int g = 0; g < SR / 4; ++g)
                if (G::sb_hi(g / G::QR, k)) {
                    ld4(cb + roff(g, ti), lr[4 * g], lr[4 * g + 1], lr[4 * g + 2], lr[4 * g + 3]);
                    if (G::LANES > 32) {
                        #pragma unroll
                        for (int w = 0; w < 4; ++w) lr[4 * g + w] *= rs;
                    }
                }
            #pragma unroll
            for (int g = 0; g < SC / 4; ++g)
                if (G::sb_hi(g / G::QC, k)) {
                    ld4(cb + coff(g, tj), lc[4 * g], lc[4 * g + 1], lc[4 * g + 2], lc[4 * g + 3]);
                    if (G::LANES > 32) {
                        #pragma unroll
                        for (int w = 0; w < 4; ++w) lc[4 * g + w] *= rs;
                    }
                }
            #pragma unroll
            for (int r = 0; r < SR; r += 2) {                      // row pairs: one FFMA2 per two elements
                if (G::rmax(r + 1) <= k) continue;
                #pragma unroll
                for (int c = 0; c < SC; ++c) {
                    if (G::cmax(c) <= k) continue;                 // column already final everywhere
                    if (G::cmin(c) > G::rmax(r + 1)) continue;     // strictly upper for every thread
                    if (c == sk) {                                  // slot that holds column k in its owners
                        if (clog(c, tj) > k) fma2_rows(a[r][c], a[r + 1][c], -lr[r], -lr[r + 1], lc[c]);
                    } else {
                        fma2_rows(a[r][c], a[r + 1][c], -lr[r], -lr[r + 1], lc[c]);
                    }
                }
            }
        }
    }

    // ---- potrf with two right-hand sides riding along (fused GP mean / variance) ---------------
    // z[r][0] / z[r][1] hold the vectors A and D for this thread-row's rows (replicated over tj).
    // At step k the diagonal thread publishes y_k = z_k / L_kk next to column k; everybody then does
    // z_i -= L_ik y_k with the row factors it loads anyway, and accumulates  sum_k yA_k yD_k
    // (= A^T M^-1 D) and  sum_k yA_k^2 (= A^T M^-1 A) on the fly: no back substitution, no reduction.
    // Shared line layout: 2 x (N + 4) words.
    static __device__ __forceinline__ void potrf_gp(T (&a)[SR][SC], T (&z)[SR][2], T *sm, int ti, int tj,
                                                    int &info, T &acc_m, T &acc_q) {
        constexpr int LINE = N + 4;
        #pragma unroll
        for (int k = 0; k < N; ++k) {
            tile_lockstep<G::LANES>(k);
            T *cb = sm + (k & 1) * LINE;
            const int ck = G::cowner(k), sk = G::cslot(k), rk = G::rowner(k), srk = G::rslot(k);
            T d, rs;
            if (G::LANES <= 32) {
                const int src = (threadIdx.x & 31 & ~(G::LANES - 1)) + rk * TC + ck;
                d = __shfl_sync(0xffffffffu, a[srk][sk], src);
                rs = dev_rsqrt<T>(d);
                if (tj == ck) {
                    #pragma unroll
                    for (int r = 0; r < SR; ++r)
                        if (G::rmax(r) >= k) a[r][sk] *= rs;
                    if (k + 1 < N) {
                        #pragma unroll
                        for (int g = 0; g < SR / 4; ++g)
                            if (G::sb_hi(g / G::QR, k))
                                st4(cb + roff(g, ti), a[4 * g][sk], a[4 * g + 1][sk], a[4 * g + 2][sk], a[4 * g + 3][sk]);
                    }
                    if (ti == rk) { cb[N] = z[srk][0] * rs; cb[N + 1] = z[srk][1] * rs; }
                }
                tile_sync<G::LANES>();
            } else {
                if (tj == ck) {
                    #pragma unroll
                    for (int g = 0; g < SR / 4; ++g)
                        if (G::sb_hi(g / G::QR, k - 1))
                            st4(cb + roff(g, ti), a[4 * g][sk], a[4 * g + 1][sk], a[4 * g + 2][sk], a[4 * g + 3][sk]);
                    if (ti == rk) {                                // the diagonal thread owns the pivot itself
                        const T r0 = dev_rsqrt<T>(a[srk][sk]);
                        cb[N] = z[srk][0] * r0; cb[N + 1] = z[srk][1] * r0;
                    }
                }
                tile_sync<G::LANES>();
                d = cb[G::memory(k)];
                rs = dev_rsqrt<T>(d);
            }
            if (info == 0 && !(d > T(0))) info = k + 1;
            const T ya = cb[N], yd = cb[N + 1];
            acc_m = fma(ya, yd, acc_m);
            acc_q = fma(ya, ya, acc_q);
            if (k + 1 == N) break;
            T lr[SR], lc[SC];
            #pragma unroll
            for (int g = 0; g < SR / 4; ++g)
                if (G::sb_hi(g / G::QR, k)) {
                    ld4(cb + roff(g, ti), lr[4 * g], lr[4 * g + 1], lr[4 * g + 2], lr[4 * g + 3]);
                    if (G::LANES > 32) {
                        #pragma unroll
                        for (int w = 0; w < 4; ++w) lr[4 * g + w] *= rs;
                    }
                }
            #pragma unroll
            for (int g = 0; g < SC / 4; ++g)
                if (G::sb_hi(g / G::QC, k)) {
                    ld4(cb + coff(g, tj), lc[4 * g], lc[4 * g + 1], lc[4 * g + 2], lc[4 * g + 3]);
                    if (G::LANES > 32) {
                        #pragma unroll
                        for (int w = 0; w < 4; ++w) lc[4 * g + w] *= rs;
                    }
                }
            #pragma unroll
            for (int r = 0; r < SR; r += 2) {
                if (G::rmax(r + 1) <= k) continue;
                fma2_rows(z[r][0], z[r + 1][0], -lr[r], -lr[r + 1], ya);
                fma2_rows(z[r][1], z[r + 1][1], -lr[r], -lr[r + 1], yd);
                #pragma unroll
                for (int c = 0; c < SC; ++c) {
                    if (G::cmax(c) <= k) continue;
                    if (G::cmin(c) > G::rmax(r + 1)) continue;
                    if (c == sk) {
                        if (clog(c, tj) > k) fma2_rows(a[r][c], a[r + 1][c], -lr[r], -lr[r + 1], lc[c]);
                    } else {
                        fma2_rows(a[r][c], a[r + 1][c], -lr[r], -lr[r + 1], lc[c]);
                    }
                }
            }
        }
    }

    // strictly-upper (logical) positions hold by-products of potrf: clear them, the later phases
    // rely on zeros there
    static __device__ __forceinline__ void clear_upper(T (&a)[SR][SC], int ti, int tj) {
        #pragma unroll
        for (int r = 0; r < SR; ++r)
            #pragma unroll
            for (int c = 0; c < SC; ++c) {
                if (G::cmax(c) <= G::rmin(r)) continue;                       // lower (or diagonal) everywhere
                if (G::cmin(c) > G::rmax(r)) a[r][c] = T(0);
                else a[r][c] = clog(c, tj) > rlog(r, ti) ? T(0) : a[r][c];
            }
    }

    // ---- trtri ---------------------------------------------------------------------------
    // In: a = L (lower, zeros above).  Out: row k of M = L^-1 (zero padded) in mbuf + k*N for
    // every k; the register tile is dead afterwards.
    static __device__ __forceinline__ void trtri(T (&a)[SR][SC], T *sm, T *mbuf, int ti, int tj) {
        #pragma unroll
        for (int k = 0; k < N; ++k) {
            tile_lockstep<G::LANES>(k);
            T *cb = sm + (k & 1) * N;                     // column k of L
            T *rb = mbuf + k * N;                         // row k of M (kept for lauum)
            const int ck = G::cowner(k), sk = G::cslot(k), rk = G::rowner(k), srk = G::rslot(k);
            if (ti == rk) {                               // owners of row k finish it: M_kc = -acc_kc / L_kk, M_kk = 1 / L_kk
                const T rdk = sm[2 * N + G::memory(k)];
                #pragma unroll
                for (int c = 0; c < SC; ++c) {
                    if (G::cmin(c) > k) continue;
                    if (G::cmax(c) < k) a[srk][c] *= -rdk;
                    else {
                        const int col = clog(c, tj);
                        a[srk][c] = col < k ? a[srk][c] * -rdk : (col == k ? rdk : a[srk][c]);
                    }
                }
                #pragma unroll
                for (int h = 0; h < SC / 4; ++h)
                    if (G::sb_lo(h / G::QC, k))
                        st4(rb + coff(h, tj), a[srk][4 * h], a[srk][4 * h + 1], a[srk][4 * h + 2], a[srk][4 * h + 3]);
            }
            if (k + 1 == N) break;
            if (tj == ck) {                               // column k of L (rows > k): publish, then clear for the accumulation
                #pragma unroll
                for (int g = 0; g < SR / 4; ++g)
                    if (G::sb_hi(g / G::QR, k))
                        st4(cb + roff(g, ti), a[4 * g][sk], a[4 * g + 1][sk], a[4 * g + 2][sk], a[4 * g + 3][sk]);
                #pragma unroll
                for (int r = 0; r < SR; ++r) {
                    if (G::rmax(r) <= k) continue;
                    if (G::rmin(r) > k) a[r][sk] = T(0);
                    else a[r][sk] = rlog(r, ti) > k ? T(0) : a[r][sk];
                }
            }
            tile_sync<G::LANES>();
            T lr[SR], mr[SC];
            #pragma unroll
            for (int g = 0; g < SR / 4; ++g)
                if (G::sb_hi(g / G::QR, k))
                    ld4(cb + roff(g, ti), lr[4 * g], lr[4 * g + 1], lr[4 * g + 2], lr[4 * g + 3]);
            // entries of the line for rows < k are zeros (cleared upper part); row k itself carries
            // L_kk and must not feed back into the finished row
            if (ti == rk) lr[srk] = T(0);
            #pragma unroll
            for (int h = 0; h < SC / 4; ++h)
                if (G::sb_lo(h / G::QC, k))
                    ld4(rb + coff(h, tj), mr[4 * h], mr[4 * h + 1], mr[4 * h + 2], mr[4 * h + 3]);
            #pragma unroll
            for (int r = 0; r < SR; r += 2) {
                if (G::rmax(r + 1) <= k) continue;
                // a row of the pair that is already finished (<= k) must stay untouched
                const T l0 = G::rmax(r) <= k ? T(0) : lr[r];
                #pragma unroll
                for (int c = 0; c < SC; ++c) {
                    if (G::cmin(c) > k) continue;
                    fma2_rows(a[r][c], a[r + 1][c], l0, lr[r + 1], mr[c]);
                }
            }
        }
    }

    // ---- lauum (full tile) ------------------------------------------------------------------
    // P = M^T M accumulated from the rows of M kept in shared memory: no publishing, no barriers.
    // Every thread ends up with its complete (both triangles) part of A^-1.
    static __device__ __forceinline__ void lauum(T (&a)[SR][SC], const T *mbuf, int ti, int tj) {
        #pragma unroll
        for (int r = 0; r < SR; ++r)
            #pragma unroll
            for (int c = 0; c < SC; ++c) a[r][c] = T(0);
#if INVGPU_LAUUM_SEG > 0
        // Rolled in segments of INVGPU_LAUUM_SEG steps: inside a segment the (tiny) loop body is pruned
        // for the segment's last step, so the code stays in the instruction cache; the rows of M are
        // zero-padded, which makes the few extra products vanish.
        constexpr int SEG = INVGPU_LAUUM_SEG;
        #pragma unroll
        for (int k0 = 0; k0 < N; k0 += SEG) {
            const int k1 = (k0 + SEG < N ? k0 + SEG : N) - 1;      // last step of the segment
            #pragma unroll 1
            for (int k = k0; k <= k1; ++k) {
                const T *rb = mbuf + k * N;
                T mi[SR], mc[SC];
                #pragma unroll
                for (int g = 0; g < SR / 4; ++g)
                    if (G::sb_lo(g / G::QR, k0))                   // written for every step of the segment
                        ld4(rb + roff(g, ti), mi[4 * g], mi[4 * g + 1], mi[4 * g + 2], mi[4 * g + 3]);
                    else if (G::sb_lo(g / G::QR, k1)) {
                        #pragma unroll
                        for (int w = 0; w < 4; ++w) mi[4 * g + w] = T(0);
                        if (G::sb_lo_rt(g / G::QR, k))
                            ld4(rb + roff(g, ti), mi[4 * g], mi[4 * g + 1], mi[4 * g + 2], mi[4 * g + 3]);
                    }
                #pragma unroll
                for (int h = 0; h < SC / 4; ++h)
                    if (G::sb_lo(h / G::QC, k0))
                        ld4(rb + coff(h, tj), mc[4 * h], mc[4 * h + 1], mc[4 * h + 2], mc[4 * h + 3]);
                    else if (G::sb_lo(h / G::QC, k1)) {
                        #pragma unroll
                        for (int w = 0; w < 4; ++w) mc[4 * h + w] = T(0);
                        if (G::sb_lo_rt(h / G::QC, k))
                            ld4(rb + coff(h, tj), mc[4 * h], mc[4 * h + 1], mc[4 * h + 2], mc[4 * h + 3]);
                    }
                #pragma unroll
                for (int r = 0; r < SR; r += 2) {
                    if (G::rmin(r) > k1) continue;
                    const T m1 = G::rmin(r + 1) > k1 ? T(0) : mi[r + 1];   // not loaded yet -> contributes nothing
                    #pragma unroll
                    for (int c = 0; c < SC; ++c) {
                        if (G::cmin(c) > k1) continue;
                        fma2_rows(a[r][c], a[r + 1][c], mi[r], m1, mc[c]);
                    }
                }
            }
        }
#else
        #pragma unroll
        for (int k = 0; k < N; ++k) {
            tile_lockstep<G::LANES>(k);
            const T *rb = mbuf + k * N;
            T mi[SR], mc[SC];
            #pragma unroll
            for (int g = 0; g < SR / 4; ++g)
                if (G::sb_lo(g / G::QR, k))
                    ld4(rb + roff(g, ti), mi[4 * g], mi[4 * g + 1], mi[4 * g + 2], mi[4 * g + 3]);
            #pragma unroll
            for (int h = 0; h < SC / 4; ++h)
                if (G::sb_lo(h / G::QC, k))
                    ld4(rb + coff(h, tj), mc[4 * h], mc[4 * h + 1], mc[4 * h + 2], mc[4 * h + 3]);
            #pragma unroll
            for (int r = 0; r < SR; r += 2) {
                if (G::rmin(r) > k) continue;
                const T m1 = G::rmin(r + 1) > k ? T(0) : mi[r + 1];
                #pragma unroll
                for (int c = 0; c < SC; ++c) {
                    if (G::cmin(c) > k) continue;
                    fma2_rows(a[r][c], a[r + 1][c], mi[r], m1, mc[c]);
                }
            }
        }
#endif
    }
};

// Calls into the rolled phases only when the geometry allows them (keeps their static_asserts
// out of rectangular instantiations).
template <typename K, bool GP, typename T, int SR, int SC>
__device__ __forceinline__ void tile_potrf_rolled(T (&a)[SR][SC], T (*z)[2], T *sm, int ti, int tj, int &info, T &m, T &q) {
    if constexpr (SR == SC) K::template potrf_rolled<GP>(a, z, sm, ti, tj, info, m, q);
}
template <typename K, typename T, int SR, int SC>
__device__ __forceinline__ void tile_trtri_rolled(T (&a)[SR][SC], T *sm, T *mbuf, int ti, int tj) {
    if constexpr (SR == SC) K::trtri_rolled(a, sm, mbuf, ti, tj);
}

// Natural-order spotrf info of one matrix, computed by a single thread using the (about to be
// NaN-filled) output matrix as scratch.  Only reached for matrices the fast path flagged.
template <typename T>
__device__ __noinline__ int exact_potrf_info_inplace(T *__restrict__ w, int n) {
    for (int j = 0; j < n; ++j) {
        T d = w[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) { const T u = w[(size_t)j * n + k]; d = fma(-u, u, d); }
        if (!(d > T(0))) return j + 1;
        const T s = dev_sqrt(d);
        w[(size_t)j * n + j] = s;
        for (int i = j + 1; i < n; ++i) {            // u_ji = (a_ji - sum_k u_kj u_ki) / u_jj, stored at (row j, col i)
            T v = w[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) v = fma(-w[(size_t)j * n + k], w[(size_t)i * n + k], v);
            w[(size_t)i * n + j] = v / s;
        }
    }
    return 0;
}

template <typename T>
__device__ __forceinline__ int exact_potrf_info(const T *__restrict__ src, T *__restrict__ w, int n) {
    for (int j = 0; j < n; ++j)
        for (int i = 0; i <= j; ++i) w[(size_t)j * n + i] = src[(size_t)j * n + i];   // upper triangle
    return exact_potrf_info_inplace<T>(w, n);
}

// Loads the logical-lower positions of the register tile from the UPPER triangle of a dense
// column-major matrix (sub-blocks below the diagonal come from their transposed twins).
template <typename T, int N, int TR, int TC, bool PERM>
__device__ __forceinline__ void tile_load_upper(T (&a)[N / TR][N / TC], const T *__restrict__ src, int ti, int tj) {
    using G = TileGeo<N, TR, TC, PERM>;
    constexpr int SR = G::SR, SC = G::SC;
    // ---- load: logical-lower positions only, always from the UPPER triangle of the input ----
    // sub-block (g, h): memory block row br, block column bc.  If br < bc the block itself is in
    // the upper triangle; otherwise its transposed twin (bc, br) is.  x[p][q] = column p, row q
    // of whichever 4x4 block is fetched.
    #pragma unroll
    for (int g = 0; g < SR / 4; ++g) {
        #pragma unroll
        for (int h = 0; h < SC / 4; ++h) {
            // does any thread need any element of this sub-block (logical row >= logical column)?
            if (G::cmin(4 * h) > G::rmax(4 * g + 3)) {
                #pragma unroll
                for (int w = 0; w < 4; ++w)
                    #pragma unroll
                    for (int v = 0; v < 4; ++v) a[4 * g + w][4 * h + v] = T(0);
                continue;
            }
            const int br = G::rblock(g, 0) + ti, bc = G::cblock(h, 0) + tj;
            if (G::rblock(g, 0) > G::cblock(h, TC - 1)) {
                // below the diagonal for every thread: fetch the transposed twin (bc, br); column p
                // of the twin is row p of this block
                #pragma unroll
                for (int p = 0; p < 4; ++p)
                    ldg4(src + (size_t)(4 * br + p) * N + 4 * bc, a[4 * g + p][4 * h], a[4 * g + p][4 * h + 1],
                         a[4 * g + p][4 * h + 2], a[4 * g + p][4 * h + 3]);
            } else if (G::rblock(g, TR - 1) < G::cblock(h, 0)) {
                // above the diagonal for every thread: the block itself
                #pragma unroll
                for (int p = 0; p < 4; ++p)
                    ldg4(src + (size_t)(4 * bc + p) * N + 4 * br, a[4 * g][4 * h + p], a[4 * g + 1][4 * h + p],
                         a[4 * g + 2][4 * h + p], a[4 * g + 3][4 * h + p]);
            } else {
                const bool twin = br >= bc;                   // per-thread choice
                const int colb = twin ? br : bc, rowb = twin ? bc : br;
                T x[4][4];
                #pragma unroll
                for (int p = 0; p < 4; ++p)
                    ldg4(src + (size_t)(4 * colb + p) * N + 4 * rowb, x[p][0], x[p][1], x[p][2], x[p][3]);
                #pragma unroll
                for (int w = 0; w < 4; ++w)
                    #pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        // element (row 4br+w, col 4bc+v): twin -> A(4bc+v, 4br+w) = x[w][v]; else x[v][w].
                        // on the diagonal block (br == bc) only v <= w is in the upper triangle of the twin.
                        T e = twin ? x[w][v] : x[v][w];
                        if (v > w) e = (br == bc) ? x[v][w] : e;
                        a[4 * g + w][4 * h + v] = e;
                    }
            }
        }
    }

}

// ------------------------------------------------------------------------------------------
// kernel: dense column-major matrices of order exactly N (16-byte aligned).
// STAGES: 7 = inverse (PERM order allowed), 1 = factor only (PERM must be false).
// ------------------------------------------------------------------------------------------
template <typename T, int N, int TR, int TC, bool PERM, typename IO, int STAGES, int MINB>
__global__ void __launch_bounds__((TileGeo<N, TR, TC, PERM>::BLOCK), MINB)
tile_spd_kernel(IO io, i64 batch, int *__restrict__ info) {
    using G = TileGeo<N, TR, TC, PERM>;
    using K = TileSpd<T, N, TR, TC, PERM>;
    constexpr int SR = G::SR, SC = G::SC;
    static_assert(!(PERM && STAGES != SPD_INVERSE), "the factor depends on the elimination order");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);

    const int grp = threadIdx.x / G::LANES;
    const int lane = threadIdx.x % G::LANES;
    const int ti = lane / TC, tj = lane % TC;
    T *sm = smem + grp * G::SMEM_WORDS;

    #pragma unroll 1
    for (i64 base = (i64)blockIdx.x * G::MPB; base < batch; base += (i64)gridDim.x * G::MPB) {
        const i64 m = base + grp;
        const bool valid = m < batch;
        const T *__restrict__ src = io.src(valid ? m : batch - 1);

        T a[SR][SC];
        tile_load_upper<T, N, TR, TC, PERM>(a, src, ti, tj);

        int st = 0;
#ifndef INVGPU_DEBUG_SKIP_COMPUTE      // debug: memory-pattern floor of the kernel (load + store only)
        if (TR == TC && PERM && INVGPU_ROLLED) { T dm = T(0), dq = T(0); tile_potrf_rolled<K, false>(a, (T(*)[2]) nullptr, sm, ti, tj, st, dm, dq); }
        else if (G::LANES <= 32 && INVGPU_LOOKAHEAD) { T dm = T(0), dq = T(0); K::template potrf_la<false>(a, nullptr, sm, ti, tj, st, dm, dq); }
        else K::potrf(a, sm, ti, tj, st);
#ifndef INVGPU_DEBUG_ONLY_POTRF
        if (STAGES & SPD_TRTRI) {
            K::clear_upper(a, ti, tj);
            tile_sync<G::LANES>();
            if (TR == TC && PERM && INVGPU_ROLLED) tile_trtri_rolled<K>(a, sm, sm + G::LINES, ti, tj);
            else if (G::LANES <= 32 && INVGPU_LOOKAHEAD) K::trtri_la(a, sm, sm + G::LINES, ti, tj);
            else K::trtri(a, sm, sm + G::LINES, ti, tj);
        }
#ifndef INVGPU_DEBUG_NO_LAUUM
        if (STAGES & SPD_LAUUM) { tile_sync<G::LANES>(); K::lauum(a, sm + G::LINES, ti, tj); }
#endif
#endif
#endif
        tile_sync<G::LANES>();                       // shared lines are reused by the next matrix

        if (!valid) continue;
        T *__restrict__ dst = io.dst(m);
        if (st) {
            if (PERM) {                               // report LAPACK's (natural-order) index
                if (lane == 0) { st = exact_potrf_info<T>(src, dst, N); if (st == 0) st = N; }
                if (G::LANES <= 32) st = __shfl_sync(__activemask(), st, (threadIdx.x & 31 & ~(G::LANES - 1)));
                else { if (lane == 0) sm[0] = (T)st; __syncthreads(); st = (int)sm[0]; __syncthreads(); }
            }
            #pragma unroll
            for (int r = 0; r < SR; ++r)
                #pragma unroll
                for (int c = 0; c < SC; ++c) a[r][c] = dev_nan<T>();
        }
        if (lane == 0 && info) info[m] = st;
        // ---- store ----
        #pragma unroll
        for (int g = 0; g < SR / 4; ++g) {
            #pragma unroll
            for (int h = 0; h < SC / 4; ++h) {
                const int br = G::rblock(g, 0) + ti, bc = G::cblock(h, 0) + tj;
                #pragma unroll
                for (int v = 0; v < 4; ++v) {
                    T x[4];
                    #pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        if (STAGES & SPD_LAUUM) x[w] = a[4 * g + w][4 * h + v];
                        else {                                     // factor: lower triangle (natural order), zero above
                            const int row = 4 * br + w, col = 4 * bc + v;
                            x[w] = (row >= col || st) ? a[4 * g + w][4 * h + v] : T(0);
                        }
                    }
                    stg4(dst + (size_t)(4 * bc + v) * N + 4 * br, x[0], x[1], x[2], x[3]);
                }
            }
        }
    }
}


// ------------------------------------------------------------------------------------------
// fused GP mean / variance kernel (reference src/gauss_bench.cu:127-265, 275-409 in one launch):
// B is loaded with diag(C) added, factored with the two right-hand sides riding along, and the
// scalars fall out of the factorisation.  Nothing but the scalars is ever written.
// scratch: N*N words per matrix slot (grid * MPB slots), used only to recompute LAPACK's
// natural-order info for a flagged (non-SPD) matrix.
// ------------------------------------------------------------------------------------------
template <typename T, int N, int TR, int TC, int MINB>
__global__ void __launch_bounds__((TileGeo<N, TR, TC, true>::BLOCK), MINB)
tile_gp_kernel(GpIO<T> io, i64 batch, int *__restrict__ info, T *__restrict__ scratch) {
    using G = TileGeo<N, TR, TC, true>;
    using K = TileSpd<T, N, TR, TC, true>;
    constexpr int SR = G::SR, SC = G::SC;
    constexpr int WORDS = 2 * (N + 4) + (G::LANES < 32 ? 8 : 0) + ((2 * (N + 4)) % 32 == 0 ? 0 : 32 - (2 * (N + 4)) % 32);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);

    const int grp = threadIdx.x / G::LANES;
    const int lane = threadIdx.x % G::LANES;
    const int ti = lane / TC, tj = lane % TC;
    T *sm = smem + grp * WORDS;

    #pragma unroll 1
    for (i64 base = (i64)blockIdx.x * G::MPB; base < batch; base += (i64)gridDim.x * G::MPB) {
        const i64 m = base + grp;
        const bool valid = m < batch;
        const i64 mm = valid ? m : batch - 1;
        const T *__restrict__ src = io.b + mm * (i64)(N * N);
        T a[SR][SC];
        tile_load_upper<T, N, TR, TC, true>(a, src, ti, tj);
        // + diag(C) on load: only threads holding a diagonal sub-block see diagonal elements
        const T *__restrict__ cv = io.c + mm * N;
        #pragma unroll
        for (int g = 0; g < SR / 4; ++g)
            #pragma unroll
            for (int h = 0; h < SC / 4; ++h) {
                if (G::rblock(g, 0) > G::cblock(h, TC - 1) || G::rblock(g, TR - 1) < G::cblock(h, 0)) continue;
                const int br = G::rblock(g, 0) + ti, bc = G::cblock(h, 0) + tj;
                if (br == bc) {
                    T c0, c1, c2, c3;
                    ldg4(cv + 4 * br, c0, c1, c2, c3);
                    a[4 * g][4 * h] += c0; a[4 * g + 1][4 * h + 1] += c1;
                    a[4 * g + 2][4 * h + 2] += c2; a[4 * g + 3][4 * h + 3] += c3;
                }
            }
        T z[SR][2];
        const T *__restrict__ av = io.a + mm * N;
        const T *__restrict__ dv = (io.d ? io.d : io.a) + mm * N;
        #pragma unroll
        for (int g = 0; g < SR / 4; ++g) {
            const int off = 4 * (G::rblock(g, 0) + ti);
            ldg4(av + off, z[4 * g][0], z[4 * g + 1][0], z[4 * g + 2][0], z[4 * g + 3][0]);
            ldg4(dv + off, z[4 * g][1], z[4 * g + 1][1], z[4 * g + 2][1], z[4 * g + 3][1]);
        }

        int st = 0;
        T acc_m = T(0), acc_q = T(0);
        if (TR == TC && INVGPU_ROLLED) tile_potrf_rolled<K, true>(a, z, sm, ti, tj, st, acc_m, acc_q);
        else if (G::LANES <= 32 && INVGPU_LOOKAHEAD) K::template potrf_la<true>(a, z, sm, ti, tj, st, acc_m, acc_q);
        else K::potrf_gp(a, z, sm, ti, tj, st, acc_m, acc_q);
        tile_sync<G::LANES>();

        if (!valid) continue;
        if (st) {                                     // rare: report LAPACK's natural-order index
            T *w = scratch + ((i64)blockIdx.x * G::MPB + grp) * (i64)(N * N);
            if (lane == 0) {
                for (int j = 0; j < N; ++j) {
                    for (int i = 0; i < j; ++i) w[(size_t)j * N + i] = src[(size_t)j * N + i];
                    w[(size_t)j * N + j] = src[(size_t)j * N + j] + cv[j];
                }
                st = exact_potrf_info_inplace<T>(w, N);
                if (st == 0) st = N;
                if (io.means) io.means[m] = dev_nan<T>();
                if (io.variances) io.variances[m] = dev_nan<T>();
                if (info) info[m] = st;
            }
            continue;
        }
        if (lane == 0) {
            if (io.means) io.means[m] = acc_m;
            if (io.variances) io.variances[m] = io.e[m] - acc_q;
            if (info) info[m] = 0;
        }
    }
}

}  // namespace invgpu

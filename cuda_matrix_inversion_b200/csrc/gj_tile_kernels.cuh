// gj_tile_kernels.cuh -- general inverse by Gauss-Jordan with partial pivoting on a 2-D REGISTER TILE.
//
// Replaces the reference's Gauss-Jordan kernel loop (src/gauss/batched_invert.cu:17-95: pivotRow /
// normalizeRow / transform_matrix, 3N launches) and its cuBLAS getrf/getriBatched path
// (src/gauss/inverse_gpu.cu:24-50) for n <= 128, and supersedes the lane = row kernel of gj_kernels.cuh
// wherever it is instantiated: the 2-D tile needs (SR + SC) operand words per thread and pivot instead of
// N, and updates with packed FFMA2 -- the same machinery as the SPD sweep (sweep_kernels.cuh).
//
// Layout: TR x TC threads per matrix, 4x4 sub-blocks dealt cyclically, FULL tile (no symmetry), held as
// vertical pairs.  Lanes are tj-major (lane = tj TR + ti): the TR owners of a pivot COLUMN are consecutive
// lanes, so the pivot search is a shuffle arg-max among them.
//
// One step (column k, natural order -> `info` is sgetrf's without further ado):
//   A  owners of column k: arg-max of |a_ik| over the rows that have not been pivots (first maximum wins,
//      like isamax / the oracle), publish the raw column with z_p := -1, the pivot, p; restart the column.
//   -- barrier (CTA-sized groups only; a warp-sized group broadcasts p and the pivot by shuffle)
//   C  owners of row p: publish the raw row with row_k := 1 and restart it.  The register slot of row p is
//      only known at run time: a `switch` over the SR slots keeps every register index static.
//   -- barrier
//   E  everybody:  a_ic += z_i * y_c  with  y_c = -row_c / pivot.  One formula covers the four cases of the
//      in-place algorithm:  i != p, c != k: a_ic - f_i row_c / piv ;  i = p: (-1)(-row_c / piv) = row_c / piv ;
//      c = k: f_i (-1 / piv) ;  i = p, c = k: 1 / piv.
// Rows are never swapped (implicit pivoting).  With rowOfCol[c] = pivot row of column c and colOfRow[r] =
// the column row r was pivot for, the registers finally hold  M[r][c] = Ainv[colOfRow[r]][rowOfCol[c]]
// and every thread scatters its elements accordingly.
//
// Runtime order n <= N: the matrix is embedded as blockdiag(A, I); padded rows can only win padded
// columns, so pivots, flags and results of A are unaffected.
// info: k (1-based) if no non-zero pivot exists for column k (sgetrf's "U(k,k) is exactly zero"; a NaN
// column counts as singular).  Flagged outputs are NaN.
#pragma once

#include "sweep_kernels.cuh"

namespace invgpu {

template <typename T, int N, int TR, int TC>
struct GjtGeo {
    static constexpr int SR = N / TR, SC = N / TC, H = SR / 2, NGR = SR / 4, NGC = SC / 4;
    static constexpr int LANES = TR * TC;
    static constexpr int BLOCK = LANES >= 64 ? LANES : INVGPU_WARP_TIER_BLOCK;
    static constexpr int MPB = BLOCK / LANES;
    static_assert(N % (4 * TR) == 0 && N % (4 * TC) == 0, "N must be a multiple of 4 TR and 4 TC");
    static_assert(TR <= 32 && (TR & (TR - 1)) == 0, "the owners of a column must fit one warp");
    // per buffer: column line z[N] | row line y[N] | p, pivot, -, -
    static constexpr int LINE = 2 * N + 4;
    // per matrix: two buffers, then rowOfCol[N], colOfRow[N] (ints stored in T-sized words)
    static constexpr int WORDS = ((2 * LINE + 2 * N * (int)sizeof(int) / (int)sizeof(T) + 31) / 32) * 32 + (LANES < 32 ? 8 : 0);
};

// publish one row of register pair I of the tile (values of this thread's columns) and restart it; which half
// of the pair is a run-time bit, so that the switch below has SR / 2 cases (code size is what limits this
// kernel: profiles/r1_gj_tile64_summary.md, stall reason no_instruction)
template <typename T, int N, int TR, int TC, int I>
__device__ __forceinline__ void gjt_publish_row(Pair2<T> (&ap)[N / TR / 2][N / TC], T *yl, int tj, bool odd) {
    using G = GjtGeo<T, N, TR, TC>;
    using PR = Pair2<T>;
    #pragma unroll
    for (int h = 0; h < G::NGC; ++h) {
        T e[4];
        #pragma unroll
        for (int v = 0; v < 4; ++v) {
            PR &p = ap[I][4 * h + v];
            const T lo = p.lo(), hi = p.hi();
            e[v] = odd ? hi : lo;
            p = PR::make(odd ? lo : T(0), odd ? T(0) : hi);
        }
        st4(yl + 4 * (TC * h + tj), e[0], e[1], e[2], e[3]);
    }
}

template <typename T, int N, int TR, int TC, int I>
struct GjtRowSwitch {
    static __device__ __forceinline__ void run(Pair2<T> (&ap)[N / TR / 2][N / TC], T *yl, int tj, int slot) {
        if ((slot >> 1) == I) gjt_publish_row<T, N, TR, TC, I>(ap, yl, tj, slot & 1);
        else if constexpr (I + 1 < N / TR / 2) GjtRowSwitch<T, N, TR, TC, I + 1>::run(ap, yl, tj, slot);
    }
};

// one elimination step: column block q (run time), position W inside the block and column group HC (static)
template <typename T, int N, int TR, int TC, int HC, int W>
__device__ __forceinline__ void gjt_step(Pair2<T> (&ap)[N / TR / 2][N / TC], T *sm, int ti, int tj, int q, unsigned &done, int &st) {
    using G = GjtGeo<T, N, TR, TC>;
    using PR = Pair2<T>;
    constexpr int sc = 4 * HC + W;                                 // column slot
    const int k = 4 * q + W;                                       // column index
    const int ck = q % TC;                                         // owner thread column
    T *zl = sm + (k & 1) * G::LINE, *yl = zl + N, *meta = zl + 2 * N;
    int *row_of_col = reinterpret_cast<int *>(sm + 2 * G::LINE), *col_of_row = row_of_col + N;

    // ---- A: pivot search among the owners of column k
    T best = T(-1), bval = T(0);
    int brow = N;
    if (tj == ck) {
        #pragma unroll
        for (int s = 0; s < G::SR; ++s) {
            const T v = (s % 2) ? ap[s / 2][sc].hi() : ap[s / 2][sc].lo();
            const T av = dev_abs(v);
            if (!((done >> s) & 1u) && av > best) { best = av; bval = v; brow = 4 * (TR * (s / 4) + ti) + s % 4; }
        }
    }
    #pragma unroll
    for (int o = TR / 2; o > 0; o >>= 1) {                         // the owners are TR consecutive lanes
        const T ob = __shfl_xor_sync(0xffffffffu, best, o);
        const T ov = __shfl_xor_sync(0xffffffffu, bval, o);
        const int orow = __shfl_xor_sync(0xffffffffu, brow, o);
        if (ob > best || (ob == best && orow < brow)) { best = ob; bval = ov; brow = orow; }
    }
    if (tj == ck) {
        #pragma unroll
        for (int g = 0; g < G::NGR; ++g) {                         // raw column, restart the slot
            sts_pair(zl + 4 * (TR * g + ti), ap[2 * g][sc]);
            sts_pair(zl + 4 * (TR * g + ti) + 2, ap[2 * g + 1][sc]);
            ap[2 * g][sc].clear(); ap[2 * g + 1][sc].clear();
        }
        if (brow < N && ((brow >> 2) % TR) == ti) sts_one(zl + brow, T(-1));       // same thread wrote that word just above
        if (ti == 0) {
            if (G::LANES > 32) { meta[0] = (T)brow; meta[1] = bval; }
            if (brow < N) { row_of_col[k] = brow; col_of_row[brow] = k; }
            else row_of_col[k] = N;
        }
    }
    // ---- C: the owners of row p publish it (register slot known only now) and restart it.
    // Warp-sized groups learn p and the pivot by shuffle from the first owner lane: one barrier per step;
    // CTA-sized groups go through shared memory and a barrier.
    int p;
    T piv;
    if (G::LANES <= 32) {
        const int srcl = (threadIdx.x & 31 & ~(G::LANES - 1)) + ck * TR;
        p = __shfl_sync(0xffffffffu, brow, srcl);
        piv = __shfl_sync(0xffffffffu, bval, srcl);
    } else {
        tile_sync<G::LANES>();
        p = (int)meta[0];
        piv = meta[1];
    }
    if (st == 0 && !(dev_abs(piv) > T(0))) st = k + 1;             // uniform inside the group
    if (p < N && ((p >> 2) % TR) == ti) {
        const int slot = 4 * ((p >> 2) / TR) + (p & 3);
        GjtRowSwitch<T, N, TR, TC, 0>::run(ap, yl, tj, slot);
        done |= 1u << slot;
        if (tj == ck) sts_one(yl + k, T(1));                       // row_k := 1 (that slot was restarted with the column)
    }
    tile_sync<G::LANES>();

    // ---- E: a_ic += z_i * y_c,  y_c = -row_c / pivot
    const T nrp = T(-1) / piv;
    PR x[G::H];
    T y[G::SC];
    #pragma unroll
    for (int g = 0; g < G::NGR; ++g) {
        T x0, x1, x2, x3;
        ld4(zl + 4 * (TR * g + ti), x0, x1, x2, x3);
        x[2 * g] = PR::make(x0, x1); x[2 * g + 1] = PR::make(x2, x3);
    }
    #pragma unroll
    for (int h = 0; h < G::NGC; ++h) {
        ld4(yl + 4 * (TC * h + tj), y[4 * h], y[4 * h + 1], y[4 * h + 2], y[4 * h + 3]);
        #pragma unroll
        for (int v = 0; v < 4; ++v) y[4 * h + v] *= nrp;
    }
    #pragma unroll
    for (int i = 0; i < G::H; ++i)
        #pragma unroll
        for (int c = 0; c < G::SC; ++c) ap[i][c].fma_bcast(x[i], y[c]);
}

// the four columns of every block of column group HC, blocks rolled over their owner
template <typename T, int N, int TR, int TC, int HC>
struct GjtGroups {
    static __device__ __forceinline__ void run(Pair2<T> (&ap)[N / TR / 2][N / TC], T *sm, int ti, int tj, unsigned &done, int &st) {
        #pragma unroll 1
        for (int t = 0; t < TC; ++t) {
            const int q = TC * HC + t;
            gjt_step<T, N, TR, TC, HC, 0>(ap, sm, ti, tj, q, done, st);
            gjt_step<T, N, TR, TC, HC, 1>(ap, sm, ti, tj, q, done, st);
            gjt_step<T, N, TR, TC, HC, 2>(ap, sm, ti, tj, q, done, st);
            gjt_step<T, N, TR, TC, HC, 3>(ap, sm, ti, tj, q, done, st);
        }
        if constexpr (HC + 1 < N / TC / 4) GjtGroups<T, N, TR, TC, HC + 1>::run(ap, sm, ti, tj, done, st);
    }
};

template <typename T, int N, int TR, int TC, typename IO, int MINB>
__global__ void __launch_bounds__((GjtGeo<T, N, TR, TC>::BLOCK), MINB)
gj_tile_kernel(IO io, int n, i64 batch, int *__restrict__ info) {
    using G = GjtGeo<T, N, TR, TC>;
    using PR = Pair2<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);

    const int grp = threadIdx.x / G::LANES;
    const int lane = threadIdx.x % G::LANES;
    const int ti = lane % TR, tj = lane / TR;                      // tj-major: the owners of a column are consecutive lanes
    T *sm = smem + (size_t)grp * G::WORDS;
    const int *row_of_col = reinterpret_cast<const int *>(sm + 2 * G::LINE), *col_of_row = row_of_col + N;

    #pragma unroll 1
    for (i64 base = (i64)blockIdx.x * G::MPB; base < batch; base += (i64)gridDim.x * G::MPB) {
        const i64 m = base + grp;
        const bool valid = m < batch;
        const T *__restrict__ src = io.src(valid ? m : batch - 1);

        PR ap[G::H][G::SC];                                        // ap[i][c] = (A(2i, c), A(2i+1, c)) of this thread's tile
        #pragma unroll
        for (int h = 0; h < G::NGC; ++h)
            #pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int c = 4 * (TC * h + tj) + v;
                #pragma unroll
                for (int i = 0; i < G::H; ++i) {
                    const int r = 4 * (TR * (i / 2) + ti) + 2 * (i % 2);
                    T e0 = (r == c) ? T(1) : T(0), e1 = (r + 1 == c) ? T(1) : T(0);   // identity padding outside n
                    if (c < n) {
                        if (r < n) e0 = __ldcs(src + (size_t)c * n + r);
                        if (r + 1 < n) e1 = __ldcs(src + (size_t)c * n + r + 1);
                    }
                    ap[i][c - 4 * (TC * h + tj) + 4 * h] = PR::make(e0, e1);
                }
            }

        unsigned done = 0;
        int st = 0;
        GjtGroups<T, N, TR, TC, 0>::run(ap, sm, ti, tj, done, st);
        tile_sync<G::LANES>();                                     // permutation arrays complete and visible

        if (valid) {
            if (lane == 0 && info) info[m] = (st > n) ? 0 : st;    // a "singular" padded column cannot happen; guard anyway
            T *__restrict__ dst = io.dst(m);
            const bool bad = st != 0 && st <= n;
            int ocol[G::SC];
            #pragma unroll
            for (int c = 0; c < G::SC; ++c) ocol[c] = row_of_col[4 * (TC * (c / 4) + tj) + c % 4];
            #pragma unroll
            for (int i = 0; i < G::H; ++i) {
                const int r = 4 * (TR * (i / 2) + ti) + 2 * (i % 2);
                const int orow0 = col_of_row[r], orow1 = col_of_row[r + 1];
                #pragma unroll
                for (int c = 0; c < G::SC; ++c) {
                    if (bad) {                                     // natural positions, all NaN
                        const int cc = 4 * (TC * (c / 4) + tj) + c % 4;
                        if (cc < n && r < n) dst[(size_t)cc * n + r] = dev_nan<T>();
                        if (cc < n && r + 1 < n) dst[(size_t)cc * n + r + 1] = dev_nan<T>();
                    } else if (ocol[c] < n) {
                        if (orow0 < n) __stcs(dst + (size_t)ocol[c] * n + orow0, ap[i][c].lo());
                        if (orow1 < n) __stcs(dst + (size_t)ocol[c] * n + orow1, ap[i][c].hi());
                    }
                }
            }
        }
        tile_sync<G::LANES>();                                     // lines and permutation arrays are reused by the next matrix
    }
}

// ==========================================================================================
// General inverse for n = 16 / 32: COLUMN-SPLIT lanes, the whole matrix in registers, no shared-memory
// broadcast.  L = N / CL consecutive lanes own CL consecutive columns each (all N rows: N x CL registers),
// which makes Gauss-Jordan with implicit partial pivoting cheap on both sides:
//   * the pivot column k lives in ONE lane with static row indices: the arg-max is a private scan, the
//     multipliers reach the other lanes by N shuffles of a statically named register;
//   * the pivot ROW p is needed by every lane only for its OWN columns -- a `switch` over the N row slots
//     reads (and restarts) it, no exchange at all.
// One uniform update  a_rc += z_r * y_c  (z_p = -1, y_c = -row_c / pivot, y_k = -1 / pivot) covers the four
// cases of the in-place algorithm, as in gj_tile_kernel.  I/O: a lane's columns are contiguous in the
// column-major input, so each lane moves them with one 1-D bulk copy into a padded slot; the permuted result
// is scattered into the (dense) staging area of the matrix and leaves with one bulk store per matrix.
// Column index k = CL * owner + j: the loop over `owner` is rolled, the CL step bodies are static.
// ==========================================================================================
template <typename T, int N, int CL, int WARPS>
struct GjcGeo {
    static constexpr int L = N / CL;                               // lanes per matrix
    static constexpr int MPW = 32 / L;                             // matrices per warp
    static constexpr int BLOCK = 32 * WARPS, MPB = MPW * WARPS;
    static constexpr int COLS_BYTES = CL * N * (int)sizeof(T);     // one lane's columns
    static constexpr int SLOT = COLS_BYTES + 16;                   // padded: consecutive lanes start one bank group further
    static constexpr int MAT_BYTES = N * N * (int)sizeof(T);
    static constexpr int EPC = 16 / (int)sizeof(T);
    static constexpr int PERM_BYTES = 2 * N * (int)sizeof(int);    // rowOfCol, colOfRow per matrix
    static constexpr size_t SMEM = (size_t)WARPS * 32 * SLOT + (size_t)MPB * PERM_BYTES + WARPS * 8 + 16;
    static_assert(L >= 1 && L <= 32 && (L & (L - 1)) == 0 && N <= 32, "lanes per matrix: a power of two; the row mask is 32 bits");
};

template <typename T, int N, int CL, int J>
__device__ __forceinline__ void gjc_step(T (&a)[N][CL], int lig, int gbase, int owner, unsigned &done, int &st, int *row_of_col, int *col_of_row) {
    const int k = CL * owner + J;
    // ---- pivot search: private to the owner lane
    T best = T(-1), bval = T(0);
    int brow = N;
    if (lig == owner) {
        #pragma unroll
        for (int r = 0; r < N; ++r) {
            const T av = dev_abs(a[r][J]);
            if (!((done >> r) & 1u) && av > best) { best = av; bval = a[r][J]; brow = r; }
        }
    }
    const int p = __shfl_sync(0xffffffffu, brow, gbase + owner);
    const T piv = __shfl_sync(0xffffffffu, bval, gbase + owner);
    if (st == 0 && !(dev_abs(piv) > T(0))) st = k + 1;
    // ---- multipliers: column k from the owner lane, z_p := -1
    T z[N];
    #pragma unroll
    for (int r = 0; r < N; ++r) {
        const T f = __shfl_sync(0xffffffffu, a[r][J], gbase + owner);
        z[r] = (r == p) ? T(-1) : f;
    }
    // ---- pivot row for this lane's own columns (run-time slot: switch), restarted; the owner restarts column k
    T y[CL];
    #pragma unroll
    for (int c = 0; c < CL; ++c) y[c] = T(0);
    #pragma unroll
    for (int r = 0; r < N; ++r) {
        if (r == p) {
            #pragma unroll
            for (int c = 0; c < CL; ++c) { y[c] = a[r][c]; a[r][c] = T(0); }
        }
    }
    if (lig == owner) {
        #pragma unroll
        for (int r = 0; r < N; ++r) a[r][J] = T(0);
        y[J] = T(1);
        if (p < N) { row_of_col[k] = p; col_of_row[p] = k; }
    }
    if (p < N) done |= 1u << p;
    // ---- a_rc += z_r * y_c,  y_c = -row_c / pivot
    const T nrp = T(-1) / piv;
    #pragma unroll
    for (int c = 0; c < CL; ++c) y[c] *= nrp;
    #pragma unroll
    for (int r = 0; r < N; ++r)
        #pragma unroll
        for (int c = 0; c < CL; ++c) a[r][c] = fma(z[r], y[c], a[r][c]);
}

template <typename T, int N, int CL, int J>
struct GjcSteps {
    static __device__ __forceinline__ void run(T (&a)[N][CL], int lig, int gbase, int owner, unsigned &done, int &st, int *roc, int *cor) {
        gjc_step<T, N, CL, J>(a, lig, gbase, owner, done, st, roc, cor);
        if constexpr (J + 1 < CL) GjcSteps<T, N, CL, J + 1>::run(a, lig, gbase, owner, done, st, roc, cor);
    }
};

template <typename T, int N, int CL, int WARPS, int MINB>
__global__ void __launch_bounds__((GjcGeo<T, N, CL, WARPS>::BLOCK), MINB)
gj_colsplit_kernel(const T *__restrict__ in, i64 in_stride, T *__restrict__ out, i64 out_stride, i64 batch, int *__restrict__ info) {
    using G = GjcGeo<T, N, CL, WARPS>;
    constexpr int L = G::L, EPC = G::EPC;
    extern __shared__ __align__(16) unsigned char smem_raw_gc[];
    const int warp = threadIdx.x >> 5, wl = threadIdx.x & 31;
    const int lig = wl % L, gbase = wl - lig, mat = wl / L;        // lane in group, first lane of the group, matrix of the warp
    unsigned char *wbase = smem_raw_gc + (size_t)warp * 32 * G::SLOT;
    unsigned char *slot = wbase + (size_t)wl * G::SLOT;            // this lane's input columns
    unsigned char *stage = wbase + (size_t)gbase * G::SLOT;        // dense N x N staging of the result (reuses the group's slots)
    int *row_of_col = reinterpret_cast<int *>(smem_raw_gc + (size_t)WARPS * 32 * G::SLOT) + (size_t)(warp * G::MPW + mat) * 2 * N;
    int *col_of_row = row_of_col + N;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(
        smem_raw_gc + (((size_t)WARPS * 32 * G::SLOT + (size_t)G::MPB * G::PERM_BYTES + 15) & ~(size_t)15)) + warp;
    if (wl == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const i64 ntiles = (batch + G::MPW - 1) / G::MPW;
    const i64 tstride = (i64)gridDim.x * WARPS;
    i64 tile = (i64)blockIdx.x * WARPS + warp;
    auto issue = [&](i64 t) {
        const i64 m = t * G::MPW + mat;
        const i64 left = batch - t * G::MPW;
        if (wl == 0) mbar_expect_tx(bar, (unsigned)((left < G::MPW ? left : G::MPW) * G::MAT_BYTES));
        __syncwarp();
        if (m < batch) bulk_load_1d(slot, in + m * in_stride + (i64)lig * CL * N, G::COLS_BYTES, bar);
    };
    if (tile < ntiles) issue(tile);
    unsigned phase = 0;
    #pragma unroll 1
    for (; tile < ntiles; tile += tstride) {
        const i64 m = tile * G::MPW + mat;
        const bool valid = m < batch;
        mbar_wait(bar, phase);
        phase ^= 1;

        T a[N][CL];                                                // a[r][c] = A(r, CL * lig + c)
        #pragma unroll
        for (int c = 0; c < CL; ++c)
            #pragma unroll
            for (int r = 0; r < N; r += EPC) {
                const T *p = reinterpret_cast<const T *>(slot + (c * N + r) * (int)sizeof(T));
                #pragma unroll
                for (int e = 0; e < EPC; ++e) a[r + e][c] = p[e];
            }
        unsigned done = 0;
        int st = 0;
        #pragma unroll 1
        for (int owner = 0; owner < L; ++owner) GjcSteps<T, N, CL, 0>::run(a, lig, gbase, owner, done, st, row_of_col, col_of_row);
        __syncwarp();                                              // permutation arrays visible; every lane is done with its slot

        // ---- scatter M[r][c] = Ainv[colOfRow[r]][rowOfCol[c]] into the dense staging area (NaN if flagged)
        T *sg = reinterpret_cast<T *>(stage);
        if (st) {
            #pragma unroll
            for (int c = 0; c < CL; ++c)
                #pragma unroll
                for (int r = 0; r < N; ++r) sg[(CL * lig + c) * N + r] = dev_nan<T>();
        } else {
            int ocol[CL];
            #pragma unroll
            for (int c = 0; c < CL; ++c) ocol[c] = row_of_col[CL * lig + c];
            #pragma unroll
            for (int r = 0; r < N; ++r) {
                const int orow = col_of_row[r];
                #pragma unroll
                for (int c = 0; c < CL; ++c) sg[ocol[c] * N + orow] = a[r][c];
            }
        }
        if (valid && lig == 0 && info) info[m] = st;
        fence_proxy_async();
        __syncwarp();
        if (valid && lig == 0) bulk_store_1d(out + m * out_stride, stage, G::MAT_BYTES);
        asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();                                              // every staging area has been read: slots reusable
        if (tile + tstride < ntiles) issue(tile + tstride);
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace invgpu

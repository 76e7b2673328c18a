// generic_smem.cuh -- any-n kernels: one warp (n <= 32) or one CTA (n <= 256) per matrix,
// the matrix resident in shared memory.  These are the "every n, both dtypes" tier of the
// engine: odd sizes, the 256 bucket and anything the register-tiled fast tiers
// (sweep_kernels.cuh, gj_kernels.cuh, gj_tile_kernels.cuh, tile_kernels.cuh) do not instantiate run here.
// Same math, same flags.  Where the working copy exceeds one CTA's shared memory (n towards 256) the same
// kernels run on a per-CTA slab of global memory (`gws`).
//
// What is computed (reference file:line each piece replaces):
//   SPD inverse     A = L L^T, M = L^-1, A^-1 = M^T M      src/inverse_cholesky_gpu.cu:251-354
//                   (4N+1 launches there, one launch here); upper triangle of the input is
//                   what gets read, like spotrf_("U") in src/inverse.c:92
//   SPD factor      L only                                  src/inverse_cholesky_gpu.cu:356-369
//   general inverse in-place Gauss-Jordan, partial pivoting src/gauss/batched_invert.cu:17-95
//                   (reference: 3N launches and zero-only pivoting)
//   GP mean / var   A^T (B+diag C)^-1 D, E - A^T (B+diag C)^-1 A
//                                                           src/gauss_bench.cu:127-265, 275-409
//                   add -> inv -> gemv -> dot there; here ONE kernel: C is added while B is
//                   loaded, the right-hand sides ride along as two extra rows of the
//                   Cholesky factor (row n = L^-1 A, row n+1 = L^-1 D), and the result is the
//                   in-kernel dot of those two rows.  No inverse is ever formed or written.
//
// info[] follows LAPACK: SPD paths report spotrf's info (k = leading minor of order k not
// positive definite), the general path sgetrf's (k = pivot k exactly zero).
#pragma once

#include "common.cuh"

namespace invgpu {

// ---------------------------------------------------------------------------------------
// Packed lower-triangular storage: L(i,j), j <= i < n, at S[i(i+1)/2 + j]; `extra` full rows
// of length n follow (the GP right-hand sides).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int row_offset(int i, int n) {
    return i < n ? packed_row(i) : packed_row(n) + (i - n) * n;
}

// Left-looking Cholesky over rows [0, nrows); rows >= n are the appended right-hand sides.
// One barrier per column: every thread recomputes the (broadcast-read) pivot itself.
template <typename T, int G>
__device__ int potrf_packed(T *S, int n, int nrows, int t) {
    int info = 0;
    for (int j = 0; j < n; ++j) {
        const T *rj = S + packed_row(j);
        T dsum = 0;
        for (int k = 0; k < j; ++k) { T v = rj[k]; dsum = fma(v, v, dsum); }
        const T d = rj[j] - dsum;
        if (!(d > T(0))) { info = j + 1; break; }   // uniform: every thread sees the same d
        const T s = dev_sqrt(d);
        const T inv = T(1) / s;
        for (int i = j + 1 + t; i < nrows; i += G) {
            T *ri = S + row_offset(i, n);
            T acc = 0;
            for (int k = 0; k < j; ++k) acc = fma(ri[k], rj[k], acc);
            ri[j] = (ri[j] - acc) * inv;
        }
        Group<G>::sync();
        if (t == 0) S[packed_row(j) + j] = s;       // nobody reads S(j,j) again before the next barrier
    }
    Group<G>::sync();
    return info;
}

// In-place inverse of the packed lower factor: thread t owns column t, rows advance in
// lock step so row i of L is still intact while it is being consumed.  Requires n <= G.
template <typename T, int G>
__device__ void trtri_packed(T *S, int n, int t) {
    for (int i = 0; i < n; ++i) {
        T *ri = S + packed_row(i);
        const T lii = ri[i];
        T acc = 0;
        if (t < i)
            for (int k = t; k < i; ++k) acc = fma(ri[k], S[packed_row(k) + t], acc);
        Group<G>::sync();
        if (t < i) ri[t] = -acc / lii;
        else if (t == i) ri[i] = T(1) / lii;
        Group<G>::sync();
    }
}

// In place  M -> M^T M  (lower triangle), row by row; rows > i are still M when row i is formed.
template <typename T, int G>
__device__ void lauum_packed(T *S, int n, int t) {
    for (int i = 0; i < n; ++i) {
        T acc = 0;
        if (t <= i)
            for (int k = i; k < n; ++k) {
                const T *rk = S + packed_row(k);
                acc = fma(rk[i], rk[t], acc);
            }
        Group<G>::sync();
        if (t <= i) S[packed_row(i) + t] = acc;
        Group<G>::sync();
    }
}

// ---------------------------------------------------------------------------------------
// group-wide reductions
// ---------------------------------------------------------------------------------------
template <typename T, int G>
__device__ __forceinline__ T group_sum(T v, T *scratch, int t) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (G == 32) return v;
    Group<G>::sync();
    if ((t & 31) == 0) scratch[t >> 5] = v;
    Group<G>::sync();
    T s = 0;
    #pragma unroll
    for (int w = 0; w < G / 32; ++w) s += scratch[w];
    return s;
}

// argmax of |v| with "smallest index wins ties" (what isamax / the oracle do).
template <typename T, int G>
__device__ __forceinline__ void group_argmax(T &best, int &idx, T *sval, int *sidx, int t) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
    }
    if (G == 32) return;
    Group<G>::sync();
    if ((t & 31) == 0) { sval[t >> 5] = best; sidx[t >> 5] = idx; }
    Group<G>::sync();
    best = sval[0]; idx = sidx[0];
    #pragma unroll
    for (int w = 1; w < G / 32; ++w) {
        const T ob = sval[w]; const int oi = sidx[w];
        if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
    }
}

// A flagged matrix gets a deterministic output: every element NaN (LAPACK leaves a half-finished
// factor there; a NaN cannot be mistaken for a result).
template <typename T, int G>
__device__ __forceinline__ void fill_nan(T *dst, int count, int t) {
    for (int idx = t; idx < count; idx += G) dst[idx] = dev_nan<T>();
}

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------
// STAGES bit mask: which phases of  potrf -> trtri -> lauum  run.  7 = full SPD inverse,
// 1 = factor only (reference decompose_cholesky_*_batched_device), 2 / 4 = the reference's
// staged `inverse_upper_stride` / `multiply_upper_stride` entry points.
//   bit 0 set  : input is an SPD matrix, its UPPER triangle is read (spotrf_("U") convention)
//   bit 0 clear: input is a lower-triangular factor (r >= c read)
//   bit 2 set  : output is the full symmetric matrix; otherwise lower triangle + zeroed upper
enum { SPD_POTRF = 1, SPD_TRTRI = 2, SPD_LAUUM = 4, SPD_INVERSE = 7 };

template <typename T, int G, typename IO, int STAGES>
__global__ void __launch_bounds__(G <= 32 ? 128 : G)
spd_generic_kernel(IO io, int n, i64 batch, int *__restrict__ info, T *gws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int groups = blockDim.x / G;
    const int g = threadIdx.x / G, t = threadIdx.x % G;
    // gws != null: the working copy does not fit shared memory (fp64, n > ~236) and lives in a per-CTA slab of
    // global memory instead (L2-resident: 148 x 2 slabs of 263 KB); same code, same barriers
    T *S = gws ? gws + ((size_t)blockIdx.x * groups + g) * packed_row(n)
               : reinterpret_cast<T *>(smem_raw) + (size_t)g * packed_row(n);
    const int nn = n * n;

    for (i64 m0 = (i64)blockIdx.x * groups; m0 < batch; m0 += (i64)gridDim.x * groups) {
        const i64 m = m0 + g;
        if (m >= batch) continue;                    // only possible in the warp tier
        const T *src = io.src(m);
        T *dst = io.dst(m);
        Group<G>::sync();                            // slab reuse
        for (int idx = t; idx < nn; idx += G) {
            const int c = idx / n, r = idx - c * n;
            if (STAGES & SPD_POTRF) { if (r <= c) S[packed_row(c) + r] = src[idx]; }   // upper -> L(c, r)
            else                    { if (r >= c) S[packed_row(r) + c] = src[idx]; }
        }
        Group<G>::sync();
        int st = 0;
        if (STAGES & SPD_POTRF) st = potrf_packed<T, G>(S, n, n, t);
        if (t == 0 && info) info[m] = st;
        if (st) { fill_nan<T, G>(dst, nn, t); continue; }
        if (STAGES & SPD_TRTRI) trtri_packed<T, G>(S, n, t);
        if (STAGES & SPD_LAUUM) lauum_packed<T, G>(S, n, t);
        for (int idx = t; idx < nn; idx += G) {      // coalesced store
            const int c = idx / n, r = idx - c * n;
            T v;
            if (r >= c) v = S[packed_row(r) + c];
            else v = (STAGES & SPD_LAUUM) ? S[packed_row(c) + r] : T(0);
            dst[idx] = v;
        }
    }
}

// In-place Gauss-Jordan with partial pivoting.  Smem holds the matrix row-major with an odd
// leading dimension (conflict-free for both "thread per row" and "thread per column").
template <typename T, int G, typename IO>
__global__ void __launch_bounds__(G <= 32 ? 128 : G)
gj_generic_kernel(IO io, int n, i64 batch, int *__restrict__ info, T *gws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int groups = blockDim.x / G;
    const int g = threadIdx.x / G, t = threadIdx.x % G;
    const int ld = n | 1;
    const size_t slab = (size_t)n * ld;
    // gws != null: working copy in a per-CTA slab of global memory (fp32 n > 240, fp64 n > 169), pivot log in smem
    T *S = gws ? gws + ((size_t)blockIdx.x * groups + g) * slab : reinterpret_cast<T *>(smem_raw) + (size_t)g * slab;
    // after the matrices: per-group pivot log, then reduction scratch (CTA tier only)
    int *piv = (gws ? reinterpret_cast<int *>(smem_raw)
                    : reinterpret_cast<int *>(reinterpret_cast<T *>(smem_raw) + (size_t)groups * slab)) + g * n;
    __shared__ T sval[8];
    __shared__ int sidx[8];
    const int nn = n * n;

    for (i64 m0 = (i64)blockIdx.x * groups; m0 < batch; m0 += (i64)gridDim.x * groups) {
        const i64 m = m0 + g;
        if (m >= batch) continue;
        const T *__restrict__ src = io.src(m);
        T *__restrict__ dst = io.dst(m);
        Group<G>::sync();
        for (int idx = t; idx < nn; idx += G) {
            const int c = idx / n, r = idx - c * n;
            S[r * ld + c] = src[idx];
        }
        Group<G>::sync();

        int st = 0;
        for (int k = 0; k < n; ++k) {
            T best = T(-1); int p = n;
            for (int i = k + t; i < n; i += G) {
                const T v = dev_abs(S[i * ld + k]);
                if (v > best) { best = v; p = i; }   // ascending i: first maximum is kept
            }
            group_argmax<T, G>(best, p, sval, sidx, t);
            if (!(best > T(0))) { st = k + 1; break; }        // uniform (NaN column counts as singular)
            const T pv = T(1) / S[p * ld + k];
            Group<G>::sync();
            for (int j = t; j < n; j += G) {                   // swap rows k <-> p, scale the new row k
                const T a = S[p * ld + j];
                const T b = S[k * ld + j];
                if (p != k) S[p * ld + j] = b;
                S[k * ld + j] = (j == k) ? pv : a * pv;
            }
            if (t == 0) piv[k] = p;
            Group<G>::sync();
            const T *rk = S + k * ld;
            for (int i = t; i < n; i += G) {                   // thread per row: eliminate column k
                if (i == k) continue;
                T *ri = S + i * ld;
                const T f = ri[k];
                for (int j = 0; j < n; ++j) ri[j] = fma(-f, rk[j], ri[j]);
                ri[k] = -f * pv;
            }
            Group<G>::sync();
        }
        if (t == 0 && info) info[m] = st;
        if (st) { fill_nan<T, G>(dst, nn, t); continue; }
        for (int i = t; i < n; i += G) {                       // undo interchanges as column swaps, own row only
            T *ri = S + i * ld;
            for (int k = n - 1; k >= 0; --k) {
                const int p = piv[k];
                if (p != k) { const T a = ri[k]; ri[k] = ri[p]; ri[p] = a; }
            }
        }
        Group<G>::sync();
        for (int idx = t; idx < nn; idx += G) {
            const int c = idx / n, r = idx - c * n;
            dst[idx] = S[r * ld + c];
        }
    }
}

// Fused GP mean / variance (see file header).
template <typename T, int G>
__global__ void __launch_bounds__(G <= 32 ? 128 : G)
gp_generic_kernel(GpIO<T> io, int n, i64 batch, int *__restrict__ info, T *gws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int groups = blockDim.x / G;
    const int g = threadIdx.x / G, t = threadIdx.x % G;
    const size_t slab = (size_t)packed_row(n) + 2 * (size_t)n;
    T *S = gws ? gws + ((size_t)blockIdx.x * groups + g) * slab : reinterpret_cast<T *>(smem_raw) + (size_t)g * slab;
    __shared__ T scratch[8];
    const int nn = n * n;

    for (i64 m0 = (i64)blockIdx.x * groups; m0 < batch; m0 += (i64)gridDim.x * groups) {
        const i64 m = m0 + g;
        if (m >= batch) continue;
        const T *__restrict__ B = io.b + m * nn;
        const T *__restrict__ a = io.a + m * n;
        const T *__restrict__ c = io.c + m * n;
        const T *__restrict__ d = io.d ? io.d + m * n : a;
        Group<G>::sync();
        for (int idx = t; idx < nn; idx += G) {
            const int cc = idx / n, r = idx - cc * n;
            if (r <= cc) S[packed_row(cc) + r] = B[idx] + (r == cc ? c[r] : T(0));   // + diag C on load
        }
        T *ra = S + packed_row(n), *rd = ra + n;
        for (int j = t; j < n; j += G) { ra[j] = a[j]; rd[j] = d[j]; }
        Group<G>::sync();
        const int st = potrf_packed<T, G>(S, n, n + 2, t);
        if (t == 0 && info) info[m] = st;
        if (st) {
            if (t == 0) {
                if (io.means) io.means[m] = dev_nan<T>();
                if (io.variances) io.variances[m] = dev_nan<T>();
            }
            continue;
        }
        T pm = 0, pq = 0;
        for (int j = t; j < n; j += G) { const T x = ra[j]; pm = fma(x, rd[j], pm); pq = fma(x, x, pq); }
        pm = group_sum<T, G>(pm, scratch, t);
        pq = group_sum<T, G>(pq, scratch, t);
        if (t == 0) {
            if (io.means) io.means[m] = pm;
            if (io.variances) io.variances[m] = io.e[m] - pq;
        }
    }
}

}  // namespace invgpu

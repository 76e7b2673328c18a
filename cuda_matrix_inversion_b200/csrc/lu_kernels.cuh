// lu_kernels.cuh -- partial-pivot LU with pivot / info output, inverse from the factors, and a multi-RHS
// solve: the `cublasSgetrfBatched` / `cublasSgetriBatched` pair of the reference's fastest GPU path
// (src/gauss/inverse_gpu.cu:24-50: LU left in devAs, PivotArray / infoArray :21-22) and the solve
// formulation its CPU side uses (`spotrs_`, src/gauss_cpu.c:87-144), for callers that want FACTORS rather
// than inverses.  LAPACK semantics throughout:
//   getrf  P A = L U in place (unit lower L below the diagonal, U on and above), ipiv 1-based ("row k was
//          interchanged with row ipiv[k]"), info = first k with U(k,k) exactly zero (the factorisation is
//          still completed, like sgetf2)
//   getri  A^-1 = U^-1 L^-1 P from the factors (strti2 + the sgetri column sweep + column interchanges);
//          info = k when U(k,k) is exactly zero (then the output is NaN-filled, nothing is inverted)
//   gesv   getrf on [A | B]: the row interchanges and the forward substitution with L ride along the
//          factorisation as nrhs extra columns; back substitution with U; A := LU, B := X
// One warp (n <= 32) or one CTA (n <= 256) per matrix, working copy row-major in shared memory with an odd
// leading dimension (or in a global scratch slab when it does not fit, `gws`), same skeleton as the any-n
// Gauss-Jordan kernel of generic_smem.cuh.
#pragma once

#include "generic_smem.cuh"

namespace invgpu {

// right-looking LU with partial pivoting on the n x ncols working copy S (ncols >= n: extra columns are
// right-hand sides).  Returns sgetrf's info; pivots (0-based) go to piv[0..n).
template <typename T, int G>
__device__ int lu_factor_rows(T *S, int ld, int n, int ncols, int *piv, T *sval, int *sidx, int t) {
    int info = 0;
    for (int k = 0; k < n; ++k) {
        T best = T(-1); int p = n;
        for (int i = k + t; i < n; i += G) {
            const T v = dev_abs(S[i * ld + k]);
            if (v > best) { best = v; p = i; }       // ascending i: the first maximum is kept (isamax)
        }
        group_argmax<T, G>(best, p, sval, sidx, t);
        if (!(best > T(0))) {                        // exactly zero (or NaN) column: sgetf2 records it and goes on
            if (!info) info = k + 1;
            if (t == 0) piv[k] = k;
            Group<G>::sync();
            continue;
        }
        if (t == 0) piv[k] = p;
        if (p != k)
            for (int j = t; j < ncols; j += G) { const T a = S[p * ld + j]; S[p * ld + j] = S[k * ld + j]; S[k * ld + j] = a; }
        Group<G>::sync();
        const T *rk = S + k * ld;
        const T s = T(1) / rk[k];
        for (int i = k + 1 + t; i < n; i += G) {     // thread per row: multiplier, then the row's trailing part
            T *ri = S + i * ld;
            const T f = ri[k] * s;
            ri[k] = f;
            for (int j = k + 1; j < ncols; ++j) ri[j] = fma(-f, rk[j], ri[j]);
        }
        Group<G>::sync();
    }
    return info;
}

// mode 0: getrf (A in place), mode 2: gesv (A in place, B -> X)
template <typename T, int G, typename IO>
__global__ void __launch_bounds__(G <= 32 ? 128 : G)
lu_factor_kernel(IO io, int n, i64 batch, int *__restrict__ pivots, int *__restrict__ info, T *B, int nrhs, T *gws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int groups = blockDim.x / G;
    const int g = threadIdx.x / G, t = threadIdx.x % G;
    const int ncols = n + nrhs;
    const int ld = ncols | 1;
    const size_t slab = (size_t)n * ld;
    T *S = gws ? gws + ((size_t)blockIdx.x * groups + g) * slab : reinterpret_cast<T *>(smem_raw) + (size_t)g * slab;
    int *piv = (gws ? reinterpret_cast<int *>(smem_raw)
                    : reinterpret_cast<int *>(reinterpret_cast<T *>(smem_raw) + (size_t)groups * slab)) + g * n;
    __shared__ T sval[8];
    __shared__ int sidx[8];
    const int nn = n * n;

    for (i64 m0 = (i64)blockIdx.x * groups; m0 < batch; m0 += (i64)gridDim.x * groups) {
        const i64 m = m0 + g;
        if (m >= batch) continue;
        const T *src = io.src(m);
        T *dst = io.dst(m);
        T *bm = nrhs ? B + m * (i64)n * nrhs : nullptr;
        Group<G>::sync();
        for (int idx = t; idx < nn; idx += G) { const int c = idx / n, r = idx - c * n; S[r * ld + c] = src[idx]; }
        for (int idx = t; idx < n * nrhs; idx += G) { const int c = idx / n, r = idx - c * n; S[r * ld + n + c] = bm[idx]; }
        Group<G>::sync();
        const int st = lu_factor_rows<T, G>(S, ld, n, ncols, piv, sval, sidx, t);
        if (t == 0 && info) info[m] = st;
        if (pivots) for (int k = t; k < n; k += G) pivots[m * n + k] = piv[k] + 1;
        for (int idx = t; idx < nn; idx += G) { const int c = idx / n, r = idx - c * n; dst[idx] = S[r * ld + c]; }
        if (nrhs) {
            if (st) {                                 // singular: LAPACK's gesv does not solve; X := NaN
                for (int idx = t; idx < n * nrhs; idx += G) bm[idx] = dev_nan<T>();
                continue;
            }
            // back substitution U X = Y: thread c owns right-hand side c for the division, thread i row i for the update
            for (int k = n - 1; k >= 0; --k) {
                const T ukk = S[k * ld + k];
                for (int c = t; c < nrhs; c += G) S[k * ld + n + c] /= ukk;
                Group<G>::sync();
                for (int i = t; i < k; i += G) {
                    const T u = S[i * ld + k];
                    for (int c = 0; c < nrhs; ++c) S[i * ld + n + c] = fma(-u, S[k * ld + n + c], S[i * ld + n + c]);
                }
                Group<G>::sync();
            }
            for (int idx = t; idx < n * nrhs; idx += G) { const int c = idx / n, r = idx - c * n; bm[idx] = S[r * ld + n + c]; }
        }
    }
}

// inverse from the factors (io.src = LU, io.dst = A^-1, may not alias unless the caller wants LU destroyed)
template <typename T, int G, typename IO>
__global__ void __launch_bounds__(G <= 32 ? 128 : G)
lu_invert_kernel(IO io, int n, i64 batch, const int *__restrict__ pivots, int *__restrict__ info, T *gws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int groups = blockDim.x / G;
    const int g = threadIdx.x / G, t = threadIdx.x % G;
    const int ld = n | 1;
    const size_t slab = (size_t)n * ld + n;                    // matrix + the work column of sgetri
    T *S = gws ? gws + ((size_t)blockIdx.x * groups + g) * slab : reinterpret_cast<T *>(smem_raw) + (size_t)g * slab;
    T *work = S + (size_t)n * ld;
    int *piv = (gws ? reinterpret_cast<int *>(smem_raw)
                    : reinterpret_cast<int *>(reinterpret_cast<T *>(smem_raw) + (size_t)groups * slab)) + g * n;
    const int nn = n * n;

    for (i64 m0 = (i64)blockIdx.x * groups; m0 < batch; m0 += (i64)gridDim.x * groups) {
        const i64 m = m0 + g;
        if (m >= batch) continue;
        const T *src = io.src(m);
        T *dst = io.dst(m);
        Group<G>::sync();
        for (int idx = t; idx < nn; idx += G) { const int c = idx / n, r = idx - c * n; S[r * ld + c] = src[idx]; }
        for (int k = t; k < n; k += G) piv[k] = pivots[m * n + k] - 1;
        Group<G>::sync();
        int st = 0;
        for (int k = 0; k < n; ++k) if (S[k * ld + k] == T(0)) { st = k + 1; break; }     // uniform (broadcast reads)
        if (t == 0 && info) info[m] = st;
        if (st) { fill_nan<T, G>(dst, nn, t); continue; }
        // U^-1 in place, column by column (strti2): thread r owns element (r, j)
        for (int j = 0; j < n; ++j) {
            T acc = 0;
            for (int r = t; r < j; r += G) {                    // G >= n is not required: strided rows, one partial each
                acc = 0;
                for (int k = r; k < j; ++k) acc = fma(S[r * ld + k], S[k * ld + j], acc);
                work[r] = acc;
            }
            Group<G>::sync();
            const T ujj = T(1) / S[j * ld + j];
            for (int r = t; r < j; r += G) S[r * ld + j] = -work[r] * ujj;
            if (t == 0) S[j * ld + j] = ujj;
            Group<G>::sync();
        }
        // X L = U^-1, columns right to left (sgetri): thread r owns row r
        for (int j = n - 2; j >= 0; --j) {
            for (int r = j + 1 + t; r < n; r += G) { work[r] = S[r * ld + j]; S[r * ld + j] = T(0); }
            Group<G>::sync();
            for (int r = t; r < n; r += G) {
                T *row = S + r * ld;
                T acc = row[j];
                for (int k = j + 1; k < n; ++k) acc = fma(-row[k], work[k], acc);
                row[j] = acc;
            }
            Group<G>::sync();
        }
        // undo the row interchanges as column interchanges, last to first; own row only
        for (int r = t; r < n; r += G) {
            T *row = S + r * ld;
            for (int j = n - 2; j >= 0; --j) {
                const int p = piv[j];
                if (p != j) { const T a = row[j]; row[j] = row[p]; row[p] = a; }
            }
        }
        Group<G>::sync();
        for (int idx = t; idx < nn; idx += G) { const int c = idx / n, r = idx - c * n; dst[idx] = S[r * ld + c]; }
    }
}

}  // namespace invgpu

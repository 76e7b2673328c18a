// gj_kernels.cuh -- register-resident Gauss-Jordan inverse with partial pivoting (general matrices).
//
// Replaces the reference's Gauss-Jordan kernel loop `invert` (src/gauss/batched_invert.cu:84-95:
// pivotRow / normalizeRow / transform_matrix, 3N launches, the whole batch streamed through DRAM
// every step) and its cuBLAS getrf/getriBatched path (src/gauss/inverse_gpu.cu:24-50) for n <= 64.
//
// Layout: "lane = row".  A matrix of (padded) order N is held by L = N / ROWS lanes of one warp; lane
// l keeps rows l, l + L, ... entirely in registers (ROWS x N values).  This makes partial pivoting
// natural on a GPU:
//   * the pivot search over column k is a shuffle arg-max over the lanes (column index k is a static
//     register index because the k loop is unrolled);
//   * rows are never swapped: the lane that wins step k simply becomes the pivot row ("implicit
//     pivoting"), publishes its scaled row through a double-buffered shared line (one __syncwarp per
//     step) and every other lane eliminates with one FFMA per element;
//   * the classic in-place trick keeps the inverse in the same registers (column k is replaced by the
//     k-th column of the transformed identity).
// With pi(k) = pivot row of step k, the registers finally hold  phys[r][c] = Ainv[pi^-1(r)][pi(c)],
// so lane r writes its values to row pi^-1(r) = "the step at which I was the pivot", column pi(c):
// for a fixed c the lanes of a matrix write one contiguous, fully used line -- coalesced without any
// staging.  Loads are coalesced the same way (column-major input, consecutive rows in consecutive
// lanes).  Arithmetic: the ROWS > 2 form scales the pivot row and subtracts like the oracle; the lean form (ROWS <= 2,
// fp32) uses rcp.approx.f32 and the multiplier form z = -a * r with two in-place FMAs per element, which differs from
// the oracle in rounding only (bounded against the fp64 truth by tests/util.py::assert_general_parity and the
// ill-conditioned test in tests/test_dropin_gpu.py).  Non-finite INPUTS give undefined output: the select-free second
// FMA (a += 0 * row) turns an Inf / NaN of a pivot row into NaN in every row, without touching info.
// Since round 2 this kernel is built only by `make lab=1`; gj_roll_kernels.cuh is the default for these sizes.
//
// Runtime order n <= N: the matrix is embedded as blockdiag(A, I); padded rows can only win padded
// columns, so pivots, flags and results of A are unaffected.
// info: k (1-based) if no non-zero pivot exists for column k (sgetrf's "U(k,k) is exactly zero"; a NaN
// column counts as singular).  Flagged outputs are NaN.
#pragma once

#include "common.cuh"

namespace invgpu {

template <typename T, int N, int ROWS>
struct GjGeo {
    static constexpr int L = N / ROWS;                  // lanes per matrix
    static constexpr int MPW = 32 / L;                  // matrices per warp
    static constexpr int WARPS = 4;
    static constexpr int BLOCK = 32 * WARPS;
    static constexpr int MPB = MPW * WARPS;             // matrices per CTA
    // per matrix: 2 pivot-row lines of N values + N pivot indices (as T-sized words), 16-byte aligned
    static constexpr int WORDS = 2 * N + ((N * (int)sizeof(int) + (int)sizeof(T) - 1) / (int)sizeof(T) + 3) / 4 * 4 + (L < 32 ? 4 : 0);
    static_assert(L >= 1 && L <= 32 && (L & (L - 1)) == 0, "lanes per matrix must be a power of two <= 32");
};

template <typename T, int N, int ROWS, typename IO, int MINB>
__global__ void __launch_bounds__((GjGeo<T, N, ROWS>::BLOCK), MINB)
gj_rowlane_kernel(IO io, int n, i64 batch, int *__restrict__ info) {
    using G = GjGeo<T, N, ROWS>;
    constexpr int L = G::L;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);

    const int lane = threadIdx.x & 31;
    const int l = lane % L;                               // lane inside the matrix' group
    const int grp = (threadIdx.x >> 5) * G::MPW + lane / L;
    T *line = smem + (size_t)grp * G::WORDS;              // two pivot-row lines
    int *piv = reinterpret_cast<int *>(line + 2 * N);     // pi(k)

    #pragma unroll 1
    for (i64 base = (i64)blockIdx.x * G::MPB; base < batch; base += (i64)gridDim.x * G::MPB) {
        const i64 m = base + grp;
        const bool valid = m < batch;
        const T *__restrict__ src = io.src(valid ? m : batch - 1);

        // a[q][c] = A(row l + L q, column c); identity padding outside n
        T a[ROWS][N];
        if (n == N) {                                              // exact order: compile-time offsets, no padding predicates
            #pragma unroll
            for (int q = 0; q < ROWS; ++q) {
                #pragma unroll
                for (int c = 0; c < N; ++c) a[q][c] = __ldcs(src + c * N + l + L * q);
            }
        } else {
            #pragma unroll
            for (int q = 0; q < ROWS; ++q) {
                const int row = l + L * q;
                #pragma unroll
                for (int c = 0; c < N; ++c) {
                    T v = (row == c) ? T(1) : T(0);
                    if (row < n && c < n) v = __ldcs(src + (size_t)c * n + row);
                    a[q][c] = v;
                }
            }
        }

        int st = 0;
        int mystep[ROWS];
        bool pivoted[ROWS];
        #pragma unroll
        for (int q = 0; q < ROWS; ++q) { mystep[q] = 0; pivoted[q] = false; }

        if constexpr (ROWS <= 2) {
            // ---- lean form (one or two rows per lane): ~70-95 issue slots per matrix and pivot instead of ~220.
            //  * fp32 pivot search: |a| as an unsigned key -> one REDUX max + one ballot per row slot (first maximum
            //    wins ties, like isamax);
            //  * the pivot lane publishes its RAW row; the uniform update  a_ic += z_i * row_c  with  z_i = -a_ik / piv
            //    needs no select per element (z_p = -1 makes the pivot row's registers exact zeros; a second in-place
            //    FMA with the multiplier 1/piv in the pivot row and 0 elsewhere turns it into row / piv), z is the new column k;
            //  * fp32 updates are packed (FFMA2: two columns per issue slot);
            //  * two rows per lane halve the shared-memory wavefronts per matrix (a broadcast LDS costs 4 bytes per
            //    lane and clock whatever the address pattern: with one row per lane that is one word per FMA).
            const unsigned gmask = (L == 32) ? 0xffffffffu : (((1u << (L & 31)) - 1u) << (lane - l));
            #pragma unroll
            for (int k = 0; k < N; ++k) {
                T *pr_line = line + (k & 1) * N;
                int pl, pq = 0;                                    // absolute lane and row slot of the pivot row
                bool none;
                if constexpr (sizeof(T) == 4) {
                    unsigned key[ROWS], mykey = 0u;
                    #pragma unroll
                    for (int q = 0; q < ROWS; ++q) {
                        const float v = fabsf((float)a[q][k]);
                        key[q] = (!pivoted[q] && v == v) ? __float_as_uint(v) : 0u;
                        mykey = max(mykey, key[q]);
                    }
                    const unsigned mx = __reduce_max_sync(gmask, mykey);
                    unsigned cand = __ballot_sync(0xffffffffu, !pivoted[0] && key[0] == mx) & gmask;
                    #pragma unroll
                    for (int q = 1; q < ROWS; ++q) {
                        const unsigned cq = __ballot_sync(0xffffffffu, !pivoted[q] && key[q] == mx) & gmask;
                        if (cand == 0u) { cand = cq; pq = q; }
                    }
                    pl = __ffs((int)cand) - 1;
                    none = mx == 0u;
                } else {
                    T best = T(-1);
                    int prow = N;
                    #pragma unroll
                    for (int q = 0; q < ROWS; ++q) {
                        const T v = dev_abs(a[q][k]);
                        if (!pivoted[q] && v > best) { best = v; prow = l + L * q; }     // NaN never wins
                    }
                    #pragma unroll
                    for (int o = L / 2; o > 0; o >>= 1) {
                        const T ob = __shfl_xor_sync(0xffffffffu, best, o);
                        const int orow = __shfl_xor_sync(0xffffffffu, prow, o);
                        if (ob > best || (ob == best && orow < prow)) { best = ob; prow = orow; }
                    }
                    none = !(best > T(0));
                    // only NaNs left: take the first row that has not been a pivot (warp-wide ballots: unconditional)
                    unsigned cand = __ballot_sync(0xffffffffu, !pivoted[0]) & gmask;
                    int cq_slot = 0;
                    #pragma unroll
                    for (int q = 1; q < ROWS; ++q) {
                        const unsigned cq = __ballot_sync(0xffffffffu, !pivoted[q]) & gmask;
                        if (cand == 0u) { cand = cq; cq_slot = q; }
                    }
                    if (prow >= N) prow = (__ffs((int)cand) - 1 - (lane - l)) + L * cq_slot;
                    pl = (lane - l) + prow % L;
                    pq = (prow / L) % ROWS;
                }
                if (st == 0 && none) st = k + 1;                   // uniform inside the group
                T mine_k = a[0][k];
                #pragma unroll
                for (int q = 1; q < ROWS; ++q) mine_k = (pq == q) ? a[q][k] : mine_k;
                const T pivv = __shfl_sync(0xffffffffu, mine_k, pl);
                T r;
                if constexpr (sizeof(T) == 4) asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(pivv));   // 1 ulp, no slow path
                else r = T(1) / pivv;
                const bool isl = lane == pl;
                bool isp[ROWS];
                T z[ROWS];
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) {
                    isp[q] = isl && pq == q;
                    z[q] = isp[q] ? T(-1) : -a[q][k] * r;          // z_p = -1: a - a = 0 exactly, then r * row
                }
                T w[ROWS];                                         // second multiplier: r for the pivot row, 0 elsewhere
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) w[q] = isp[q] ? r : T(0);
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) {
                    if (isp[q]) {                                  // one branch per row slot: static register names, no selects
                        #pragma unroll
                        for (int c4 = 0; c4 < N; c4 += 4) {
                            if constexpr (sizeof(T) == 4) {
                                // 64-bit stores straight from the FFMA2 register pairs (a merged 128-bit store costs four staging moves)
                                const unsigned sa = (unsigned)__cvta_generic_to_shared(pr_line + c4);
                                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sa), "f"((float)a[q][c4]), "f"((float)a[q][c4 + 1]) : "memory");
                                asm volatile("st.shared.v2.f32 [%0+8], {%1, %2};" ::"r"(sa), "f"((float)a[q][c4 + 2]), "f"((float)a[q][c4 + 3]) : "memory");
                            } else {
                                *reinterpret_cast<double2 *>(pr_line + c4) = make_double2((double)a[q][c4], (double)a[q][c4 + 1]);
                                *reinterpret_cast<double2 *>(pr_line + c4 + 2) = make_double2((double)a[q][c4 + 2], (double)a[q][c4 + 3]);
                            }
                        }
                        piv[k] = l + L * q;
                        pivoted[q] = true;
                        mystep[q] = k;
                    }
                }
                __syncwarp();
                // a += z * row (exact zero in the pivot row), then a += w * row (row / piv in the pivot row, a + 0 elsewhere):
                // both in place and select-free
                #pragma unroll
                for (int c4 = 0; c4 < N; c4 += 4) {
                    if constexpr (sizeof(T) == 4) {
                        const float4 v = *reinterpret_cast<const float4 *>(pr_line + c4);
                        #pragma unroll
                        for (int q = 0; q < ROWS; ++q) {
                            const float2 z2 = make_float2((float)z[q], (float)z[q]), w2 = make_float2((float)w[q], (float)w[q]);
                            float2 lo = make_float2((float)a[q][c4], (float)a[q][c4 + 1]), hi = make_float2((float)a[q][c4 + 2], (float)a[q][c4 + 3]);
                            lo = __ffma2_rn(z2, make_float2(v.x, v.y), lo);
                            hi = __ffma2_rn(z2, make_float2(v.z, v.w), hi);
                            lo = __ffma2_rn(w2, make_float2(v.x, v.y), lo);
                            hi = __ffma2_rn(w2, make_float2(v.z, v.w), hi);
                            a[q][c4] = lo.x; a[q][c4 + 1] = lo.y; a[q][c4 + 2] = hi.x; a[q][c4 + 3] = hi.y;
                        }
                    } else {
                        const double2 u = *reinterpret_cast<const double2 *>(pr_line + c4);
                        const double2 v = *reinterpret_cast<const double2 *>(pr_line + c4 + 2);
                        #pragma unroll
                        for (int q = 0; q < ROWS; ++q) {
                            if (isp[q]) {
                                a[q][c4] = r * (T)u.x; a[q][c4 + 1] = r * (T)u.y; a[q][c4 + 2] = r * (T)v.x; a[q][c4 + 3] = r * (T)v.y;
                            } else {
                                a[q][c4] = fma(z[q], (T)u.x, a[q][c4]);
                                a[q][c4 + 1] = fma(z[q], (T)u.y, a[q][c4 + 1]);
                                a[q][c4 + 2] = fma(z[q], (T)v.x, a[q][c4 + 2]);
                                a[q][c4 + 3] = fma(z[q], (T)v.y, a[q][c4 + 3]);
                            }
                        }
                    }
                }
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) a[q][k] = isp[q] ? r : z[q];
            }
        } else {
        #pragma unroll
        for (int k = 0; k < N; ++k) {
            T *pr_line = line + (k & 1) * N;
            // ---- pivot search: arg-max of |a[.][k]| over the rows that have not been pivots yet;
            //      the smallest row index wins ties (first maximum, like isamax)
            T best = T(-1);
            int prow = N;
            #pragma unroll
            for (int q = 0; q < ROWS; ++q) {
                const T v = dev_abs(a[q][k]);
                if (!pivoted[q] && v > best) { best = v; prow = l + L * q; }
            }
            #pragma unroll
            for (int o = L / 2; o > 0; o >>= 1) {
                const T ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int orow = __shfl_xor_sync(0xffffffffu, prow, o);
                if (ob > best || (ob == best && orow < prow)) { best = ob; prow = orow; }
            }
            if (st == 0 && !(best > T(0))) st = k + 1;             // uniform inside the group
            const int pl = prow % L;                               // pivot lane (garbage-safe: prow <= N)
            const int pq = (prow / L) % ROWS;                      // its row slot
            // ---- the pivot lane scales its row and publishes it
            T mine_k = a[0][k];
            #pragma unroll
            for (int q = 1; q < ROWS; ++q) mine_k = (pq == q) ? a[q][k] : mine_k;
            const T pv = T(1) / __shfl_sync(0xffffffffu, mine_k, (lane - l) + pl);
            if (l == pl) {
                #pragma unroll
                for (int c4 = 0; c4 < N; c4 += 4) {
                    T t[4];
                    #pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const int c = c4 + w;
                        T x = a[0][c];
                        #pragma unroll
                        for (int q = 1; q < ROWS; ++q) x = (pq == q) ? a[q][c] : x;
                        t[w] = (c == k) ? pv : x * pv;
                    }
                    if (sizeof(T) == 4) {
                        *reinterpret_cast<float4 *>(pr_line + c4) = make_float4((float)t[0], (float)t[1], (float)t[2], (float)t[3]);
                    } else {
                        *reinterpret_cast<double2 *>(pr_line + c4) = make_double2((double)t[0], (double)t[1]);
                        *reinterpret_cast<double2 *>(pr_line + c4 + 2) = make_double2((double)t[2], (double)t[3]);
                    }
                }
                piv[k] = prow;
            }
            __syncwarp();
            // ---- everybody eliminates column k (the pivot row itself just takes the scaled row)
            bool isp[ROWS];
            T f[ROWS];
            #pragma unroll
            for (int q = 0; q < ROWS; ++q) {
                isp[q] = (l + L * q == prow);
                f[q] = isp[q] ? T(0) : a[q][k];
                if (isp[q]) { pivoted[q] = true; mystep[q] = k; }
            }
            #pragma unroll
            for (int c4 = 0; c4 < N; c4 += 4) {
                T p[4];
                if (sizeof(T) == 4) {
                    const float4 v = *reinterpret_cast<const float4 *>(pr_line + c4);
                    p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
                } else {
                    const double2 u = *reinterpret_cast<const double2 *>(pr_line + c4);
                    const double2 v = *reinterpret_cast<const double2 *>(pr_line + c4 + 2);
                    p[0] = u.x; p[1] = u.y; p[2] = v.x; p[3] = v.y;
                }
                #pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int c = c4 + w;
                    #pragma unroll
                    for (int q = 0; q < ROWS; ++q) {
                        if (c == k) a[q][c] = isp[q] ? pv : -f[q] * pv;
                        else a[q][c] = fma(-f[q], p[w], isp[q] ? p[w] : a[q][c]);
                    }
                }
            }
        }
        }
        __syncwarp();

        if (valid) {
            if (l == 0 && info) info[m] = (st > n) ? 0 : st;       // a "singular" padded column cannot happen; guard anyway
            T *__restrict__ dst = io.dst(m);
            const bool bad = st != 0 && st <= n;
            if (n == N && !bad) {                                  // the common case: 32-bit offsets, no bounds predicates
                #pragma unroll
                for (int c = 0; c < N; ++c) {
                    const int ocol = piv[c] * N;
                    #pragma unroll
                    for (int q = 0; q < ROWS; ++q) __stcs(dst + ocol + mystep[q], a[q][c]);
                }
            } else
            #pragma unroll
            for (int c = 0; c < N; ++c) {
                const int ocol = piv[c];                           // pi(c): broadcast read
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) {
                    const int orow = mystep[q];                    // pi^-1(my row)
                    if (bad) { if (l + L * q < n && c < n) dst[(size_t)c * n + l + L * q] = dev_nan<T>(); }
                    else if (orow < n && ocol < n) __stcs(dst + (size_t)ocol * n + orow, a[q][c]);
                }
            }
        }
        __syncwarp();                                              // piv / lines are reused by the next matrix
    }
}

}  // namespace invgpu

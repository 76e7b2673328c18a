// gj_roll_kernels.cuh -- register-resident Gauss-Jordan inverse with partial pivoting, lane = row, ROLLED pivot loop.
//
// Same job and same layout as gj_rowlane_kernel (gj_kernels.cuh: replaces the reference's `invert` launch loop,
// src/gauss/batched_invert.cu:84-95, and its cuBLAS getrf/getriBatched path, src/gauss/inverse_gpu.cu:24-50): a matrix
// of (padded) order N is held by L = N / ROWS lanes of one warp, lane l keeps rows l, l + L entirely in registers, rows
// are never swapped (implicit pivoting), the permutation is undone by the final store.  What is new:
//
//  * ROTATING REGISTER WINDOW.  The unrolled kernel needs the pivot column as a static register index, i.e. N copies
//    of the step body: 44 KB of code at n = 32, ~150 KB at n = 64, and ncu showed instruction fetch as the top stall
//    (I-cache hit rate 60-78 %, profiles/r1_gj_*_summary.md).  Here the pivot column is ALWAYS register 0: every step
//    writes its results one column down (the FMA's destination simply is the neighbouring register) and appends the
//    new inverse column at the end of the window, so after N steps the window is back in natural order.  Two pivots
//    per loop iteration keep the FFMA2 register pairs aligned (the window moves by one PAIR per iteration); the loop
//    itself stays rolled: two step bodies of code for any N.
//  * DEFERRED ROW SCALING.  The pivot row is published RAW and never scaled inside the loop; every other row does ONE
//    fused multiply-add per element,  a_ic += z_i * row_c  with  z_i = -a_ik / pivot  (z = 0 in the pivot row: exact),
//    and the new column k is  z_i  (1 in the pivot row).  In the augmented picture [A | I] this is Gauss-Jordan without
//    normalisation: A ends as a (permuted) diagonal of the pivots, so each row is multiplied ONCE at the end by the
//    reciprocal of its own pivot.  The lean unrolled kernel spent a second FFMA2 pass per step on that scaling.
//  * fp32 pivot search: |a| as an unsigned key -> one REDUX max + one ballot per row slot (first maximum wins ties, like
//    isamax); fp64: shuffle arg-max.
//
// Arithmetic differs from the oracle's scale-then-subtract form only in rounding (rcp.approx + multiplier form in fp32);
// the parity tests bound it against the fp64 truth.  Non-finite inputs give undefined output (0 * Inf in the pivot row).
// info: k (1-based) if no non-zero pivot exists for column k (sgetrf's "U(k,k) is exactly zero"; a NaN column counts as
// singular).  Flagged outputs are NaN.  Runtime order n <= N: the matrix is embedded as blockdiag(A, I).
#pragma once

#include "common.cuh"

namespace invgpu {

template <typename T, int N, int ROWS>
struct GjRollGeo {
    static constexpr int L = N / ROWS;                  // lanes per matrix
    static constexpr int MPW = 32 / L;                  // matrices per warp
    static constexpr int WARPS = 4;
    static constexpr int BLOCK = 32 * WARPS;
    static constexpr int MPB = MPW * WARPS;             // matrices per CTA
    static constexpr int LINE = N + 4;                  // a published pivot row: N window entries + the pending column
    // per matrix: 2 lines + N pivot indices (as T-sized words), 16-byte aligned
    static constexpr int WORDS = 2 * LINE + ((N * (int)sizeof(int) + (int)sizeof(T) - 1) / (int)sizeof(T) + 3) / 4 * 4;
    static_assert(L >= 1 && L <= 32 && (L & (L - 1)) == 0 && N % 2 == 0, "lanes per matrix must be a power of two <= 32");
};

#ifndef GJR_TWO_REDUX
#define GJR_TWO_REDUX 1
#endif

template <typename T> struct GjPair { T x, y; };

template <typename T> __device__ __forceinline__ GjPair<T> pair_fma(T z, GjPair<T> r, GjPair<T> a) {
    GjPair<T> o; o.x = fma(z, r.x, a.x); o.y = fma(z, r.y, a.y); return o;
}
template <> __device__ __forceinline__ GjPair<float> pair_fma<float>(float z, GjPair<float> r, GjPair<float> a) {
    const float2 o = __ffma2_rn(make_float2(z, z), make_float2(r.x, r.y), make_float2(a.x, a.y));
    GjPair<float> p; p.x = o.x; p.y = o.y; return p;
}
// two neighbouring pairs of a published row with one 128-bit (fp32) / two 128-bit (fp64) broadcast loads; p is 16-byte aligned
template <typename T> __device__ __forceinline__ void load_two_pairs(const T *p, GjPair<T> &r0, GjPair<T> &r1) {
    r0 = *reinterpret_cast<const GjPair<T> *>(p);
    r1 = *reinterpret_cast<const GjPair<T> *>(p + 2);
}
template <> __device__ __forceinline__ void load_two_pairs<float>(const float *p, GjPair<float> &r0, GjPair<float> &r1) {
    const float4 v = *reinterpret_cast<const float4 *>(p);
    r0.x = v.x; r0.y = v.y; r1.x = v.z; r1.y = v.w;
}
// publish one register pair of the pivot row: 64-bit stores straight from the FFMA2 register pairs (a merged 128-bit store
// costs four staging moves per store: the first version of this kernel spent 19 % of its issue slots on them; a variant with
// four pivots per iteration and quad-aligned 128-bit publish stores still needed the moves and measured 7 % slower)
template <typename T> __device__ __forceinline__ void store_pair(T *p, GjPair<T> v) { *reinterpret_cast<GjPair<T> *>(p) = v; }
template <> __device__ __forceinline__ void store_pair<float>(float *p, GjPair<float> v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "f"(v.x), "f"(v.y) : "memory");
}
template <typename T> __device__ __forceinline__ T fast_rcp(T x) { return T(1) / x; }
template <> __device__ __forceinline__ float fast_rcp<float>(float x) { float r; asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Pivot search over the column values v[q] of the rows that have not been pivots: absolute lane `pl` and row slot `pq` of
// the first maximum of |v|, `none` when no non-zero (non-NaN) candidate exists.  Uniform inside the matrix' lane group.
template <typename T, int ROWS, int L>
__device__ __forceinline__ void pivot_search(const T (&v)[ROWS], const bool (&pivoted)[ROWS], unsigned gmask, int lane, int l,
                                             int &pl, int &pq, bool &none, T &pv) {
    pq = 0;
#if GJR_TWO_REDUX
    if constexpr (sizeof(T) == 4 && L == 32) {
        // (full-warp groups only: measured on B200, n = 32 / 64: +3.6 % / +2.4 %; with four 8-lane groups per warp, n = 16, the
        // partial-mask reductions cost more than the ballots they replace: 0.357 -> 0.304 of the roofline, so those keep them)
        // fp32: max of |a| as an unsigned key, then the smallest candidate row together with the sign of its value,
        // (row << 1 | sign) -- the row dominates the order, so this is the first maximum in row order (isamax / the oracle) and
        // the pivot is sign | max: two REDUX, no ballots, no shuffle of the pivot value
        unsigned key[ROWS], mykey = 0u;
        #pragma unroll
        for (int q = 0; q < ROWS; ++q) {
            const float av = fabsf((float)v[q]);
            key[q] = (!pivoted[q] && av == av) ? __float_as_uint(av) : 0u;
            mykey = max(mykey, key[q]);
        }
        const unsigned mx = __reduce_max_sync(gmask, mykey);
        unsigned cand = 0xffffffffu;
        #pragma unroll
        for (int q = 0; q < ROWS; ++q) {
            const unsigned code = ((unsigned)(l + L * q) << 1) | (__float_as_uint((float)v[q]) >> 31);
            cand = (!pivoted[q] && key[q] == mx) ? min(cand, code) : cand;
        }
        const unsigned best = __reduce_min_sync(gmask, cand);
        const int prow = (int)(best >> 1);
        pl = (lane - l) + prow % L;
        pq = prow / L;
        none = mx == 0u;
        pv = (T)__uint_as_float(mx | (best << 31));
        return;
    }
#endif
    if constexpr (sizeof(T) == 4) {
        unsigned key[ROWS], mykey = 0u;
        #pragma unroll
        for (int q = 0; q < ROWS; ++q) {
            const float av = fabsf((float)v[q]);
            key[q] = (!pivoted[q] && av == av) ? __float_as_uint(av) : 0u;
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
        int prow = L * ROWS;
        #pragma unroll
        for (int q = 0; q < ROWS; ++q) {
            const T av = dev_abs(v[q]);
            if (!pivoted[q] && av > best) { best = av; prow = l + L * q; }      // NaN never wins
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
        if (prow >= L * ROWS) prow = (__ffs((int)cand) - 1 - (lane - l)) + L * cq_slot;
        pl = (lane - l) + prow % L;
        pq = (prow / L) % ROWS;
    }
    T mine = v[0];
    #pragma unroll
    for (int q = 1; q < ROWS; ++q) mine = (pq == q) ? v[q] : mine;
    pv = __shfl_sync(0xffffffffu, mine, pl);
}

// EXACT: the runtime order is N (compile-time offsets, no padding predicates); otherwise n <= N, embedded as blockdiag(A, I)
template <typename T, int N, int ROWS, typename IO, int MINB, bool EXACT>
__global__ void __launch_bounds__((GjRollGeo<T, N, ROWS>::BLOCK), MINB)
gj_roll_kernel(IO io, int n_runtime, i64 batch, int *__restrict__ info) {
    const int n = EXACT ? N : n_runtime;
    using G = GjRollGeo<T, N, ROWS>;
    constexpr int L = G::L, H = N / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);

    const int lane = threadIdx.x & 31;
    const int l = lane % L;                               // lane inside the matrix' group
    const int grp = (threadIdx.x >> 5) * G::MPW + lane / L;
    T *line = smem + (size_t)grp * G::WORDS;              // two pivot-row lines
    int *piv = reinterpret_cast<int *>(line + 2 * G::LINE);   // pi(k)
    const unsigned gmask = (L == 32) ? 0xffffffffu : (((1u << (L & 31)) - 1u) << (lane - l));

    #pragma unroll 1
    for (i64 base = (i64)blockIdx.x * G::MPB; base < batch; base += (i64)gridDim.x * G::MPB) {
        const i64 m = base + grp;
        const bool valid = m < batch;
        const T *__restrict__ src = io.src(valid ? m : batch - 1);

        // ap[q][i] = A(row l + L q, columns 2i, 2i + 1); identity padding outside n
        GjPair<T> ap[ROWS][H];
        if constexpr (EXACT) {
            #pragma unroll
            for (int q = 0; q < ROWS; ++q) {
                #pragma unroll
                for (int i = 0; i < H; ++i) {
                    ap[q][i].x = __ldcs(src + (2 * i) * N + l + L * q);
                    ap[q][i].y = __ldcs(src + (2 * i + 1) * N + l + L * q);
                }
            }
        } else {
            #pragma unroll
            for (int q = 0; q < ROWS; ++q) {
                const int row = l + L * q;
                #pragma unroll
                for (int c = 0; c < N; ++c) {
                    T v = (row == c) ? T(1) : T(0);
                    if (row < n && c < n) v = __ldcs(src + (size_t)c * n + row);
                    if (c & 1) ap[q][c >> 1].y = v; else ap[q][c >> 1].x = v;
                }
            }
        }

        int st = 0;
        int mystep[ROWS];
        bool pivoted[ROWS];
        T rscale[ROWS];
        #pragma unroll
        for (int q = 0; q < ROWS; ++q) { mystep[q] = 0; pivoted[q] = false; rscale[q] = T(1); }

        #pragma unroll 1
        for (int kk = 0; kk < H; ++kk) {
            T ca[ROWS];                                            // the new column 2kk (pending: it joins the window after step B)
            // ---------------- step A: pivot column = register pair 0, .x
            {
                T *pr = line;
                T v[ROWS];
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) v[q] = ap[q][0].x;
                int pl, pq; bool none;
                T pv;
                pivot_search<T, ROWS, L>(v, pivoted, gmask, lane, l, pl, pq, none, pv);
                if (st == 0 && none) st = 2 * kk + 1;              // uniform inside the group
                const T r = fast_rcp<T>(pv);
                T z[ROWS];
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) {
                    const bool isp = (lane == pl) && (pq == q);
                    z[q] = isp ? T(0) : -v[q] * r;
                    ca[q] = isp ? T(1) : z[q];
                    if (isp) {                                     // one branch per row slot: static register names
                        #pragma unroll
                        for (int i = 0; i < H; ++i) store_pair<T>(pr + 2 * i, ap[q][i]);
                        piv[2 * kk] = l + L * q;
                        pivoted[q] = true; mystep[q] = 2 * kk; rscale[q] = r;
                    }
                }
                __syncwarp();
                #pragma unroll
                for (int i2 = 0; i2 < H; i2 += 2) {                // two pairs per 128-bit (fp32) broadcast load
                    GjPair<T> r0, r1;
                    load_two_pairs<T>(pr + 2 * i2, r0, r1);
                    #pragma unroll
                    for (int q = 0; q < ROWS; ++q) {
                        ap[q][i2] = pair_fma<T>(z[q], r0, ap[q][i2]);
                        ap[q][i2 + 1] = pair_fma<T>(z[q], r1, ap[q][i2 + 1]);
                    }
                }
            }
            // ---------------- step B: pivot column = register pair 0, .y; the window moves on by one pair
            {
                T *pr = line + G::LINE;
                T v[ROWS];
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) v[q] = ap[q][0].y;
                int pl, pq; bool none;
                T pv;
                pivot_search<T, ROWS, L>(v, pivoted, gmask, lane, l, pl, pq, none, pv);
                if (st == 0 && none) st = 2 * kk + 2;
                const T r = fast_rcp<T>(pv);
                T z[ROWS], cb[ROWS];
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) {
                    const bool isp = (lane == pl) && (pq == q);
                    z[q] = isp ? T(0) : -v[q] * r;
                    cb[q] = isp ? T(1) : z[q];
                    if (isp) {
                        #pragma unroll
                        for (int i = 1; i < H; ++i) store_pair<T>(pr + 2 * i, ap[q][i]);
                        pr[N] = ca[q];                             // the pending column of the pivot row
                        piv[2 * kk + 1] = l + L * q;
                        pivoted[q] = true; mystep[q] = 2 * kk + 1; rscale[q] = r;
                    }
                }
                __syncwarp();
                const T rca = pr[N];
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) ca[q] = fma(z[q], rca, ca[q]);
                {
                    const GjPair<T> rr = *reinterpret_cast<const GjPair<T> *>(pr + 2);
                    #pragma unroll
                    for (int q = 0; q < ROWS; ++q) ap[q][0] = pair_fma<T>(z[q], rr, ap[q][1]);
                }
                #pragma unroll
                for (int i2 = 2; i2 < H; i2 += 2) {
                    GjPair<T> r0, r1;
                    load_two_pairs<T>(pr + 2 * i2, r0, r1);
                    #pragma unroll
                    for (int q = 0; q < ROWS; ++q) {
                        ap[q][i2 - 1] = pair_fma<T>(z[q], r0, ap[q][i2]);
                        ap[q][i2] = pair_fma<T>(z[q], r1, ap[q][i2 + 1]);
                    }
                }
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) { ap[q][H - 1].x = ca[q]; ap[q][H - 1].y = cb[q]; }
            }
        }
        __syncwarp();

        if (valid) {
            if (l == 0 && info) info[m] = (st > n) ? 0 : st;       // a "singular" padded column cannot happen; guard anyway
            T *__restrict__ dst = io.dst(m);
            const bool bad = st != 0 && st <= n;
            if (EXACT && !bad) {                                   // the common case: 32-bit offsets, no bounds predicates
                #pragma unroll
                for (int c = 0; c < N; ++c) {
                    const int ocol = piv[c] * N;
                    #pragma unroll
                    for (int q = 0; q < ROWS; ++q) __stcs(dst + ocol + mystep[q], ((c & 1) ? ap[q][c >> 1].y : ap[q][c >> 1].x) * rscale[q]);
                }
            } else
            #pragma unroll
            for (int c = 0; c < N; ++c) {
                const int ocol = piv[c];                           // pi(c): broadcast read
                #pragma unroll
                for (int q = 0; q < ROWS; ++q) {
                    const int orow = mystep[q];                    // pi^-1(my row)
                    if (bad) { if (l + L * q < n && c < n) dst[(size_t)c * n + l + L * q] = dev_nan<T>(); }
                    else if (orow < n && ocol < n) __stcs(dst + (size_t)ocol * n + orow, ((c & 1) ? ap[q][c >> 1].y : ap[q][c >> 1].x) * rscale[q]);
                }
            }
        }
        __syncwarp();                                              // piv / lines are reused by the next matrix
    }
}

}  // namespace invgpu

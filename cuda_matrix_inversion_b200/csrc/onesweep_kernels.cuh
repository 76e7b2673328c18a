// onesweep_kernels.cuh -- SPD inverse by Cholesky with the three sweeps MERGED into one.
//
// The classic route  potrf (A = L L^T)  ->  trtri (M = L^-1)  ->  lauum (A^-1 = M^T M)  visits the
// matrix three times.  All three are sequences of rank-1 updates, and at pivot k they touch disjoint
// parts of the lower triangle:
//     rows > k , cols > k   :  a_ic -= L_ik L_ck            (potrf,  trailing matrix)
//     rows > k , cols <= k  :  acc_ic += L_ik M_kc          (trtri,  row i of M under construction)
//     rows <= k, cols <= i  :  p_ic += M_ki M_kc            (lauum,  finished part of A^-1)
// so ONE vector z_k = [ M_k,0..k | L_k+1..n-1,k ] broadcast per pivot updates the whole triangle:
//     t_ic += z_i z_c      (with the trailing matrix stored negated, t = -a, the signs all agree).
// Same arithmetic, same order of operations per element as the three-sweep route -- but one publish,
// one barrier and one pair of vector loads per pivot instead of three, and no parking of L^-1 in shared
// memory.  Shared-memory traffic (the limiter of the three-sweep kernel on B200) drops
// to about half, barriers to a third.
//
// Layout: TR x TC threads per matrix, 4x4 sub-blocks dealt cyclically (TileGeo with PERM = false).
// Pivots are taken in NATURAL order, so `info` is LAPACK's spotrf info without further ado and a 4-block
// straddles the "<= k | > k" boundary only in block k/4, which is assembled by the diagonal thread.
// Load: upper triangle only (spotrf_("U"), reference src/inverse.c:92); store: both triangles, each
// thread writing its lower blocks in natural and mirrored position, 16 bytes at a time.
#pragma once

#include "tile_kernels.cuh"

namespace invgpu {

// ------------------------------------------------------------------------------------------
// Asynchronous staging of the NEXT matrix while the current one is being inverted: every thread
// copies exactly the 16-byte chunks its register tile is built from (cp.async, L1 bypassed) into a
// private, bank-staggered strip of shared memory and picks them up at the top of the next iteration.
// The enumeration of chunks below is the one of tile_load_upper (same static skips, same twin rule).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <typename T, int N, int TR, int TC>
struct TileStage {
    using G = TileGeo<N, TR, TC, false>;
    static constexpr int SR = G::SR, SC = G::SC;
    // words per thread: the whole tile (upper-only blocks are never fetched, the space is simply unused)
    // plus one 16-byte pad so that the strips of neighbouring lanes start 4 banks apart
    static constexpr int THREAD_WORDS = SR * SC + 16 / (int)sizeof(T);
    static constexpr int MATRIX_WORDS = THREAD_WORDS * G::LANES;

    static __device__ __forceinline__ void copy4(T *dst, const T *src) {
        cp_async16(dst, src);
        if (sizeof(T) == 8) cp_async16(dst + 2, src + 2);
    }

    // issue the copies for matrix `src` into this thread's strip
    static __device__ __forceinline__ void prefetch(T *strip, const T *__restrict__ src, int ti, int tj) {
        int chunk = 0;
        #pragma unroll
        for (int g = 0; g < SR / 4; ++g) {
            #pragma unroll
            for (int h = 0; h < SC / 4; ++h) {
                if (G::cmin(4 * h) > G::rmax(4 * g + 3)) continue;
                const int br = G::rblock(g, 0) + ti, bc = G::cblock(h, 0) + tj;
                int colb, rowb;
                if (G::rblock(g, 0) > G::cblock(h, TC - 1)) { colb = br; rowb = bc; }
                else if (G::rblock(g, TR - 1) < G::cblock(h, 0)) { colb = bc; rowb = br; }
                else { const bool twin = br >= bc; colb = twin ? br : bc; rowb = twin ? bc : br; }
                #pragma unroll
                for (int p = 0; p < 4; ++p) { copy4(strip + 4 * chunk, src + (size_t)(4 * colb + p) * N + 4 * rowb); ++chunk; }
            }
        }
    }

    // build the register tile (logical-lower positions; zeros in blocks that are upper for everybody)
    static __device__ __forceinline__ void consume(T (&a)[SR][SC], const T *strip, int ti, int tj) {
        int chunk = 0;
        #pragma unroll
        for (int g = 0; g < SR / 4; ++g) {
            #pragma unroll
            for (int h = 0; h < SC / 4; ++h) {
                if (G::cmin(4 * h) > G::rmax(4 * g + 3)) {
                    #pragma unroll
                    for (int w = 0; w < 4; ++w)
                        #pragma unroll
                        for (int v = 0; v < 4; ++v) a[4 * g + w][4 * h + v] = T(0);
                    continue;
                }
                const int br = G::rblock(g, 0) + ti, bc = G::cblock(h, 0) + tj;
                T x[4][4];
                #pragma unroll
                for (int p = 0; p < 4; ++p) { ld4(strip + 4 * chunk, x[p][0], x[p][1], x[p][2], x[p][3]); ++chunk; }
                if (G::rblock(g, 0) > G::cblock(h, TC - 1)) {          // twin fetched: x[p][q] = element (row p, col q)
                    #pragma unroll
                    for (int w = 0; w < 4; ++w)
                        #pragma unroll
                        for (int v = 0; v < 4; ++v) a[4 * g + w][4 * h + v] = x[w][v];
                } else if (G::rblock(g, TR - 1) < G::cblock(h, 0)) {   // block itself: x[p][q] = element (row q, col p)
                    #pragma unroll
                    for (int w = 0; w < 4; ++w)
                        #pragma unroll
                        for (int v = 0; v < 4; ++v) a[4 * g + w][4 * h + v] = x[v][w];
                } else {
                    const bool twin = br >= bc;
                    #pragma unroll
                    for (int w = 0; w < 4; ++w)
                        #pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            T e = twin ? x[w][v] : x[v][w];
                            if (v > w) e = (br == bc) ? x[v][w] : e;
                            a[4 * g + w][4 * h + v] = e;
                        }
                }
            }
        }
    }
};

// STAGE: prefetch the next matrix through shared memory with cp.async (pays off only where the tile is
// tiny and the kernel is purely load-latency bound, n = 8; for n >= 16 the extra shared-memory traffic
// costs more than the latency it hides -- measured on B200)
template <typename T, int N, int TR, int TC, bool STAGE, typename IO, int MINB>
__global__ void __launch_bounds__((TileGeo<N, TR, TC, false>::BLOCK), MINB)
onesweep_spd_kernel(IO io, i64 batch, int *__restrict__ info) {
    using G = TileGeo<N, TR, TC, false>;
    constexpr int SR = G::SR, SC = G::SC;
    static_assert(G::LANES <= 32, "warp-sized groups only");
    using S = TileStage<T, N, TR, TC>;
    constexpr int LINE_WORDS = ((2 * N + 31) / 32) * 32 + (G::LANES < 32 ? 8 : 0);
    constexpr int WORDS = LINE_WORDS + (STAGE ? S::MATRIX_WORDS : 0);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *smem = reinterpret_cast<T *>(smem_raw);

    const int grp = threadIdx.x / G::LANES;
    const int lane = threadIdx.x % G::LANES;
    const int ti = lane / TC, tj = lane % TC;
    T *sm = smem + grp * WORDS;
    T *strip = sm + LINE_WORDS + lane * S::THREAD_WORDS;
    const int glane0 = threadIdx.x & 31 & ~(G::LANES - 1);

    if (STAGE) {   // prologue: stage the first matrix of this group
        const i64 m0 = (i64)blockIdx.x * G::MPB + grp;
        S::prefetch(strip, io.src(m0 < batch ? m0 : batch - 1), ti, tj);
    }
    #pragma unroll 1
    for (i64 base = (i64)blockIdx.x * G::MPB; base < batch; base += (i64)gridDim.x * G::MPB) {
        const i64 m = base + grp;
        const bool valid = m < batch;

        T a[SR][SC];
        if (STAGE) {
            cp_async_commit_wait_all();                              // this thread's own copies have landed
            S::consume(a, strip, ti, tj);
            // stage the next matrix of this group behind the arithmetic of the current one
            const i64 mn = m + (i64)gridDim.x * G::MPB;
            if (base + (i64)gridDim.x * G::MPB < batch) S::prefetch(strip, io.src(mn < batch ? mn : batch - 1), ti, tj);
        } else {
            tile_load_upper<T, N, TR, TC, false>(a, io.src(valid ? m : batch - 1), ti, tj);
        }
        #pragma unroll
        for (int r = 0; r < SR; ++r)
            #pragma unroll
            for (int c = 0; c < SC; ++c) a[r][c] = -a[r][c];       // trailing matrix is kept negated

        int st = 0;
        #pragma unroll
        for (int k = 0; k < N; ++k) {
            tile_lockstep<G::LANES>(k);
            T *z = sm + (k & 1) * N;
            const int q = k / 4, w = k % 4;
            const int rk = G::rowner(k), ck = G::cowner(k), srk = G::rslot(k), sk = G::cslot(k);
            const int gk = srk / 4, hk = sk / 4;
            // ---- pivot
            const T d = -__shfl_sync(0xffffffffu, a[srk][sk], glane0 + rk * TC + ck);
            if (st == 0 && !(d > T(0))) st = k + 1;
            const T rs = dev_rsqrt<T>(d);
            const T nrs = -rs;
            // ---- owners of column k: L_ik = a_ik / L_kk for rows > k, publish, restart as acc_ik
            if (tj == ck) {
                #pragma unroll
                for (int g = 0; g < SR / 4; ++g) {
                    if (G::rblock(g, TR - 1) < q) continue;          // no thread has rows > k here
                    const int blk = G::rblock(g, 0) + ti;
                    if (blk > q) {
                        st4(z + 4 * blk, a[4 * g][sk] * nrs, a[4 * g + 1][sk] * nrs, a[4 * g + 2][sk] * nrs, a[4 * g + 3][sk] * nrs);
                        a[4 * g][sk] = T(0); a[4 * g + 1][sk] = T(0); a[4 * g + 2][sk] = T(0); a[4 * g + 3][sk] = T(0);
                    }
                }
            }
            // ---- owners of row k: M_kc = -acc_kc / L_kk for cols < k, publish, restart as p_kc
            if (ti == rk) {
                #pragma unroll
                for (int h = 0; h < SC / 4; ++h) {
                    if (G::cblock(h, 0) > q) continue;               // no thread has cols <= k here
                    const int blk = G::cblock(h, 0) + tj;
                    if (blk < q) {
                        st4(z + 4 * blk, a[srk][4 * h] * nrs, a[srk][4 * h + 1] * nrs, a[srk][4 * h + 2] * nrs, a[srk][4 * h + 3] * nrs);
                        a[srk][4 * h] = T(0); a[srk][4 * h + 1] = T(0); a[srk][4 * h + 2] = T(0); a[srk][4 * h + 3] = T(0);
                    }
                }
                // ---- the diagonal thread assembles block q = [ M_k,4q..k-1 , 1/L_kk | L_k+1..4q+3,k ]
                if (tj == ck) {
                    T e[4];
                    #pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        if (v < w) { e[v] = a[srk][4 * hk + v] * nrs; a[srk][4 * hk + v] = T(0); }
                        else if (v == w) { e[v] = rs; a[srk][sk] = T(0); }
                        else { e[v] = a[4 * gk + v][sk] * nrs; a[4 * gk + v][sk] = T(0); }
                    }
                    st4(z + 4 * q, e[0], e[1], e[2], e[3]);
                }
            }
            tile_sync<G::LANES>();
            // ---- one rank-1 update of the whole lower triangle
            T x[SR], y[SC];
            #pragma unroll
            for (int g = 0; g < SR / 4; ++g)
                ld4(z + 4 * (G::rblock(g, 0) + ti), x[4 * g], x[4 * g + 1], x[4 * g + 2], x[4 * g + 3]);
            #pragma unroll
            for (int h = 0; h < SC / 4; ++h)
                ld4(z + 4 * (G::cblock(h, 0) + tj), y[4 * h], y[4 * h + 1], y[4 * h + 2], y[4 * h + 3]);
            #pragma unroll
            for (int r = 0; r < SR; ++r)
                #pragma unroll
                for (int c = 0; c < SC; ++c) {
                    if (G::cmin(c) > G::rmax(r)) continue;           // strictly upper for every thread
                    a[r][c] = fma(x[r], y[c], a[r][c]);
                }
        }
        tile_sync<G::LANES>();                                       // the lines are reused by the next matrix

        if (!valid) continue;
        if (lane == 0 && info) info[m] = st;
        T *__restrict__ dst = io.dst(m);
        #pragma unroll
        for (int g = 0; g < SR / 4; ++g) {
            #pragma unroll
            for (int h = 0; h < SC / 4; ++h) {
                const int br = G::rblock(g, 0) + ti, bc = G::cblock(h, 0) + tj;
                if (st) {                                             // flagged: this thread's natural blocks, all NaN
                    #pragma unroll
                    for (int v = 0; v < 4; ++v)
                        stg4(dst + (size_t)(4 * bc + v) * N + 4 * br, dev_nan<T>(), dev_nan<T>(), dev_nan<T>(), dev_nan<T>());
                    continue;
                }
                if (G::cblock(h, 0) > G::rblock(g, TR - 1)) continue; // strictly upper for every thread
                if (br == bc) {                                       // diagonal block: symmetrise in registers
                    #pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        T e[4];
                        #pragma unroll
                        for (int ww = 0; ww < 4; ++ww) e[ww] = (ww >= v) ? a[4 * g + ww][4 * h + v] : a[4 * g + v][4 * h + ww];
                        stg4(dst + (size_t)(4 * bc + v) * N + 4 * br, e[0], e[1], e[2], e[3]);
                    }
                } else if (br > bc) {
                    #pragma unroll
                    for (int v = 0; v < 4; ++v)                       // natural position
                        stg4(dst + (size_t)(4 * bc + v) * N + 4 * br, a[4 * g][4 * h + v], a[4 * g + 1][4 * h + v],
                             a[4 * g + 2][4 * h + v], a[4 * g + 3][4 * h + v]);
                    #pragma unroll
                    for (int ww = 0; ww < 4; ++ww)                    // mirror image
                        stg4(dst + (size_t)(4 * br + ww) * N + 4 * bc, a[4 * g + ww][4 * h], a[4 * g + ww][4 * h + 1],
                             a[4 * g + ww][4 * h + 2], a[4 * g + ww][4 * h + 3]);
                }
            }
        }
    }
}

}  // namespace invgpu

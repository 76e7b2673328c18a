// tc_kernels.cuh -- n = 128 fp32 on the 5th-generation tensor cores: blocked factorisation with the
// MATRIX ITSELF AS THE TMEM ACCUMULATOR (128 lanes x 128 columns of fp32 = one 128x128 matrix; lane = row).
//
// north_star: "Tensor cores are used only for the trailing-update GEMM of blocked n >= 64 factorizations, and
// only where ncu shows it wins."  This is that kernel for the fused GP mean / variance (BASELINE configs[3];
// replaces addDiagonal -> getrf/getriBatched -> gemv -> dot of reference src/gauss_bench.cu:38-265 like the CUDA-core
// kernels of sweep_kernels.cuh do):
//
//   one CTA of 128 threads per evaluation, thread t = row t of B + diag C, warp w <-> TMEM lanes 32w..32w+31.
//   Right-looking blocked Cholesky, panel width 32:
//     panel p:  tcgen05.ld the 32 panel columns of every row (lane = row: 32 registers per thread)
//               warp p factors the 32x32 diagonal block (lane = row, column k published through shared memory,
//               rsqrt + one Newton step per pivot) and forward-substitutes both right-hand sides along the way
//               warps > p solve their rows against it (row TRSM: 496 FMAs per thread, broadcast LDS operands)
//               every row writes its 32 multipliers to shared memory as a K-major UMMA operand, split
//               x = hi + lo with hi = the TF32-representable head (3xTF32: hi*hi + hi*lo + lo*hi ~ fp32 accuracy)
//               one thread issues 12 tcgen05.mma.kind::tf32 (4 K-steps x 3 split products, M = 128,
//               N = 128 - 32(p+1), A negated by the instruction descriptor):  T[:, 32(p+1):] -= L_p L_p^T,
//               tcgen05.commit -> mbarrier; the next panel's tcgen05.ld waits on it
//   so the n^3/3 flops of the factorisation shrink to the 34 % that are panel work on the CUDA cores; the
//   right-hand sides ride along in registers (32 FMAs per thread, vector and panel) and the result is
//   sum_k y_k(a) y_k(d) -- no inverse, no solve vectors and no intermediates reach HBM.
//   Only the UPPER triangle of B is read (spotrf_("U") convention of the reference's CPU path, src/gauss_cpu.c:54):
//   row i of the lower triangle = the first i+1 elements of column i of the column-major input.
//   Pivots are taken in natural order, so `info` is spotrf's directly.
#pragma once

#include <cuda_runtime.h>
#ifndef INVGPU_TC_STAGE
#define INVGPU_TC_STAGE 0
#endif

#include <stdint.h>

#include "common.cuh"

namespace invgpu {
namespace tc {

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
// bounded wait: a protocol error traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t a = smem_addr(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return;
        if (spin > (1u << 26)) __trap();
    }
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (bytes: multiple of 16, both sides 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols) {      // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {    // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async;" ::: "memory"); }

// 32 consecutive columns of the calling thread's TMEM lane <-> 32 registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    #pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    #pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 :: "r"(taddr),
                    "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                    "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                    "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                    "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
                    "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
                    "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
                    "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
                    "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared-memory operand descriptor, K-major, no swizzle ("interleave"): core matrices of 8 rows x 16 bytes are
// 128 contiguous bytes; the next 8-row group lies SBO bytes further, the next 16-byte K chunk LBO bytes further.
// Bit layout as in cute::UMMA::SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start >> 4 in [0,14), LBO >> 4 in
// [16,30), SBO >> 4 in [32,46), version 1 in [46,48), layout type 0 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// Instruction descriptor of kind::tf32 (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10),
// negate A (1 << 13), both K-major, N >> 3 in [17,23), M >> 4 in [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n, bool negate_a) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((negate_a ? 1u : 0u) << 13) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}

constexpr int GP_N = 128;          // matrix order == TMEM lanes == threads per CTA
constexpr int GP_CHUNK_STRIDE = GP_N * 16;                       // LBO: bytes between 16-byte K chunks of an operand (2048)

template <int PW>
struct __align__(16) GpShared {
    float lcol[PW * PW];           // lcol[k * PW + i] = L11(k + 1 + i, k): column k of the current diagonal block, aligned
    float rinv[PW];                // 1 / L11(k, k)
    float ya[PW], yd[PW];          // forward-substituted right-hand sides of the current panel
    float red[2 * 4];
    uint64_t mma_done;
    uint64_t stage_full;           // the staged chunk of the next evaluation has landed (bulk copies, complete_tx)
    uint32_t tmem_base;
    int info;
};
template <int PW> struct GpGeo {
    static constexpr int PANEL_BYTES = GP_N * PW * 4;            // one operand copy (hi or lo)
    static constexpr size_t SMEM_BYTES = 2 * PANEL_BYTES + sizeof(GpShared<PW>);
};

// x[j] -= l * L11(j, K) for the columns j > K of the panel; col[i] = L11(K + 1 + i, K) (broadcast 128-bit reads).
// Triangular and fully unrolled: every register index is static and only the FMAs that matter are issued (a rolled
// full-width form with a rotating register window doubled both the FMAs and the shared-memory wavefronts, which are
// what bounds this kernel: one broadcast word per FMA).
template <int PW, int K>
struct Tail {
    static constexpr int NQ = (PW - K - 1 + 3) / 4;
    float4 v[NQ > 0 ? NQ : 1];
    __device__ __forceinline__ void load(const float *col) {
        #pragma unroll
        for (int q = 0; q < NQ; ++q) v[q] = reinterpret_cast<const float4 *>(col)[q];
    }
    __device__ __forceinline__ void apply(float (&x)[PW], float l) const {
        #pragma unroll
        for (int q = 0; q < NQ; ++q) {
            if (K + 1 + 4 * q + 0 < PW) x[K + 1 + 4 * q + 0] = fmaf(-l, v[q].x, x[K + 1 + 4 * q + 0]);
            if (K + 1 + 4 * q + 1 < PW) x[K + 1 + 4 * q + 1] = fmaf(-l, v[q].y, x[K + 1 + 4 * q + 1]);
            if (K + 1 + 4 * q + 2 < PW) x[K + 1 + 4 * q + 2] = fmaf(-l, v[q].z, x[K + 1 + 4 * q + 2]);
            if (K + 1 + 4 * q + 3 < PW) x[K + 1 + 4 * q + 3] = fmaf(-l, v[q].w, x[K + 1 + 4 * q + 3]);
        }
    }
};
template <int PW, int K>
__device__ __forceinline__ void rank1_tail(float (&x)[PW], float l, const float *col) {
    Tail<PW, K> tl;
    tl.load(col);
    tl.apply(x, l);
}

// reciprocal square root of a pivot (rsqrt.approx + one Newton step: full fp32 accuracy); a non-positive / NaN pivot is
// recorded (spotrf's info = the first one) and replaced by 1 so that the rest of the panel stays finite
__device__ __forceinline__ float pivot_rsqrt_newton(float d, int k1, int &bad_at) {
    const bool bad = !(d > 0.f);                                  // uniform in the warp
    bad_at = (bad && bad_at == 0) ? k1 : bad_at;
    d = bad ? 1.f : d;
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return r * fmaf(-0.5f * d * r, r, 1.5f);
}

// Per-evaluation register state of a row-thread.
template <int PW>
struct RowState {
    float x[PW];                   // the row's entries in the current panel -> its multipliers
    float ua, ud;                  // right-hand sides a, d of this row (forward substitution in progress)
    float ya, yd;                  // their finished values (set in the row's own panel)
    float dself;                   // the row's own diagonal element, kept up to date locally inside its diagonal block
    float rkeep;                   // 1 / L_kk of the row's own pivot (kept by the pivot lane, stored once per block)
    int bad_at;
};

// The warp that owns the diagonal block of a panel (lanes lane0 .. lane0 + PW - 1 = the block's rows; for PW = 16 the
// other half-warp holds rows below the block, which simply follow as TRSM rows, or finished rows above it, whose
// registers are don't-cares).  Pivot k:  shfl(dself, pivot lane) -> rsqrt -> l = x_k * r -> dself -= l^2: every lane
// keeps its OWN diagonal element up to date locally, so the pivot chain never waits for the shared-memory exchange of
// column k (store l, __syncwarp, broadcast loads, rank-1 tail), which runs beside it.  x[k] becomes the multiplier.
// NM evaluations are advanced TOGETHER: their pivot chains are independent, so one warp interleaves NM dependency
// chains and shares every __syncwarp / __syncthreads between them (this phase is latency-bound: 54 % of the time line of
// the single-evaluation kernel, profiles/r2_tc_gp128_summary.md).
// Measured alternatives: starting pivot k+1's shuffle + rsqrt before step k's column loads (-5 %), fetching L(k+1, k) by a
// fourth shuffle (-13 %: the SM retires one warp shuffle per clock), staggering the CTAs of an SM in time (0 %).
template <int PW, int NM, int K>
struct DiagStep {
    static __device__ __forceinline__ void run(RowState<PW> (&s)[NM], GpShared<PW> *sh, int lane, int lane0) {
        const int pl = lane0 + K;                                 // pivot lane
        const int jb = lane - lane0;                              // row inside the block
        float l[NM];
        #pragma unroll
        for (int e = 0; e < NM; ++e) {
            const float r = pivot_rsqrt_newton(__shfl_sync(0xffffffffu, s[e].dself, pl), K + 1, s[e].bad_at);
            l[e] = s[e].x[K] * r;
            s[e].x[K] = l[e];
            s[e].dself = fmaf(-l[e], l[e], s[e].dself);
            s[e].rkeep = (lane == pl) ? r : s[e].rkeep;                      // one store per block instead of one per pivot
            if (jb > K && jb < PW) sh[e].lcol[K * PW + jb - K - 1] = l[e];   // aligned column: entry i = L11(K + 1 + i, K)
        }
        __syncwarp();
        Tail<PW, K> tl[NM];
        #pragma unroll
        for (int e = 0; e < NM; ++e) tl[e].load(sh[e].lcol + K * PW);
        #pragma unroll
        for (int e = 0; e < NM; ++e) tl[e].apply(s[e].x, l[e]);
        DiagStep<PW, NM, K + 1>::run(s, sh, lane, lane0);
    }
};
template <int PW, int NM> struct DiagStep<PW, NM, PW> {
    static __device__ __forceinline__ void run(RowState<PW> (&)[NM], GpShared<PW> *, int, int) {}
};

// Forward substitution of both right-hand sides through the diagonal block, by the warp that owns it, AFTER the block is
// factored and WHILE the other warps run their row TRSMs (this warp has no TRSM rows): y_k = u_k / L_kk in the pivot lane,
// broadcast by shuffle, u_i -= L_ik y_k in the rows below it (x[K] is the lane's multiplier).  Inside the pivot loop these
// two shuffles stalled every step of the in-order pivot chain: 18 % of the kernel (timing experiment, r2_tc_gp128_summary.md).
template <int PW, int NM, int K>
struct DiagRhsStep {
    static __device__ __forceinline__ void run(RowState<PW> (&s)[NM], GpShared<PW> *sh, int lane, int lane0) {
        const int pl = lane0 + K;
        #pragma unroll
        for (int e = 0; e < NM; ++e) {
            const float r = sh[e].rinv[K];
            const float ta = __shfl_sync(0xffffffffu, s[e].ua * r, pl);
            const float td = __shfl_sync(0xffffffffu, s[e].ud * r, pl);
            s[e].ya = (lane == pl) ? ta : s[e].ya;
            s[e].yd = (lane == pl) ? td : s[e].yd;
            s[e].ua = fmaf(-s[e].x[K], ta, s[e].ua);
            s[e].ud = fmaf(-s[e].x[K], td, s[e].ud);
        }
        DiagRhsStep<PW, NM, K + 1>::run(s, sh, lane, lane0);
    }
};
template <int PW, int NM> struct DiagRhsStep<PW, NM, PW> {
    static __device__ __forceinline__ void run(RowState<PW> (&)[NM], GpShared<PW> *, int, int) {}
};

// a row below the diagonal block (another warp): l_ik = x_k / L_kk and the row's own rank-1 tail
template <int PW, int NM, int K>
struct TrsmStep {
    static __device__ __forceinline__ void run(RowState<PW> (&s)[NM], const GpShared<PW> *sh) {
        float l[NM];
        Tail<PW, K> tl[NM];
        #pragma unroll
        for (int e = 0; e < NM; ++e) {
            l[e] = s[e].x[K] * sh[e].rinv[K];                              // (the compiler merges these into 128-bit broadcast loads)
            s[e].x[K] = l[e];
            tl[e].load(sh[e].lcol + K * PW);
        }
        #pragma unroll
        for (int e = 0; e < NM; ++e) tl[e].apply(s[e].x, l[e]);
        TrsmStep<PW, NM, K + 1>::run(s, sh);
    }
};
template <int PW, int NM> struct TrsmStep<PW, NM, PW> {
    static __device__ __forceinline__ void run(RowState<PW> (&)[NM], const GpShared<PW> *) {}
};

// the right-hand sides of a row below the block, once the block's y are known: u_i -= sum_k L_ik y_k
template <int PW>
__device__ __forceinline__ void rhs_row_update(RowState<PW> &s, const GpShared<PW> &sh) {
    #pragma unroll
    for (int c = 0; c < PW / 4; ++c) {
        const float4 a = reinterpret_cast<const float4 *>(sh.ya)[c], d = reinterpret_cast<const float4 *>(sh.yd)[c];
        s.ua = fmaf(-s.x[4 * c + 0], a.x, s.ua); s.ud = fmaf(-s.x[4 * c + 0], d.x, s.ud);
        s.ua = fmaf(-s.x[4 * c + 1], a.y, s.ua); s.ud = fmaf(-s.x[4 * c + 1], d.y, s.ud);
        s.ua = fmaf(-s.x[4 * c + 2], a.z, s.ua); s.ud = fmaf(-s.x[4 * c + 2], d.z, s.ud);
        s.ua = fmaf(-s.x[4 * c + 3], a.w, s.ua); s.ud = fmaf(-s.x[4 * c + 3], d.w, s.ud);
    }
}

template <int PW> __device__ __forceinline__ void tmem_ld_panel(uint32_t taddr, float (&x)[PW]);
template <> __device__ __forceinline__ void tmem_ld_panel<32>(uint32_t taddr, float (&x)[32]) { tmem_ld32(taddr, x); }
template <> __device__ __forceinline__ void tmem_ld_panel<16>(uint32_t taddr, float (&x)[16]) { tmem_ld16(taddr, x); }

template <int PW, int NM> struct GpGeoM {
    static constexpr int PANEL_BYTES = GP_N * PW * 4;            // one operand copy (hi or lo) of one evaluation
    static constexpr int TMEM_COLS = 128 * NM;
    // next evaluation staged through shared memory, chunk by chunk, while the current one is factored (INVGPU_TC_STAGE=1 at
    // build time).  Measured: 9.56 ms vs 8.43 ms without it (200 000 evaluations) -- the kernel is bound by the SM's shared-
    // memory / LSU data pipe (63 % busy, profiles/r2_tc_gp128_summary.md), and the stage adds traffic there; off by default.
    static constexpr bool STAGED = (INVGPU_TC_STAGE != 0 && NM == 1 && PW == 16);
    static constexpr int STAGE_STRIDE = 144;                     // bytes per staged row: 128 + 16 (conflict-free 128-bit reads)
    static constexpr int STAGE_BYTES = STAGED ? GP_N * STAGE_STRIDE : 0;
    static constexpr size_t SMEM_BYTES = (size_t)NM * 2 * PANEL_BYTES + STAGE_BYTES + NM * sizeof(GpShared<PW>) + 64;
};

// Fused GP mean / variance, n = 128 fp32, panel width PW (16 or 32), NM (1 or 2) evaluations per CTA advanced together
// (evaluation e of the CTA's group lives in TMEM columns 128 e .. 128 e + 127).  grid = persistent, block = 128.
template <int PW, int NM, int MINB>
__global__ void __launch_bounds__(128, MINB)
tc_gp128_kernel(GpIO<float> io, i64 batch, int *__restrict__ info) {
    using Geo = GpGeoM<PW, NM>;
    constexpr int NP = GP_N / PW;                                 // panels
    // no pointer arithmetic through integers here: the compiler must keep seeing the shared address space (a generic
    // LD.E / ST.E costs ~30 clocks more than LDS / STS, and these accesses sit beside the pivot chain)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *ops = smem_raw;                                // [NM][hi, lo][PANEL_BYTES]
    GpShared<PW> *sh = reinterpret_cast<GpShared<PW> *>(smem_raw + NM * 2 * Geo::PANEL_BYTES + Geo::STAGE_BYTES);
    uint64_t *mma_done = &sh[0].mma_done;
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;

    if (w == 0) tmem_alloc(&sh[0].tmem_base, Geo::TMEM_COLS);
    if (t == 0) { mbar_init(mma_done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sh[0].tmem_base;
    const uint32_t my_lane = tmem + ((uint32_t)(32 * w) << 16);           // this warp's 32 TMEM lanes
    uint32_t phase = 0;
    // where this row's multipliers go inside an operand copy (K-major core matrices, see umma_desc_kmajor)
    const uint32_t row_off = (uint32_t)(t >> 3) * 128u + (uint32_t)(t & 7) * 16u;

    // STAGED (one evaluation per CTA, PW = 16): chunk c (columns 32c..32c+31) of the NEXT evaluation is fetched while the
    // current one is still being factored.  Once the two panels covering those columns are done no MMA touches them again,
    // so the TMEM cells are free: right after that panel's first barrier every row that needs the chunk issues ONE 128-byte
    // bulk copy (cp.async.bulk, no registers, no warp waits) into a shared-memory stage; a panel later the rows move their
    // 32 values stage -> registers -> TMEM.  Only the very first evaluation of a CTA is loaded up front (the load phase is
    // 24 % of a CTA's time line).  Measured slower than loading up front (see GpGeoM::STAGED), like fetching through registers
    // (9.70 ms): four CTAs per SM already hide each other's load phases, and the SM-level bound is the LSU data pipe.
    constexpr bool STAGED = Geo::STAGED;
    unsigned char *stage = smem_raw + NM * 2 * Geo::PANEL_BYTES;
    bool prefetched = false;
    uint32_t stage_phase = 0;
    if (STAGED && t == 0) { mbar_init(&sh[0].stage_full, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (STAGED) __syncthreads();
    for (i64 m0 = (i64)blockIdx.x * NM; m0 < batch; m0 += (i64)gridDim.x * NM) {
        RowState<PW> s[NM];
        i64 mm[NM];
        #pragma unroll
        for (int e = 0; e < NM; ++e) {
            mm[e] = (m0 + e < batch) ? m0 + e : batch - 1;                    // a ragged last group recomputes the last evaluation
            const i64 m = mm[e];
            const float *__restrict__ brow = io.b + m * (GP_N * GP_N) + (i64)t * GP_N;   // column t == row t of the symmetric input
            s[e].ua = io.a[m * GP_N + t];
            s[e].ud = io.d ? io.d[m * GP_N + t] : s[e].ua;
            s[e].ya = 0.f; s[e].yd = 0.f; s[e].bad_at = 0; s[e].dself = 0.f; s[e].rkeep = 1.f;
            if (t == 0) sh[e].info = 0;
            if (STAGED && prefetched) continue;
            const float cdiag = io.c[m * GP_N + t];
            // ---- load: columns 32c..32c+31 of this row for c <= w (lower triangle), + diag C, into this row's TMEM lane
            float v32[32];
            #pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c <= w) {                                                 // warp-uniform
                    #pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 v = __ldg(reinterpret_cast<const float4 *>(brow + 32 * c) + q);
                        v32[4 * q + 0] = v.x; v32[4 * q + 1] = v.y; v32[4 * q + 2] = v.z; v32[4 * q + 3] = v.w;
                    }
                    if (c == w) {
                        #pragma unroll
                        for (int j = 0; j < 32; ++j) if (j == lane) v32[j] += cdiag;
                    }
                    tmem_st32(my_lane + 128 * e + 32 * c, v32);
                }
            }
        }
        tmem_st_wait();
        const i64 m_next = m0 + (i64)gridDim.x * NM;
        const bool has_next = STAGED && m_next < batch;
        int staged_chunk = -1;                                                // chunk sitting in the stage, not yet in TMEM
        const float *__restrict__ nrow = io.b + (has_next ? m_next : m0) * (GP_N * GP_N) + (i64)t * GP_N;

        #pragma unroll 1
        for (int p = 0; p < NP; ++p) {
            const int wd = (p * PW) >> 5;                                     // the warp that owns the diagonal block
            const int lane0 = (p * PW) & 31;                                  // its first lane
            if (STAGED && t == 0 && has_next && ((p + 1) * PW) % 32 == 0)     // bytes the rows will copy after this panel's first barrier
                mbar_expect_tx(&sh[0].stage_full, (uint32_t)(4 - (((p + 1) * PW) / 32 - 1)) * 32u * 128u);
            if (w >= wd) {
                if (p > 0) { mbar_wait(mma_done, phase); tc_fence_after(); }  // the updates of panel p-1 have landed
                #pragma unroll
                for (int e = 0; e < NM; ++e) tmem_ld_panel<PW>(my_lane + 128 * e + PW * p, s[e].x);
            }
            if (p > 0) phase ^= 1;                                            // every thread tracks the barrier's phase
            if (w == wd) {
                #pragma unroll
                for (int e = 0; e < NM; ++e) {
                    float dself = 0.f;                                        // this lane's own diagonal element
                    #pragma unroll
                    for (int j = 0; j < PW; ++j) dself = (lane - lane0 == j) ? s[e].x[j] : dself;
                    s[e].dself = dself; s[e].bad_at = 0;
                }
                DiagStep<PW, NM, 0>::run(s, sh, lane, lane0);
                #pragma unroll
                for (int e = 0; e < NM; ++e) {
                    if (lane >= lane0 && lane < lane0 + PW) sh[e].rinv[lane - lane0] = s[e].rkeep;
                    if (lane == 0 && s[e].bad_at != 0 && sh[e].info == 0) sh[e].info = p * PW + s[e].bad_at;
                }
            }
            __syncthreads();
            // the columns of chunk `pc` are final for every row now: start fetching that chunk of the next evaluation
            const int pc = (has_next && ((p + 1) * PW) % 32 == 0) ? ((p + 1) * PW) / 32 - 1 : -1;
            if (STAGED && pc >= 0) {
                if (pc <= w) bulk_g2s(stage + t * Geo::STAGE_STRIDE, nrow + 32 * pc, 128, &sh[0].stage_full);
                staged_chunk = pc;
            }
            if (w > wd) TrsmStep<PW, NM, 0>::run(s, sh);
            else if (w == wd) {                                               // in the shadow of the other warps' TRSM
                DiagRhsStep<PW, NM, 0>::run(s, sh, lane, lane0);
                #pragma unroll
                for (int e = 0; e < NM; ++e)
                    if (lane >= lane0 && lane < lane0 + PW) { sh[e].ya[lane - lane0] = s[e].ya; sh[e].yd[lane - lane0] = s[e].yd; }
            }
            if (w >= wd && p < NP - 1) {                                      // publish the rows' multipliers as hi + lo
                #pragma unroll
                for (int e = 0; e < NM; ++e) {
                    unsigned char *l_hi = ops + (size_t)(2 * e) * Geo::PANEL_BYTES, *l_lo = l_hi + Geo::PANEL_BYTES;
                    #pragma unroll
                    for (int c = 0; c < PW / 4; ++c) {
                        float4 hi, lo;
                        hi.x = __uint_as_float(__float_as_uint(s[e].x[4 * c + 0]) & 0xffffe000u); lo.x = s[e].x[4 * c + 0] - hi.x;
                        hi.y = __uint_as_float(__float_as_uint(s[e].x[4 * c + 1]) & 0xffffe000u); lo.y = s[e].x[4 * c + 1] - hi.y;
                        hi.z = __uint_as_float(__float_as_uint(s[e].x[4 * c + 2]) & 0xffffe000u); lo.z = s[e].x[4 * c + 2] - hi.z;
                        hi.w = __uint_as_float(__float_as_uint(s[e].x[4 * c + 3]) & 0xffffe000u); lo.w = s[e].x[4 * c + 3] - hi.w;
                        *reinterpret_cast<float4 *>(l_hi + c * GP_CHUNK_STRIDE + row_off) = hi;
                        *reinterpret_cast<float4 *>(l_lo + c * GP_CHUNK_STRIDE + row_off) = lo;
                    }
                }
                fence_async_smem();                                           // generic-proxy writes -> visible to the tensor core
            }
            tc_fence_before();
            __syncthreads();
            if (w > wd) {                                                     // the block's y are complete: rows below catch up
                #pragma unroll
                for (int e = 0; e < NM; ++e) rhs_row_update<PW>(s[e], sh[e]);
            }
            if (p < NP - 1 && t == 0) {
                tc_fence_after();
                const int n_cols = GP_N - PW * (p + 1);
                const uint32_t idesc = umma_idesc_tf32(128, n_cols, true);
                const uint32_t b_rows = (uint32_t)(PW * (p + 1) / 8) * 128u;   // B operand = rows PW(p+1).. of the same panel
                #pragma unroll
                for (int e = 0; e < NM; ++e) {
                    const uint32_t d_addr = tmem + 128 * e + PW * (p + 1);
                    const uint32_t s_hi = smem_addr(ops + (size_t)(2 * e) * Geo::PANEL_BYTES), s_lo = s_hi + Geo::PANEL_BYTES;
                    #pragma unroll
                    for (int ks = 0; ks < PW / 8; ++ks) {
                        const uint32_t koff = (uint32_t)ks * 2u * GP_CHUNK_STRIDE;
                        const uint64_t a_hi = umma_desc_kmajor(s_hi + koff, GP_CHUNK_STRIDE, 128);
                        const uint64_t a_lo = umma_desc_kmajor(s_lo + koff, GP_CHUNK_STRIDE, 128);
                        const uint64_t b_hi = umma_desc_kmajor(s_hi + koff + b_rows, GP_CHUNK_STRIDE, 128);
                        const uint64_t b_lo = umma_desc_kmajor(s_lo + koff + b_rows, GP_CHUNK_STRIDE, 128);
                        umma_tf32(d_addr, a_lo, b_hi, idesc, 1u);             // small terms first
                        umma_tf32(d_addr, a_hi, b_lo, idesc, 1u);
                        umma_tf32(d_addr, a_hi, b_hi, idesc, 1u);
                    }
                }
                umma_commit(mma_done);
            }
            if (STAGED && staged_chunk >= 0 && staged_chunk != pc) {          // staged one panel ago: stage -> registers -> TMEM
                const int c = staged_chunk;
                if (c <= w) {
                    mbar_wait(&sh[0].stage_full, stage_phase);
                    float v32[32];
                    #pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 v = reinterpret_cast<const float4 *>(stage + t * Geo::STAGE_STRIDE)[q];
                        v32[4 * q + 0] = v.x; v32[4 * q + 1] = v.y; v32[4 * q + 2] = v.z; v32[4 * q + 3] = v.w;
                    }
                    if (c == w) {
                        const float cnext = io.c[m_next * GP_N + t];
                        #pragma unroll
                        for (int j = 0; j < 32; ++j) if (j == lane) v32[j] += cnext;
                    }
                    tmem_st32(my_lane + 32 * c, v32);
                    fence_async_smem();                                       // the stage is rewritten by bulk copies later
                }
                stage_phase ^= 1;
                staged_chunk = -1;
            }
        }
        if (STAGED && staged_chunk >= 0) {                                    // the last chunk (rows 96..127 only) of the next evaluation
            const int c = staged_chunk;
            if (c <= w) {
                mbar_wait(&sh[0].stage_full, stage_phase);
                float v32[32];
                #pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 v = reinterpret_cast<const float4 *>(stage + t * Geo::STAGE_STRIDE)[q];
                    v32[4 * q + 0] = v.x; v32[4 * q + 1] = v.y; v32[4 * q + 2] = v.z; v32[4 * q + 3] = v.w;
                }
                if (c == w) {
                    const float cnext = io.c[m_next * GP_N + t];
                    #pragma unroll
                    for (int j = 0; j < 32; ++j) if (j == lane) v32[j] += cnext;
                }
                tmem_st32(my_lane + 32 * c, v32);
                fence_async_smem();
            }
            stage_phase ^= 1;
            staged_chunk = -1;
        }
        if (STAGED && has_next) prefetched = true;
        // ---- epilogue: means = sum ya*yd, variances = E - sum ya^2 (every row finalised its y in its own panel)
        #pragma unroll
        for (int e = 0; e < NM; ++e) {
            float pm = s[e].ya * s[e].yd, pq = s[e].ya * s[e].ya;
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) { pm += __shfl_xor_sync(0xffffffffu, pm, o); pq += __shfl_xor_sync(0xffffffffu, pq, o); }
            if (lane == 0) { sh[e].red[w] = pm; sh[e].red[4 + w] = pq; }
        }
        __syncthreads();
        if (t < NM && m0 + t < batch) {
            const GpShared<PW> &r = sh[t];
            const i64 m = m0 + t;
            const int st = r.info;
            const float sm = (r.red[0] + r.red[1]) + (r.red[2] + r.red[3]);
            const float sq = (r.red[4] + r.red[5]) + (r.red[6] + r.red[7]);
            if (info) info[m] = st;
            if (io.means) io.means[m] = st ? dev_nan<float>() : sm;
            if (io.variances) io.variances[m] = st ? dev_nan<float>() : io.e[m] - sq;
        }
        __syncthreads();                                                     // red / info reuse; all tensor-core work of this group is done
    }
    tc_fence_before();
    __syncthreads();
    if (w == 0) tmem_dealloc(tmem, Geo::TMEM_COLS);
}

}  // namespace tc
}  // namespace invgpu

// mixed_kernels.cuh -- persistent-CTA scheduler over batches of MIXED dimensions.
//
// The reference only sketches this case: "size-bucketed stream queues" (README.md:41-44, buckets
// 32 / 128 / 512 / 1024; BASELINE.json narrows them to 32 / 128 / 256).  Here each bucket is a work
// list drained by a persistent grid: CTAs (bucket 128 / 256: one matrix per CTA) or warps (bucket
// 32: one matrix per warp) pull the next work unit with an atomic ticket until the list is empty,
// so a few large matrices never leave SMs idle behind a static partition.  Lists are sorted by
// descending n (longest work first).  The three bucket kernels run concurrently on three streams.
// Per matrix the math is the any-n shared-memory tier (generic_smem.cuh): potrf -> trtri -> lauum
// in packed storage, spotrf info semantics, NaN output for flagged matrices.
#pragma once

#include "generic_smem.cuh"

namespace invgpu {

// G = 32: one matrix per warp, 8 warps per CTA; G = 128 / 256: one matrix per CTA.
template <typename T, int G>
__global__ void __launch_bounds__(G <= 32 ? 256 : G)
mixed_spd_kernel(const MixedItem *__restrict__ items, i64 count, int nmax, int *__restrict__ info,
                 unsigned long long *__restrict__ ticket, T *gws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long s_unit;
    constexpr int PER_UNIT = (G <= 32) ? 8 : 1;                  // matrices per work unit
    const int g = (G <= 32) ? threadIdx.x / 32 : 0;
    const int t = (G <= 32) ? threadIdx.x % 32 : threadIdx.x;
    // gws != null: working copy in a per-CTA slab of global memory (fp64 with nmax > 236 exceeds shared memory)
    T *S = gws ? gws + ((size_t)blockIdx.x * PER_UNIT + g) * packed_row(nmax)
               : reinterpret_cast<T *>(smem_raw) + (size_t)g * packed_row(nmax);

    for (;;) {
        if (threadIdx.x == 0) s_unit = atomicAdd(ticket, 1ULL);
        __syncthreads();
        const i64 unit = (i64)s_unit;
        __syncthreads();                                          // everyone has read s_unit before the next grab
        if (unit * PER_UNIT >= count) break;
        const i64 m = unit * PER_UNIT + g;
        if (m >= count) continue;                                 // warp tier only: ragged last unit
        const MixedItem it = items[m];
        const int n = it.n, nn = n * n;
        const T *src = static_cast<const T *>(it.in);
        T *dst = static_cast<T *>(it.out);
        for (int idx = t; idx < nn; idx += G) {                   // upper triangle -> L(c, r)
            const int c = idx / n, r = idx - c * n;
            if (r <= c) S[packed_row(c) + r] = src[idx];
        }
        Group<G>::sync();
        const int st = potrf_packed<T, G>(S, n, n, t);
        if (t == 0 && info) info[it.index] = st;
        if (st) { fill_nan<T, G>(dst, nn, t); Group<G>::sync(); continue; }
        trtri_packed<T, G>(S, n, t);
        lauum_packed<T, G>(S, n, t);
        for (int idx = t; idx < nn; idx += G) {
            const int c = idx / n, r = idx - c * n;
            dst[idx] = r >= c ? S[packed_row(r) + c] : S[packed_row(c) + r];
        }
        Group<G>::sync();                                         // slab reuse
    }
}

}  // namespace invgpu

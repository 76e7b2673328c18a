// common.cuh -- shared device/host helpers for the batched-inversion engine (sm_100a).
//
// Layout contract (reference README.md:46-52, src/helper.cu:45,103-118): every matrix is
// column-major with lda == n.  Batches reach the kernels through an "IO policy":
//   StridedIO : matrix k at base + k*stride      (dense, what the *_gpu host wrappers build)
//   PtrIO     : matrix k at ptrs[k]              (the reference's `Array *devAs` flavour; the
//               pointer array may live in pinned host memory, read through UVA like the
//               reference kernels do, src/gauss/batched_invert.cu:120)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace invgpu {

typedef long long i64;

template <typename T>
struct StridedIO {
    const T *in;
    T *out;
    i64 in_stride;   // elements between consecutive input matrices
    i64 out_stride;  // elements between consecutive outputs
    __device__ __forceinline__ const T *src(i64 k) const { return in + k * in_stride; }
    __device__ __forceinline__ T *dst(i64 k) const { return out + k * out_stride; }
};

template <typename T>
struct PtrIO {
    T *const *in;
    T *const *out;
    __device__ __forceinline__ const T *src(i64 k) const { return in[k]; }
    __device__ __forceinline__ T *dst(i64 k) const { return out[k]; }
};

// Mixed-dimension batches: one work item per matrix (device pointers, column-major n x n, lda = n).
struct MixedItem {
    const void *in;
    void *out;
    int n;
    int index;          // position in the caller's arrays (for info[])
};
// IO policy of the padded tiers: matrix k of order n <= N is embedded as blockdiag(A, I) in the N x N
// register tile of a fixed-order kernel ((A (+) I)^-1 = A^-1 (+) I; the identity part costs nothing but
// the FMAs of its pivots); loads / stores are bounds-checked scalar accesses (lda = n: no alignment).
template <typename T>
struct PadIO {
    // exactly one addressing mode is set: a work list (mixed dimensions), per-matrix pointer arrays
    // (the reference's `Array *devAs` flavour, uniform order n) or a strided dense batch (uniform n)
    const MixedItem *items = nullptr;
    T *const *in_ptrs = nullptr;
    T *const *out_ptrs = nullptr;
    const T *in_base = nullptr;
    T *out_base = nullptr;
    i64 in_stride = 0, out_stride = 0;
    int n = 0;
    __device__ __forceinline__ void get(i64 k, const T *&src, T *&dst, int &order, i64 &info_index) const {
        if (items) {
            const MixedItem it = items[k];
            src = static_cast<const T *>(it.in); dst = static_cast<T *>(it.out); order = it.n; info_index = it.index;
        } else if (in_ptrs) {
            src = in_ptrs[k]; dst = out_ptrs[k]; order = n; info_index = k;
        } else {
            src = in_base + k * in_stride; dst = out_base + k * out_stride; order = n; info_index = k;
        }
    }
};
template <typename IO> struct IoTraits { static constexpr bool PADDED = false; };
template <typename T> struct IoTraits<PadIO<T>> { static constexpr bool PADDED = true; };

// Inputs of the fused GP kernels (reference include/gauss_cpu.h:16-58): dense, contiguous.
template <typename T>
struct GpIO {
    const T *a;   // batch x n
    const T *b;   // batch x n x n
    const T *c;   // batch x n      (diagonal added on load)
    const T *d;   // batch x n      (mean right-hand side; may be null)
    const T *e;   // batch          (variance offset;      may be null)
    T *means;     // batch          (may be null)
    T *variances; // batch          (may be null)
};

template <typename T> __device__ __forceinline__ T dev_sqrt(T x);
template <> __device__ __forceinline__ float dev_sqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double dev_sqrt<double>(double x) { return sqrt(x); }

template <typename T> __device__ __forceinline__ T dev_abs(T x);
template <> __device__ __forceinline__ float dev_abs<float>(float x) { return fabsf(x); }
template <> __device__ __forceinline__ double dev_abs<double>(double x) { return fabs(x); }

template <typename T> __device__ __forceinline__ T dev_nan();
template <> __device__ __forceinline__ float dev_nan<float>() { return __int_as_float(0x7fc00000); }
template <> __device__ __forceinline__ double dev_nan<double>() { return __longlong_as_double(0x7ff8000000000000LL); }

// Group = the threads that cooperate on one matrix: a warp (G == 32) or the whole CTA.
template <int G>
struct Group {
    static __device__ __forceinline__ void sync() {
        if (G == 32) __syncwarp(); else __syncthreads();
    }
};

__host__ __device__ __forceinline__ int packed_row(int i) { return (i * (i + 1)) >> 1; }

}  // namespace invgpu

// compile-only probe for the rolled sweep kernel
#include "../cuda_matrix_inversion_b200/csrc/generic_smem.cuh"
#include "../cuda_matrix_inversion_b200/csrc/sweep_kernels.cuh"
using namespace invgpu;
#ifndef PT
#define PT float
#endif
#ifndef PN
#define PN 128
#endif
#ifndef PP
#define PP 16
#endif
#ifndef PMINB
#define PMINB 2
#endif
#ifndef PUNROLLED
template __global__ void invgpu::sweep_rolled_kernel<PT, PN, PP, StridedIO<PT>, PMINB>(StridedIO<PT>, i64, int *);
#endif
#ifdef PUNROLLED
template __global__ void invgpu::sweep_unrolled_kernel<PT, 32, 4, 4, StridedIO<PT>, 5>(StridedIO<PT>, i64, int *);
#endif

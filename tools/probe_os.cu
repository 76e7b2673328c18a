// compile-only probe for the one-sweep kernels
#include "../cuda_matrix_inversion_b200/csrc/generic_smem.cuh"
#include "../cuda_matrix_inversion_b200/csrc/onesweep_kernels.cuh"
using namespace invgpu;
#ifndef PT
#define PT float
#endif
#ifndef PN
#define PN 32
#endif
#ifndef PTR
#define PTR 4
#endif
#ifndef PTC
#define PTC 4
#endif
#ifndef PMINB
#define PMINB 5
#endif
#ifdef PROLLED
template __global__ void invgpu::onesweep_rolled_kernel<PT, PN, PTR, StridedIO<PT>, PMINB>(StridedIO<PT>, i64, int *);
#else
template __global__ void invgpu::onesweep_spd_kernel<PT, PN, PTR, PTC, false, StridedIO<PT>, PMINB>(StridedIO<PT>, i64, int *);
#endif

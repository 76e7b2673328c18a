// compile-only probe for the sweep kernels
#include "../cuda_matrix_inversion_b200/csrc/generic_smem.cuh"
#include "../cuda_matrix_inversion_b200/csrc/sweep_kernels.cuh"
using namespace invgpu;
#ifndef PT
#define PT float
#endif
#ifndef PN
#define PN 32
#endif
#ifndef PTR
#define PTR 2
#endif
#ifndef PTC
#define PTC 4
#endif
#ifndef PUNROLL
#define PUNROLL false
#endif
#ifndef PMINB
#define PMINB 3
#endif
#ifndef PBLK
#define PBLK 1
#endif
#ifndef PBLK
#define PBLK 1
#endif
#ifdef PGP
template __global__ void invgpu::sweep_gp_kernel<PT, PN, PTR, PTC, PUNROLL, PMINB, PBLK>(GpIO<PT>, i64, int *, PT *);
#else
template __global__ void invgpu::sweep_spd_kernel<PT, PN, PTR, PTC, PUNROLL, StridedIO<PT>, PMINB, PBLK>(StridedIO<PT>, i64, int *);
#endif

// compile-only probe: one instantiation, for quick `ptxas -v` / SASS feedback
#include "../cuda_matrix_inversion_b200/csrc/generic_smem.cuh"
#include "../cuda_matrix_inversion_b200/csrc/tile_kernels.cuh"
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
#ifndef PPERM
#define PPERM true
#endif
#ifndef PSTAGES
#define PSTAGES SPD_INVERSE
#endif
#ifndef PMINB
#define PMINB 4
#endif
template __global__ void invgpu::tile_spd_kernel<PT, PN, PTR, PTC, PPERM, StridedIO<PT>, PSTAGES, PMINB>(StridedIO<PT>, i64, int *);

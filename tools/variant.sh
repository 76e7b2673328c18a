#!/bin/bash
# build a variant of the fp32 tile TU with extra -D flags into lib/variants/<name>/libinvgpu.so
# usage: tools/variant.sh name "-DINVGPU_LOCKSTEP=0 ..."
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2
out=cuda_matrix_inversion_b200/lib/variants/$name
mkdir -p $out
L=cuda_matrix_inversion_b200/lib
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $flags -c cuda_matrix_inversion_b200/csrc/inst_spd_f32.cu -o $out/inst_spd_f32.o &
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $flags -c cuda_matrix_inversion_b200/csrc/capi.cu -o $out/capi.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libinvgpu.so $out/capi.o $out/inst_spd_f32.o $L/inst_spd_f64.o $L/inst_spd_factor.o $L/mats_io.o -cudart static
echo built $out

#!/bin/bash
# build a variant of the fp32 SPD tile TU (and capi.cu, whose dispatch names the template arguments)
# with extra -D flags into lib/variants/<name>/libinvgpu.so:   tools/variant.sh name "-DINVGPU_LOCKSTEP=0 ..."
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2
L=cuda_matrix_inversion_b200/lib
out=$L/variants/$name
mkdir -p $out
NV="nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $flags"
$NV -c cuda_matrix_inversion_b200/csrc/inst_spd_f32.cu -o $out/inst_spd_f32.o &
$NV -c cuda_matrix_inversion_b200/csrc/inst_onesweep_f32.cu -o $out/inst_onesweep_f32.o &
$NV -c cuda_matrix_inversion_b200/csrc/capi.cu -o $out/capi.o &
wait
others=$(ls $L/inst_*.o | grep -v -e inst_spd_f32.o -e inst_onesweep_f32.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libinvgpu.so $out/capi.o $out/inst_spd_f32.o $out/inst_onesweep_f32.o $others $L/mats_io.o -cudart static
echo built $out

#!/bin/bash
# build a variant of the sweep TUs (and capi.cu, whose dispatch names the template arguments) with extra -D
# flags into lib/variants/<name>/libinvgpu.so:   tools/variant.sh name "-DINVGPU_SWEEP_LOCKSTEP=0 ..." [tu ...]
# use it with  INVGPU_LIB=cuda_matrix_inversion_b200/lib/variants/<name>/libinvgpu.so python tools/kbench.py ...
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2; shift 2
tus=${@:-inst_sweep_f32 inst_sweep_f64 capi}
L=cuda_matrix_inversion_b200/lib
out=$L/variants/$name
mkdir -p $out
NV="nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $flags"
objs=""
for t in $tus; do $NV -c cuda_matrix_inversion_b200/csrc/$t.cu -o $out/$t.o & objs="$objs $out/$t.o"; done
wait
others=""
for o in $L/inst_*.o $L/capi.o; do b=$(basename $o .o); case " $tus " in *" $b "*) ;; *) others="$others $o";; esac; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libinvgpu.so $objs $others $L/mats_io.o -cudart static
echo built $out

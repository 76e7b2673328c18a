#!/bin/bash
mkdir -p gpurun_out
V=cuda_matrix_inversion_b200/lib/variants
{
for name in A B C; do for n in 16 32; do
  echo "== variant $name $n f32"; INVGPU_LIB=$V/$name/libinvgpu.so timeout 120 python tools/kbench.py spd $n f32
done; done
} > gpurun_out/g_kbench.log 2>&1
grep -E "==|ms" gpurun_out/g_kbench.log | sed 's/"op": "spd", //; s/"units_per_s.*"hbm_frac"/"hbm_frac"/; s/"tier.*"info_nonzero"/"info_nz"/'

#!/bin/bash
mkdir -p gpurun_out
{
for v in 0 2 3; do for n in 64 128; do echo "== sweep v$v $n f32"; INVGPU_SWEEP_VARIANT=$v timeout 120 python tools/kbench.py spd $n f32; done; done
} > gpurun_out/n_kbench.log 2>&1
grep -E "==|ms" gpurun_out/n_kbench.log | sed 's/"op": "[a-z]*", //; s/"units_per_s.*"hbm_frac"/"hbm_frac"/; s/"tier.*"info_nonzero"/"info_nz"/'

#!/bin/bash
mkdir -p gpurun_out
{
for n in 64 128; do echo "== gj tile $n f32"; timeout 120 python tools/kbench.py general $n f32; done
for n in 32 64; do echo "== gj tile $n f64"; timeout 120 python tools/kbench.py general $n f64; done
} > gpurun_out/v_kbench.log 2>&1
grep -E "==|ms|rror" gpurun_out/v_kbench.log | sed 's/"op": "[a-z]*", //; s/"units_per_s.*"hbm_frac"/"hbm_frac"/; s/"tier.*"info_nonzero"/"info_nz"/'

#!/bin/bash
mkdir -p gpurun_out
{
for n in 16 32 64 128; do for v in 0 4 5; do echo "== spd v$v $n f32"; INVGPU_SWEEP_VARIANT=$v timeout 120 python tools/kbench.py spd $n f32; done; done
for n in 32 64 128; do for v in 0 4; do echo "== spd v$v $n f64"; INVGPU_SWEEP_VARIANT=$v timeout 120 python tools/kbench.py spd $n f64; done; done
for v in 0 4 5; do echo "== gp v$v 128 f32"; INVGPU_SWEEP_VARIANT=$v timeout 120 python tools/kbench.py gp 128 f32 25000; done
echo "== gp v4 64 f32"; INVGPU_SWEEP_VARIANT=4 timeout 120 python tools/kbench.py gp 64 f32
} > gpurun_out/o_kbench.log 2>&1
grep -E "==|ms|rror" gpurun_out/o_kbench.log | sed 's/"op": "[a-z]*", //; s/"units_per_s.*"hbm_frac"/"hbm_frac"/; s/"tier.*"info_nonzero"/"info_nz"/'

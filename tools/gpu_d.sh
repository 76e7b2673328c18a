#!/bin/bash
mkdir -p gpurun_out
{
for v in 0 1 2 3; do echo "== sweep v$v 32 f32"; INVGPU_SWEEP_VARIANT=$v timeout 120 python tools/kbench.py spd 32 f32; done
echo "== onesweep 32 f32"; INVGPU_SPD_KERNEL=onesweep timeout 120 python tools/kbench.py spd 32 f32
for v in 0 1; do for n in 16 64 128; do echo "== sweep v$v $n f32"; INVGPU_SWEEP_VARIANT=$v timeout 120 python tools/kbench.py spd $n f32; done; done
echo "== onesweep 16 f32"; INVGPU_SPD_KERNEL=onesweep timeout 120 python tools/kbench.py spd 16 f32
for n in 16 32 64 128; do echo "== sweep $n f64"; timeout 120 python tools/kbench.py spd $n f64; done
for n in 16 32; do echo "== onesweep $n f64"; INVGPU_SPD_KERNEL=onesweep timeout 120 python tools/kbench.py spd $n f64; done
} > gpurun_out/d_kbench.log 2>&1
grep -E "==|ms" gpurun_out/d_kbench.log | sed 's/"op": "spd", //; s/"units_per_s.*"hbm_frac"/"hbm_frac"/; s/"tier.*"info_nonzero"/"info_nz"/'

#!/bin/bash
mkdir -p gpurun_out
{
for v in 0 6 7; do echo "== spd v$v 32 f32"; INVGPU_SWEEP_VARIANT=$v timeout 120 python tools/kbench.py spd 32 f32; done
echo "== spd v6 32 f32 odd batch"; INVGPU_SWEEP_VARIANT=6 timeout 120 python tools/kbench.py spd 32 f32 100003
} > gpurun_out/p_kbench.log 2>&1
grep -E "==|ms|rror|rap" gpurun_out/p_kbench.log | sed 's/"op": "[a-z]*", //; s/"units_per_s.*"hbm_frac"/"hbm_frac"/; s/"tier.*"info_nonzero"/"info_nz"/'
tail -n 5 gpurun_out/p_kbench.log

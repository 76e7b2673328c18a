#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; tail -3 gpurun_out/f_pytest.log
{
for v in 0 1 2 3; do echo "== sweep v$v 32 f32"; INVGPU_SWEEP_VARIANT=$v timeout 120 python tools/kbench.py spd 32 f32; done
for v in 0 1; do for n in 16; do echo "== sweep v$v $n f32"; INVGPU_SWEEP_VARIANT=$v timeout 120 python tools/kbench.py spd $n f32; done; done
for n in 16 32; do echo "== sweep $n f64"; timeout 120 python tools/kbench.py spd $n f64; done
} > gpurun_out/f_kbench.log 2>&1
grep -E "==|ms" gpurun_out/f_kbench.log | sed 's/"op": "spd", //; s/"units_per_s.*"hbm_frac"/"hbm_frac"/; s/"tier.*"info_nonzero"/"info_nz"/'

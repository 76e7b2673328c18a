#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest.log 2>&1; tail -n 12 gpurun_out/h_pytest.log
{
for n in 32 64 128; do
  echo "== gp sweep $n f32"; timeout 120 python tools/kbench.py gp $n f32 $([ $n = 128 ] && echo 25000)
  echo "== gp tile $n f32"; INVGPU_GP_KERNEL=tile timeout 120 python tools/kbench.py gp $n f32 $([ $n = 128 ] && echo 25000)
done
for n in 32 64 128; do
  echo "== gp sweep $n f64"; timeout 120 python tools/kbench.py gp $n f64 $([ $n = 128 ] && echo 25000)
  echo "== gp tile $n f64"; INVGPU_GP_KERNEL=tile timeout 120 python tools/kbench.py gp $n f64 $([ $n = 128 ] && echo 25000)
done
for n in 16 32 64 128; do echo "== spd sweep $n f32"; timeout 120 python tools/kbench.py spd $n f32; done
} > gpurun_out/h_kbench.log 2>&1
grep -E "==|ms" gpurun_out/h_kbench.log | sed 's/"op": "[a-z]*", //; s/"units_per_s.*"hbm_frac"/"hbm_frac"/; s/"tier.*"info_nonzero"/"info_nz"/'

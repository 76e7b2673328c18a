#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "general or legacy or smoke" > gpurun_out/t_pytest.log 2>&1; tail -n 15 gpurun_out/t_pytest.log
{
for n in 16 32 64 128; do
  echo "== gj tile $n f32"; timeout 120 python tools/kbench.py general $n f32
  echo "== gj old $n f32"; INVGPU_GJ_KERNEL=rowlane timeout 120 python tools/kbench.py general $n f32
done
for n in 32 64; do echo "== gj tile $n f64"; timeout 120 python tools/kbench.py general $n f64; echo "== gj old $n f64"; INVGPU_GJ_KERNEL=rowlane timeout 120 python tools/kbench.py general $n f64; done
} > gpurun_out/t_kbench.log 2>&1
grep -E "==|ms|rror" gpurun_out/t_kbench.log | sed 's/"op": "[a-z]*", //; s/"units_per_s.*"hbm_frac"/"hbm_frac"/; s/"tier.*"info_nonzero"/"info_nz"/'

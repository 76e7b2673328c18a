#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; tail -5 gpurun_out/c_pytest.log
{
for cfg in "64 f32" "128 f32" "64 f64" "128 f64"; do
  echo "== sweep $cfg"; timeout 120 python tools/kbench.py spd $cfg
done
echo "== rolled 32 f32"; INVGPU_SPD_KERNEL=rolled timeout 120 python tools/kbench.py spd 32 f32
} > gpurun_out/c_kbench.log 2>&1
cat gpurun_out/c_kbench.log

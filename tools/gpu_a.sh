#!/bin/bash
# rolled one-sweep vs default, plus a full ncu capture of the default n=32 one-sweep kernel
mkdir -p gpurun_out
{
for k in default rolled; do
  for cfg in "32 f32" "64 f32" "128 f32" "64 f64" "128 f64"; do
    echo "== $k $cfg"; INVGPU_SPD_KERNEL=$k timeout 120 python tools/kbench.py spd $cfg
  done
done
} > gpurun_out/a_kbench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:onesweep_spd -s 2 -c 1 -o gpurun_out/os32_r1 \
    python tools/kbench.py spd 32 f32 262144 > gpurun_out/a_ncu.log 2>&1
cat gpurun_out/a_kbench.log

#!/bin/bash
mkdir -p gpurun_out
tools/bin/microbench > gpurun_out/microbench.log 2>&1
{
for cfg in "64 f32" "128 f32"; do
  echo "== rolled $cfg"; INVGPU_SPD_KERNEL=rolled timeout 120 python tools/kbench.py spd $cfg
done
echo "== gp 128 f32"; timeout 120 python tools/kbench.py gp 128 f32 25000
echo "== gp 64 f32"; timeout 120 python tools/kbench.py gp 64 f32 
} > gpurun_out/b_kbench.log 2>&1
INVGPU_SPD_KERNEL=rolled timeout 300 ncu --set full --clock-control none --import-source on -k regex:onesweep_rolled -s 2 -c 1 -o gpurun_out/osr128_r1 \
    python tools/kbench.py spd 128 f32 16384 > gpurun_out/b_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tile_gp -s 2 -c 1 -o gpurun_out/gp128_r1 \
    python tools/kbench.py gp 128 f32 25000 > gpurun_out/b_ncu2.log 2>&1
cat gpurun_out/microbench.log gpurun_out/b_kbench.log

#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_spd -s 2 -c 1 -o gpurun_out/sw32_r1 \
    python tools/kbench.py spd 32 f32 262144 > gpurun_out/e_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_spd -s 2 -c 1 -o gpurun_out/sw128_r1 \
    python tools/kbench.py spd 128 f32 16384 > gpurun_out/e_ncu2.log 2>&1
tail -3 gpurun_out/e_ncu1.log gpurun_out/e_ncu2.log

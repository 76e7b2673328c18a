#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_gp -s 2 -c 1 -o gpurun_out/swgp128_r1 \
    python tools/kbench.py gp 128 f32 25000 > gpurun_out/i_ncu1.log 2>&1
tail -n 3 gpurun_out/i_ncu1.log

#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_spd -s 2 -c 1 -o gpurun_out/sw64_r1 \
    python tools/kbench.py spd 64 f32 65536 > gpurun_out/m_ncu1.log 2>&1
tail -n 2 gpurun_out/m_ncu1.log

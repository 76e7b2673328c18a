#!/bin/bash
mkdir -p gpurun_out
{
for v in 6 7; do echo "== spd v$v 32 f32"; INVGPU_TRACE=1 INVGPU_SWEEP_VARIANT=$v timeout 120 python tools/kbench.py spd 32 f32 262144; done
} > gpurun_out/q_kbench.log 2>&1
grep -E "==|ms|rror|rap|invgpu" gpurun_out/q_kbench.log | sed 's/"op": "[a-z]*", //; s/"units_per_s.*"hbm_frac"/"hbm_frac"/; s/"tier.*"info_nonzero"/"info_nz"/' | sort | uniq -c
INVGPU_SWEEP_VARIANT=6 timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_spd -s 2 -c 1 -o gpurun_out/sw32tma_r1 \
    python tools/kbench.py spd 32 f32 262144 > gpurun_out/q_ncu.log 2>&1
tail -n 2 gpurun_out/q_ncu.log

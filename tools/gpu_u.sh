#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "general or legacy" > gpurun_out/u_pytest.log 2>&1; tail -n 6 gpurun_out/u_pytest.log
{
for n in 64 128; do echo "== gj tile $n f32"; timeout 120 python tools/kbench.py general $n f32; done
for n in 32 64 128; do echo "== gj tile $n f64"; timeout 120 python tools/kbench.py general $n f64; done
} > gpurun_out/u_kbench.log 2>&1
grep -E "==|ms|rror" gpurun_out/u_kbench.log | sed 's/"op": "[a-z]*", //; s/"units_per_s.*"hbm_frac"/"hbm_frac"/; s/"tier.*"info_nonzero"/"info_nz"/'
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gj_tile -s 2 -c 1 -o gpurun_out/gjt64_r1 \
    python tools/kbench.py general 64 f32 65536 > gpurun_out/u_ncu.log 2>&1
tail -n 2 gpurun_out/u_ncu.log

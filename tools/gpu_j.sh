#!/bin/bash
mkdir -p gpurun_out
INVGPU_MIXED_TIMING=1 timeout 600 python tools/mixed_bench.py > gpurun_out/j_mixed.log 2>&1
tail -n 12 gpurun_out/j_mixed.log
timeout 600 python tools/mixed_bench.py 2>&1 | tail -n 2

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; tail -n 6 gpurun_out/k_pytest.log
INVGPU_MIXED_TIMING=1 timeout 600 python tools/mixed_bench.py > gpurun_out/k_mixed.log 2>&1
tail -n 8 gpurun_out/k_mixed.log
timeout 600 python tools/mixed_bench.py 2>&1 | tail -n 2

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r_pytest.log 2>&1; tail -n 8 gpurun_out/r_pytest.log
{
echo "== gp 128 f32"; timeout 120 python tools/kbench.py gp 128 f32 25000
echo "== mixed"; timeout 300 python tools/mixed_bench.py
} 2>&1 | grep -E "==|ms" | cut -c1-300

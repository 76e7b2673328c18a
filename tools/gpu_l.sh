#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/l_pytest.log 2>&1; tail -n 8 gpurun_out/l_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 1200 python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -n 3 gpurun_out/bench_r1c.err
bash tools/profile.sh r1c

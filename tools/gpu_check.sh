#!/bin/bash
# Run ON THE GPU BOX (through gpurun): the whole round-end sequence -- GPU parity tests, smoke, the bench
# line (with extras), then tools/profile.sh (plain bench, ncu launch list of the same command, one
# `ncu --set full` capture of the headline kernel).  Outputs land in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash tools/gpu_check.sh r1'
tag=${1:-r1}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${tag}.log 2>&1; tail -n 4 gpurun_out/pytest_${tag}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 1500 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; tail -n 3 gpurun_out/bench_${tag}.err
bash tools/profile.sh ${tag}

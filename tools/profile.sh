#!/bin/bash
# Run ON THE GPU BOX (through gpurun): plain bench run, then the ncu launch list of the same command,
# then one `--set full` capture of the headline kernel.  Outputs land in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/profile.sh r1b'
set -u
tag=${1:-r1}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/bench_${tag}_plain.json 2> gpurun_out/bench_${tag}_plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sweep_|tile_|gj_|onesweep_|_generic_|mixed_" -c 40 --csv \
    --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 10 --warmup 3 --no-extra > gpurun_out/ncu_launches_${tag}.log 2>&1
timeout 300 python tools/kbench.py spd 32 f32 262144 > gpurun_out/kbench_${tag}.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_spd -s 2 -c 1 -o gpurun_out/sweep32_${tag} \
    python tools/kbench.py spd 32 f32 262144 > gpurun_out/ncu_full_${tag}.log 2>&1
tail -n 2 gpurun_out/ncu_full_${tag}.log
cut -c1-600 gpurun_out/bench_${tag}_plain.json

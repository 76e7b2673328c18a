#!/bin/bash
# usage: tools/probe.sh [-DPT=float -DPN=32 -DPTR=4 -DPTC=2 -DPMINB=3 ...]
#   -> registers, spills, SASS opcode histogram of the hot path (up to the final EXIT)
set -e
cd "$(dirname "$0")/.."
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xptxas -v "$@" -c ${PROBE_SRC:-tools/probe.cu} -o /tmp/probe.o 2>&1 | grep -E "registers|spill|error" || true
cuobjdump -sass /tmp/probe.o > /tmp/probe.sass
last=$(grep -nE '^\s+/\*[0-9a-f]{4,6}\*/\s+EXIT' /tmp/probe.sass | tail -1 | cut -d: -f1)
head -n "$last" /tmp/probe.sass > /tmp/probe_hot.sass
echo "hot-path instructions: $(grep -cE '^\s+/\*[0-9a-f]{4,6}\*/' /tmp/probe_hot.sass)  (whole function: $(grep -cE '^\s+/\*[0-9a-f]{4,6}\*/' /tmp/probe.sass))"
grep -oE '^\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P[0-9] )?[A-Z0-9_.]+' /tmp/probe_hot.sass | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${TOPN:-18} | tr '\n' ';'; echo

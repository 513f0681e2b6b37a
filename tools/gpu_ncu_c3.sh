#!/bin/bash
# ncu capture of the C3 (84-aa windows vs PR/RT/INT) forward kernel
set -u
mkdir -p gpurun_out
CMD="python tools/bench_configs.py --c2 0 --c4 0 --c3 300000 --steps 1 --verify 10"
$CMD > gpurun_out/c3_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_forward -s 1 -c 1 -f -o gpurun_out/prof_c3 $CMD > gpurun_out/c3_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/c3_ncu.log

#!/bin/bash
# ncu --set full capture of the long-pair forward kernel (strip dataflow) on a C4-shaped batch, after a plain run.
set -u
mkdir -p gpurun_out
CMD="python tools/bench_configs.py --c2 0 --c3 0 --c4 ${PAIRS:-300} --steps 1 --verify 2"
GOTOH_B200_LONG=flow $CMD > gpurun_out/c4_plain.log 2>&1 && \
GOTOH_B200_LONG=flow ncu --set full --clock-control none --import-source on -k regex:k_forward_flow -s 1 -c 1 -f -o gpurun_out/prof_c4_flow $CMD > gpurun_out/c4_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/c4_ncu.log

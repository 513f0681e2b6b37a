#!/bin/bash
# ncu launch list + one --set full capture of the forward kernel (after a plain run of the same command).
set -u
mkdir -p gpurun_out
CMD="python bench.py --pairs 60000 --steps 1 --warmup 1 --verify 0 --lite"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"; tail -2 gpurun_out/plain.log
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_forward -s 1 -c 1 -o gpurun_out/prof_forward $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log

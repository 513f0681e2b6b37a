#!/bin/bash
# ncu launch list + one --set full capture of a config's forward kernel (each after a plain run of the same command).
#   bash tools/gpu_ncu.sh [config] [pairs] [tag]
set -u
CFG=${1:-c2}; PAIRS=${2:-60000}; TAG=${3:-$CFG}
mkdir -p gpurun_out
CMD="python bench.py --config $CFG --pairs $PAIRS --steps 1 --warmup 1 --verify 0 --lite"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"; tail -2 gpurun_out/plain_$TAG.log
$CMD > gpurun_out/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_forward -s 1 -c 1 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full_$TAG.log

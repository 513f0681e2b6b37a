#!/bin/bash
# end-of-round pass on one GPU: parity, smoke, bench, configs, gotoh2 bench, ncu launch lists + --set full captures
set -u
mkdir -p gpurun_out
bash tools/gpu_check.sh
bash tools/gpu_round.sh
NCU=0 bash tools/gpu_gotoh2.sh > gpurun_out/gotoh2_run.log 2>&1; tail -c 300 gpurun_out/gotoh2_run.log
# launch list of the default bench command (resident arm first: 1 warm-up + 2 steps)
python bench.py --steps 2 --warmup 1 --verify 20 > gpurun_out/plain_default.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 2 --warmup 1 --verify 20 > gpurun_out/ncu_default.log 2>&1
echo "ncu default launches rc=$?"
bash tools/gpu_ncu.sh
bash tools/gpu_ncu_c3.sh

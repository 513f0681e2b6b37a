#!/bin/bash
# configs C2b/C3/C4 (full 10k pairs, default long-pair kernel) on one GPU; logs to gpurun_out/
set -u
mkdir -p gpurun_out
timeout 1200 python tools/bench_configs.py --c4 ${C4:-10000} --c4-modes flow --steps 3 > gpurun_out/configs.log 2> gpurun_out/configs.err; echo "configs rc=$?"
cut -c1-420 gpurun_out/configs.log; tail -3 gpurun_out/configs.err

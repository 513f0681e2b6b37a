#!/bin/bash
# end-of-round pass on one GPU: parity, smoke, every bench config, latency table, ncu launch lists + --set full captures
set -u
mkdir -p gpurun_out
BENCH_ARGS="--steps 5 --warmup 3" bash tools/gpu_check.sh
bash tools/gpu_bench_cfg.sh "c2b c3 c1" --steps 5 --warmup 3
bash tools/gpu_bench_cfg.sh "c4" --steps 3 --warmup 3
bash tools/gpu_bench_cfg.sh "c5" --steps 2 --warmup 1
python tools/bench_latency.py --reps 7 > gpurun_out/latency.json 2> gpurun_out/latency.err; echo "latency rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "reference arm rc=$?"; tail -c 300 gpurun_out/bench_reference.log
bash tools/gpu_ncu.sh c2 60000 c2
bash tools/gpu_ncu.sh c3 300000 c3
bash tools/gpu_ncu.sh c4 600 c4

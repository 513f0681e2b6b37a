#!/bin/bash
# One GPU-box pass: parity tests, smoke, a short bench.  Logs go to gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout ${TEST_TIMEOUT:-900} python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout ${BENCH_TIMEOUT:-600} python bench.py --pairs ${PAIRS:-1000000} --steps ${STEPS:-5} --warmup ${WARMUP:-3} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/bench.err
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -c 3000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err

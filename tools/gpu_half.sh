#!/bin/bash
# half-warp wavefronts: parity on the GPU, then C3 with and without them
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "half_warp or fuzz or c3 or golden or drop_in" > gpurun_out/pytest_half.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_half.log
timeout 600 python tools/bench_configs.py --c2 0 --c4 0 --steps 3 2>&1 | cut -c1-330
GOTOH_B200_HALF=0 timeout 600 python tools/bench_configs.py --c2 0 --c4 0 --steps 3 2>&1 | cut -c1-330

mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long or hcv or c4 or golden" 2>&1 | tail -3
timeout 900 python tools/bench_configs.py --c2 0 --c3 0 --c4 10000 --c4-modes flow --steps 3 2>&1 | cut -c1-400
timeout 900 python tools/bench_configs.py --c2 0 --c3 0 --c4 600 --c4-modes flow --steps 3 2>&1 | cut -c1-400

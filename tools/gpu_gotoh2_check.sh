set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gotoh2.py tests/test_callers.py -m gpu -x -q > gpurun_out/pytest_gotoh2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gotoh2.log
bash tools/gpu_gotoh2.sh

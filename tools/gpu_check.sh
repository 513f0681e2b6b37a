#!/bin/bash
# GPU parity suite + smoke + default bench on one GPU; logs to gpurun_out/
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > gpurun_out/gpu.txt 2>&1
( time timeout ${TEST_TIMEOUT:-1500} python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
( time timeout 900 python bench.py ${BENCH_ARGS:-} ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
    e = d["e2e"]
    print(d["config"]["name"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", round(e["value"]), round(e["s_per_step"], 4),
          "compact", round(e["compact"]["value"]), round(e["compact"]["s_per_step"], 4), "ceiling", e.get("host_ceiling"), "roof", round(d["roofline"]["frac"], 3), d["clocks"], d["bit_exact_verified_pairs"])
except Exception as ex:
    print("no bench line:", ex)
PY

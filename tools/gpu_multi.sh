#!/bin/bash
# bench.py under torchrun on N GPUs of one box + the in-library multi-device test: bash tools/gpu_multi.sh N [config] [extra bench args]
set -u
N=${1:-2}; CFG=${2:-c2}; shift; shift || true
mkdir -p gpurun_out
nproc > gpurun_out/nproc_${N}gpu.txt; nvidia-smi topo -m >> gpurun_out/nproc_${N}gpu.txt 2>&1
python -m pytest tests/test_gpu_scale.py -m gpu -q -k multi_device -rs > gpurun_out/pytest_multi_${N}gpu.log 2>&1; echo "multi-device test rc=$?"; tail -3 gpurun_out/pytest_multi_${N}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $CFG "$@" > gpurun_out/bench_${CFG}_${N}gpu.log 2> gpurun_out/bench_${CFG}_${N}gpu.err
echo "bench rc=$?"; tail -c 400 gpurun_out/bench_${CFG}_${N}gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${CFG}_${N}gpu.log").read().strip().splitlines()[-1])
e = d["e2e"]
print(d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "| e2e strings", round(e["value"]), round(e["s_per_step"], 4), "d2h GB/s", round(e["d2h_gbs"], 1), "ceiling", e.get("host_ceiling"),
      "| compact", round(e["compact"]["value"]), round(e["compact"]["s_per_step"], 4), "|", d["clocks"], d["bit_exact_verified_pairs"], "cpu", d["cpu_baseline"]["cores"])
PY

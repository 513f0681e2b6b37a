#!/bin/bash
# bench.py under torchrun on N GPUs of one box: bash tools/gpu_multi.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err
echo "bench rc=$?"; tail -c 400 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${N}gpu.log").read().strip().splitlines()[-1])
print(d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["s_per_step"], d["clocks"])
PY

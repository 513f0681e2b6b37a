#!/bin/bash
# bench.py under torchrun on N GPUs of one box (+ the in-library multi-device test): bash tools/gpu_multi.sh N "c2 c5" [extra bench args]
set -u
N=${1:-2}; CFGS=${2:-c2}; shift; shift || true
mkdir -p gpurun_out
nproc > gpurun_out/nproc_${N}gpu.txt; nvidia-smi topo -m >> gpurun_out/nproc_${N}gpu.txt 2>&1
python -m pytest tests/test_gpu_scale.py -m gpu -q -k multi_device -rs > gpurun_out/pytest_multi_${N}gpu.log 2>&1; echo "multi-device test rc=$?"; tail -2 gpurun_out/pytest_multi_${N}gpu.log
for CFG in $CFGS; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $CFG "$@" > gpurun_out/bench_${CFG}_${N}gpu.log 2> gpurun_out/bench_${CFG}_${N}gpu.err
echo "bench $CFG rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_${CFG}_${N}gpu.err | tail -3
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${CFG}_${N}gpu.log").read().strip().splitlines()[-1])
    e = d["e2e"]
    print(d["config"]["name"], d["n_gpus"], "gpus value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "aln/s %.3g" % d["alignments_per_s"], "| e2e strings", round(e["value"]), round(e["s_per_step"], 4), "d2h GB/s", round(e["d2h_gbs"], 1),
          "ceiling", round(e["host_ceiling"]["gbs"], 1), "frac", round(e["host_ceiling"]["d2h_frac_of_ceiling"], 3), "| compact", round(e["compact"]["value"]), round(e["compact"]["s_per_step"], 4), "aln/s %.3g" % e["compact"]["alignments_per_s"], "|", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d["bit_exact_verified_pairs"], "cpu cores", d["cpu_baseline"]["cores"])
except Exception as ex:
    print("no bench line:", ex)
PY
done

#!/bin/bash
# bench.py for a list of configs on one GPU: bash tools/gpu_bench_cfg.sh "c2 c3 c4" [extra bench args]; logs gpurun_out/bench_<cfg>.log
set -u
mkdir -p gpurun_out
CFGS=${1:-"c2 c3 c4"}; shift || true
for c in $CFGS; do
  ( time timeout 1500 python bench.py --config $c "$@" ) > gpurun_out/bench_$c.log 2> gpurun_out/bench_$c.err; echo "bench $c rc=$?"; tail -c 400 gpurun_out/bench_$c.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$c.log").read().strip().splitlines()[-1])
    e = d["e2e"]
    print(d["config"]["name"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "aln/s", round(d["alignments_per_s"]), "| e2e", round(e["value"]), round(e["s_per_step"], 4),
          "| compact", round(e["compact"]["value"]), round(e["compact"]["s_per_step"], 4), "| ceiling", (e.get("host_ceiling") or {}).get("gbs"), "d2h_frac", (e.get("host_ceiling") or {}).get("d2h_frac_of_ceiling"),
          "| roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["achieved"]), "/", round(d["roofline"]["peak"]), "| cpu", round(d["cpu_baseline"]["value"], 3), d["cpu_baseline"]["cores"], "|", d["clocks"], d["bit_exact_verified_pairs"])
except Exception as ex:
    print("no bench line:", ex)
PY
done

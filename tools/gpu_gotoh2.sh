#!/bin/bash
# gotoh2 (live aligner) pass: bench on the three caller shapes, then the ncu launch list and one --set full
# capture each of the forward and reverse kernels on a small run of the same tool.
set -u
mkdir -p gpurun_out
timeout ${BENCH_TIMEOUT:-900} python tools/bench_gotoh2.py ${BENCH_ARGS:-} > gpurun_out/gotoh2_bench.jsonl 2> gpurun_out/gotoh2_bench.err; echo "bench_gotoh2 rc=$?"
tail -c 6000 gpurun_out/gotoh2_bench.jsonl; tail -5 gpurun_out/gotoh2_bench.err
if [ "${NCU:-1}" = "1" ]; then
CMD="python tools/bench_gotoh2.py --remap 0 --aa 0 --reads 4000 --steps 1 --cpu-seconds 0"
$CMD > gpurun_out/gotoh2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/gotoh2_launches.csv $CMD > gpurun_out/gotoh2_ncu_launches.log 2>&1
echo "ncu launches rc=$?"; tail -2 gpurun_out/gotoh2_plain.log
$CMD > gpurun_out/gotoh2_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'^k2f_x2$|^k2f$|^k2r$' -s 2 -c 2 -f -o gpurun_out/prof_gotoh2 $CMD > gpurun_out/gotoh2_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/gotoh2_ncu_full.log
fi

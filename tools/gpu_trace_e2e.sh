# e2e of the one-shot calls, strings vs compact vs tight, with the library's per-slab trace (diagnostics)
# usage: bash tools/gpu_trace_e2e.sh "c3 1000000" "c2 200000" ...
mkdir -p gpurun_out
for cfg in "$@"; do set -- $cfg
  for f in strings compact tight; do
    python tools/trace_e2e.py --config $1 --pairs $2 --reps 3 --format $f 2>/dev/null
    GOTOH_B200_TRACE=1 python tools/trace_e2e.py --config $1 --pairs $2 --reps 1 --format $f > /dev/null 2> gpurun_out/trace_$1_$f.log
  done
done

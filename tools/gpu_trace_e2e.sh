# e2e of the one-shot call under different pipeline depths (diagnostics)
mkdir -p gpurun_out
for cfg in "4 2" "6 2" "8 2" "6 3" "8 4" "8 3"; do set -- $cfg
  GOTOH_B200_WORKSPACES=$1 GOTOH_B200_BUILDERS=$2 python tools/trace_e2e.py --reps 3 2>/dev/null
done
GOTOH_B200_WORKSPACES=8 GOTOH_B200_BUILDERS=2 GOTOH_B200_SLAB_MB=6144 python tools/trace_e2e.py --reps 3 2>/dev/null
GOTOH_B200_WORKSPACES=8 GOTOH_B200_BUILDERS=4 GOTOH_B200_SLAB_MB=1536 python tools/trace_e2e.py --reps 3 2>/dev/null
GOTOH_B200_TRACE=1 python tools/trace_e2e.py --reps 2 > gpurun_out/trace_e2e.out 2> gpurun_out/trace_e2e.err
cat gpurun_out/trace_e2e.out

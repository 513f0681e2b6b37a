# e2e of the one-shot call under different pipeline settings (diagnostics)
mkdir -p gpurun_out
for d in 0 1 2 3; do GOTOH_B200_FWD_DEPTH=$d python tools/trace_e2e.py --reps 3 2>/dev/null; done
GOTOH_B200_FWD_DEPTH=1 GOTOH_B200_SLAB_MB=6144 python tools/trace_e2e.py --reps 3 2>/dev/null
GOTOH_B200_FWD_DEPTH=2 GOTOH_B200_SLAB_MB=1536 python tools/trace_e2e.py --reps 3 2>/dev/null
GOTOH_B200_TRACE=1 python tools/trace_e2e.py --reps 2 > gpurun_out/trace_e2e.out 2> gpurun_out/trace_e2e.err
cat gpurun_out/trace_e2e.out

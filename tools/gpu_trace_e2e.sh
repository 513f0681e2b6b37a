# e2e of the one-shot call under different pipeline settings (diagnostics)
mkdir -p gpurun_out
for mb in 2048 1536 1024 2560 2048; do GOTOH_B200_SLAB_MB=$mb python tools/trace_e2e.py --reps 5 2>/dev/null; done

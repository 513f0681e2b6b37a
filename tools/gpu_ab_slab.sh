# slab size A/B on C3 (pinned inputs come from bench.py; the trace tool uses pageable ones)
for mb in 1536 768 384; do for f in strings compact; do
GOTOH_B200_SLAB_MB=$mb python tools/trace_e2e.py --config c3 --pairs 1000000 --reps 5 --format $f 2>/dev/null
done; done
for b in 2 3; do GOTOH_B200_BUILDERS=$b GOTOH_B200_SLAB_MB=768 python tools/trace_e2e.py --config c3 --pairs 1000000 --reps 5 --format strings 2>/dev/null; done

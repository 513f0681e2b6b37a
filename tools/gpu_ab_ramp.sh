for r in 1 0; do for f in strings compact; do
GOTOH_B200_RAMP=$r python tools/trace_e2e.py --config c3 --pairs 1000000 --reps 5 --format $f 2>/dev/null
GOTOH_B200_RAMP=$r python tools/trace_e2e.py --config c2 --pairs 1000000 --reps 3 --format $f 2>/dev/null
done; done

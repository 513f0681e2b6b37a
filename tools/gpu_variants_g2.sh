#!/bin/bash
# A/B of library variants on the gotoh2 bench (reads + aa windows)
for v in variants/*.so; do
  cp $v micall-lite_b200/lib/libgotoh_b200.so; touch micall-lite_b200/lib/libgotoh_b200.so
  echo "== $v"; python tools/bench_gotoh2.py --remap 0 --cpu-seconds 0 --steps 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['config'][:12], 'gcups %.0f fwd %.0f (%.3f ms) rev %.0f' % (d['gcups'], d['gcups_forward'], d['ms_forward'], d['gcups_reverse']))
"
done

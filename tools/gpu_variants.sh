#!/bin/bash
# A/B kernels builds: run the lite bench with each prebuilt library variant (variants/*.so in the repo root)
for v in variants/*.so; do
  cp $v micall-lite_b200/lib/libgotoh_b200.so; touch micall-lite_b200/lib/libgotoh_b200.so
  echo "== $v"; python bench.py --pairs 400000 --steps 3 --warmup 2 --verify 50 --lite 2>&1 | tail -1 | cut -c1-200
done

#!/bin/bash
# A/B kernel builds: run the lite bench (resident arm only) with each prebuilt library variant (variants/*.so, git-ignored)
# usage: bash tools/gpu_variants.sh "c2 c3"
cp micall-lite_b200/lib/libgotoh_b200.so /tmp/libgotoh_b200.keep
for v in variants/*.so; do
  cp $v micall-lite_b200/lib/libgotoh_b200.so; touch micall-lite_b200/lib/libgotoh_b200.so
  for c in ${1:-c2}; do echo "== $v $c"; python bench.py --config $c --steps 3 --warmup 2 --verify 50 --lite 2>&1 | tail -1 | cut -c1-170; done
done
cp /tmp/libgotoh_b200.keep micall-lite_b200/lib/libgotoh_b200.so

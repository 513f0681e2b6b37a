#!/usr/bin/env python
"""Diagnostics: time the one-shot call gotoh_b200_align_batch on C2 reads (pinned host buffers) under a few
environment settings; with GOTOH_B200_TRACE=1 the library prints the per-slab host phases and GPU timeline on stderr.
    python tools/trace_e2e.py [--pairs N] [--reps R]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "micall-lite_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1000000)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import gotoh_b200
    from gotoh_b200 import packing, workloads
    from gotoh_b200.api import Aligner, PinnedArray
    al = Aligner()
    n = a.pairs
    ref, qb, qo = workloads.c2_reads_packed(n, seed=20260101)
    rb, ro = packing.pack([ref])
    ridx = np.zeros(n, np.int32)
    out_off = packing.out_offsets(ro, ridx, qo)
    pin = [PinnedArray(al, qb.shape, np.uint8), PinnedArray(al, (int(out_off[-1]),), np.uint8),
           PinnedArray(al, (int(out_off[-1]),), np.uint8), PinnedArray(al, (n,), np.int32), PinnedArray(al, (n,), np.int32)]
    pin[0].array[:] = qb
    outs = (pin[1].array, pin[2].array, pin[3].array, pin[4].array)
    cells = float(np.diff(qo).sum()) * len(ref)
    best = None
    for it in range(a.reps + 1):
        t0 = time.perf_counter()
        al.align_packed(rb, ro, ridx, pin[0].array, qo, 10, 3, 1, gotoh_b200.NT, out_off=out_off, out=outs, device_mask=1)
        dt = time.perf_counter() - t0
        sys.stderr.write("[trace_e2e] call %d: %.1f ms\n" % (it, dt * 1e3))
        if it and (best is None or dt < best):
            best = dt
    print("pairs %d  best %.1f ms  e2e %.0f GCUPS  env %s" % (n, best * 1e3, cells / best / 1e9,
          {k: v for k, v in os.environ.items() if k.startswith("GOTOH_B200")}))


if __name__ == "__main__":
    main()

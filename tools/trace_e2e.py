#!/usr/bin/env python
"""Diagnostics: time the one-shot calls on a bench config (pinned host buffers) in one of the three result forms; with
GOTOH_B200_TRACE=1 the library prints the per-slab host phases and GPU timeline on stderr.
    python tools/trace_e2e.py [--config c2] [--pairs N] [--reps R] [--format strings|tight|compact]
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
    ap.add_argument("--config", default="c2")
    ap.add_argument("--pairs", type=int, default=1000000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--format", default="strings", choices=["strings", "tight", "compact"])
    a = ap.parse_args()
    import bench
    from gotoh_b200.api import Aligner, PinnedArray
    al = Aligner()
    b = bench.make_batches(a.config, a.pairs, 0, 1)[0]
    cap = int(b.out_off[-1])
    if a.format == "compact":
        words = int(((np.minimum(b.rlen, b.qlen) * 5 // 4 + 47) >> 4).sum())
        pin = [PinnedArray(al, (b.n * 8,), np.int32), PinnedArray(al, (words,), np.uint32), PinnedArray(al, (b.n,), np.int64)]
        call = lambda: al.align_packed_compact(b.rb, b.ro, b.ridx, b.qb, b.qo, b.gip, b.gep, b.term, b.matrix, out=tuple(p.array for p in pin))
    elif a.format == "tight":
        pin = [PinnedArray(al, (cap,), np.uint8), PinnedArray(al, (cap,), np.uint8), PinnedArray(al, (b.n,), np.int64),
               PinnedArray(al, (b.n,), np.int32), PinnedArray(al, (b.n,), np.int32)]
        call = lambda: al.align_packed_tight(b.rb, b.ro, b.ridx, b.qb, b.qo, b.gip, b.gep, b.term, b.matrix, out=tuple(p.array for p in pin))
    else:
        pin = [PinnedArray(al, (cap,), np.uint8), PinnedArray(al, (cap,), np.uint8), PinnedArray(al, (b.n,), np.int32), PinnedArray(al, (b.n,), np.int32)]
        call = lambda: al.align_packed(b.rb, b.ro, b.ridx, b.qb, b.qo, b.gip, b.gep, b.term, b.matrix, out_off=b.out_off, out=tuple(p.array for p in pin))
    best = None
    for it in range(a.reps + 1):
        t0 = time.perf_counter()
        call()
        dt = time.perf_counter() - t0
        sys.stderr.write("[trace_e2e] call %d: %.2f ms\n" % (it, dt * 1e3))
        if it and (best is None or dt < best):
            best = dt
    print("%s %s pairs %d  best %.2f ms  e2e %.0f GCUPS  %.2f M aln/s  env %s" % (a.config, a.format, b.n, best * 1e3, b.cells / best / 1e9, b.n / best / 1e6,
          {k: v for k, v in os.environ.items() if k.startswith("GOTOH_B200")}))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Resident-GCUPS of the other BASELINE.json configurations (parity-test shapes, not the bench line):

    C2b  1M 251-nt reads vs HXB2 pol, align_it(ref, q, 10, 10, 0)   (reference_distances.py:31-41 gap model)
    C3   1M ~84-aa windows vs PR/RT/INT, align_it_aa(ref, q, 40, 10, term) for term in {1, 0}
    C4   consensus vs full-length HCV genomes (~9.6 kb x ~9.6 kb), align_it(ref, q, 15, 3, 1)

Each line also re-checks a sample of the timed run against the oracle (bit-exact or abort).
    python tools/bench_configs.py [--c3 N] [--c4 N] [--c2 N] [--steps K]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "micall-lite_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402


def run(name, al, ora, matrix, rb, ro, ridx, qb, qo, gip, gep, term, steps, verify):
    from gotoh_b200 import packing
    t0 = time.perf_counter()
    plan = al.plan(rb, ro, ridx, qb, qo, gip, gep, term, matrix)
    t_create = time.perf_counter() - t0
    plan.run()
    ms, fms = [], []
    for _ in range(steps):
        d, f = plan.run()
        ms.append(d)
        fms.append(f)
    out = plan.fetch()
    n = len(qo) - 1
    cells = plan.cells
    bad = 0
    idx = list(range(0, n, max(1, n // verify)))
    sub_q = [qb[qo[k]:qo[k + 1]] for k in idx]
    sqo = np.zeros(len(idx) + 1, np.int64)
    np.cumsum([len(q) for q in sub_q], out=sqo[1:])
    sridx = np.array([0 if ridx is None else ridx[k] for k in idx], np.int32) if ridx is not None else None
    exp = ora.align_batch(matrix, rb, ro, sridx, np.concatenate(sub_q), sqo, gip, gep, term)
    for t, k in enumerate(idx):
        ln = int(exp[3][t])
        s0, e0 = int(exp[2][t]), int(plan.out_off[k])
        ok = (ln == int(out[2][k]) and int(exp[4][t]) == int(out[3][k]) and
              (exp[0][s0:s0 + ln] == out[0][e0:e0 + ln]).all() and (exp[1][s0:s0 + ln] == out[1][e0:e0 + ln]).all())
        bad += (not ok)
    line = {"config": name, "pairs": n, "cells": cells, "gcups": cells / (min(ms) * 1e-3) / 1e9,
            "gcups_forward": cells / (min(fms) * 1e-3) / 1e9, "alignments_per_s": n / (min(ms) * 1e-3),
            "ms_per_step": min(ms), "plan_create_s": t_create, "pairs_int16x2": plan.stat(5), "pairs_int32": plan.stat(6),
            "chunks": plan.stat(7), "verified": len(idx), "mismatches": bad}
    plan.close()
    print(json.dumps(line), flush=True)
    if bad:
        raise SystemExit("%s: %d of %d verified pairs differ from the oracle" % (name, bad, len(idx)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c2", type=int, default=1000000)
    ap.add_argument("--c3", type=int, default=1000000)
    ap.add_argument("--c4", type=int, default=2000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--verify", type=int, default=100)
    ap.add_argument("--c4-modes", default="flow,cta,warp", help="long-pair kernels to time on C4 (GOTOH_B200_LONG values)")
    a = ap.parse_args()
    import gotoh_b200
    from gotoh_b200 import packing, workloads
    from gotoh_b200.api import Aligner
    from oracle.oracle import Oracle, have_reference
    al = Aligner()
    ora = Oracle("reference" if have_reference() else "port")
    if a.c2:
        ref, qb, qo = workloads.c2_reads_packed(a.c2, seed=20260102)
        rb, ro = packing.pack([ref])
        run("C2b align_it(10,10,0)", al, ora, gotoh_b200.NT, rb, ro, np.zeros(a.c2, np.int32), qb, qo, 10, 10, 0, a.steps, a.verify)
    if a.c3:
        refs, ridx, qb, qo = workloads.c3_queries_packed(a.c3)
        rb, ro = packing.pack(refs)
        for term in (1, 0):
            run("C3 align_it_aa(40,10,%d)" % term, al, ora, gotoh_b200.HIV25, rb, ro, ridx, qb, qo, 40, 10, term, a.steps, a.verify)
    if a.c4:
        refs, ridx, qb, qo = workloads.c4_pairs_packed(a.c4)
        rb, ro = packing.pack(refs)
        for mode in a.c4_modes.split(","):
            os.environ["GOTOH_B200_LONG"] = mode
            run("C4 align_it(15,3,1) HCV genomes, long-pair kernel=%s" % mode, al, ora, gotoh_b200.NT, rb, ro, ridx, qb, qo,
                15, 3, 1, max(1, a.steps - 1), min(a.verify, 6))
        os.environ.pop("GOTOH_B200_LONG")


if __name__ == "__main__":
    main()

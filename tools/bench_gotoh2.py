#!/usr/bin/env python
"""GCUPS of the live aligner path (gotoh2.Aligner.align on the GPU) on the shapes its two
callers produce, each re-checked against the oracle and timed beside the reference's own `_gotoh2.c`:

    G-remap  consensus vs HCV seed genomes (~9.6 kb x ~9.6 kb), Aligner(15, 3, global, HYPHY_NUC)   remap.py:33,248
    G-aa     amino-acid consensus windows vs PR/RT/INT, Aligner(40, 10, local, EmpHIV25)           aln2counts.py:34-37
    G-reads  251-nt reads vs the HXB2 pol seed (3039 nt), Aligner(10, 3, local, HYPHY_NUC)

A unit is one cell of the (l1+1) x (l2+1) grid `_gotoh2.c:137-201` fills.  `gcups` uses the CUDA-event time of
the kernels (inputs resident), `e2e_gcups` the wall time of gotoh_b200_gotoh2_align_batch with host buffers.
    python tools/bench_gotoh2.py [--remap N] [--aa N] [--reads N] [--steps K] [--cpu-seconds S]
"""
import argparse
import ctypes
import json
import multiprocessing
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "micall-lite_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402


def _cpu_worker(args):
    kind, pairs, gop, gep, glob, model = args
    from oracle.oracle2 import Oracle2
    ora = Oracle2(kind)
    t0 = time.perf_counter()
    for a, b in pairs:
        ora.align(a, b, gop, gep, glob, model)
    return time.perf_counter() - t0


def cpu_baseline(pairs, gop, gep, glob, model, seconds):
    """The reference's `_gotoh2.c` (oracle/_ref) - or the oracle port - over all host cores on a bounded sample."""
    from oracle import oracle2
    kind = "reference" if oracle2.have_reference() else "port"
    cores = os.cpu_count() or 1
    # probe one pair to size the sample
    t = _cpu_worker((kind, pairs[:1], gop, gep, glob, model))
    per_core = max(1, min(len(pairs) // cores, int(seconds / max(t, 1e-6))))
    sample = pairs[:per_core * cores]
    shards = [(kind, sample[r::cores], gop, gep, glob, model) for r in range(cores)]
    t0 = time.perf_counter()
    with multiprocessing.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_worker, shards)
    dt = time.perf_counter() - t0
    cells = sum((len(a) + 1) * (len(b) + 1) for a, b in sample)
    return {"value": cells / dt / 1e9, "unit": "GCUPS", "alignments_per_s": len(sample) / dt, "cores": cores, "kind": kind,
            "sample": "%d pairs of this workload over %d processes, %.1f s" % (len(sample), cores, dt)}


def run(name, pairs, gop, gep, glob, model, steps, verify, cpu_seconds):
    from gotoh_b200.gotoh2 import Aligner
    from oracle.oracle2 import Oracle2
    al = Aligner(gop, gep, glob, model)
    lib = al._libobj.lib
    stats = (ctypes.c_double * 11)()
    best = None
    e2e = []
    out = None
    lazy = []
    for it in range(steps + 1):
        t0 = time.perf_counter()
        lz = al.align_batch(pairs, lazy=True)
        lazy.append(time.perf_counter() - t0)
        assert lz[0] == lz[0] and len(lz) == len(pairs)
        t0 = time.perf_counter()
        out = al.align_batch(pairs)
        dt = time.perf_counter() - t0
        lib.gotoh_b200_gotoh2_last_stats(stats, 11)
        s = list(stats)
        if it == 0:
            continue                      # warm-up
        e2e.append(dt)
        if best is None or s[1] < best[1]:
            best = s
    # end to end through the C ABI: packed HOST arrays in, packed host arrays out (cleaning, packing, H2D, kernels, D2H)
    from gotoh_b200 import packing
    import numpy as np
    uniq, s1_idx = {}, np.zeros(len(pairs), np.int32)      # a first sequence shared by many pairs is passed once
    for k, (a, _) in enumerate(pairs):
        s1_idx[k] = uniq.setdefault(a, len(uniq))
    b1, o1 = packing.pack(list(uniq), "seq1")
    b2, o2 = packing.pack([b for _, b in pairs], "seq2")
    out_off = packing.out_offsets(o1, s1_idx, o2)
    out1 = np.zeros(int(out_off[-1]), np.uint8); out2 = np.zeros(int(out_off[-1]), np.uint8)
    out_len = np.zeros(len(pairs), np.int32); out_score = np.zeros(len(pairs), np.int32)
    mat = np.ascontiguousarray(al.matrix, dtype=np.int32)
    cabi = []
    for it in range(steps + 1):
        t0 = time.perf_counter()
        rc = lib.gotoh_b200_gotoh2_align_batch(b1.ctypes.data, o1.ctypes.data, len(uniq), s1_idx.ctypes.data, b2.ctypes.data, o2.ctypes.data,
                                               len(pairs), gop, gep, int(glob), al.alphabet.encode("ascii"), mat.ctypes.data,
                                               out1.ctypes.data, out2.ctypes.data, out_off.ctypes.data, out_len.ctypes.data,
                                               out_score.ctypes.data, 0)
        dt = time.perf_counter() - t0
        al._libobj.check(rc)
        if it:
            cabi.append(dt)
    ora = Oracle2("port")
    idx = list(range(0, len(pairs), max(1, len(pairs) // verify)))
    bad = sum(out[k] != ora.align(pairs[k][0], pairs[k][1], gop, gep, glob, model) for k in idx)
    cells = best[0]
    line = {"config": name, "pairs": len(pairs), "grid_cells": cells, "gcups": cells / (best[1] * 1e-3) / 1e9,
            "alignments_per_s": len(pairs) / (best[1] * 1e-3), "ms_kernels": best[1], "ms_forward": best[2],
            "ms_reverse": best[3], "ms_walk_emit": best[4], "gcups_forward": cells / (best[2] * 1e-3) / 1e9,
            "gcups_reverse": cells / (best[3] * 1e-3) / 1e9, "gpu_launches": best[5], "arena_bytes": best[6],
            "chunks": best[7], "forward_tasks_int16x2": best[10], "e2e_gcups_c_abi": cells / min(cabi) / 1e9, "e2e_s_c_abi": min(cabi),
            "h2d_bytes": best[8], "d2h_bytes": best[9], "e2e_gcups_python_api": cells / min(e2e) / 1e9, "e2e_gcups_python_api_lazy": cells / min(lazy[1:]) / 1e9, "verified": len(idx), "mismatches": bad,
            "params": {"gop": gop, "gep": gep, "is_global": glob, "model": model}}
    if cpu_seconds > 0:
        line["cpu_baseline"] = cpu_baseline(pairs, gop, gep, glob, model, cpu_seconds)
    print(json.dumps(line), flush=True)
    if bad:
        raise SystemExit("%s: %d of %d verified pairs differ from the oracle" % (name, bad, len(idx)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--remap", type=int, default=200)
    ap.add_argument("--aa", type=int, default=200000)
    ap.add_argument("--reads", type=int, default=20000)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--verify", type=int, default=8)
    ap.add_argument("--cpu-seconds", type=float, default=4.0, help="per-core CPU budget of the baseline sample (0: skip)")
    a = ap.parse_args()
    from gotoh_b200 import workloads
    if a.reads:
        ref, reads = workloads.c2_reads(a.reads, seed=11)
        run("G-reads Aligner(10,3,local,HYPHY_NUC) 251-nt reads vs HXB2 pol seed", [(ref, r) for r in reads],
            10, 3, False, "HYPHY_NUC", a.steps, a.verify, a.cpu_seconds)
    if a.aa:
        refs, qs = workloads.c3_queries(a.aa, seed=12)
        run("G-aa Aligner(40,10,local,EmpHIV25) aa windows vs PR/RT/INT (aln2counts.py:34-37)",
            [(refs[k % 3], q) for k, q in enumerate(qs)], 40, 10, False, "EmpHIV25", a.steps, max(a.verify, 50), a.cpu_seconds)
    if a.remap:
        refs, ridx, qb, qo = workloads.c4_pairs_packed(a.remap)
        pairs = [(refs[int(ridx[k])], qb[qo[k]:qo[k + 1]].tobytes().decode()) for k in range(a.remap)]
        run("G-remap Aligner(15,3,global,HYPHY_NUC) consensus vs HCV seed genomes (remap.py:33,248)", pairs,
            15, 3, True, "HYPHY_NUC", max(1, a.steps - 1), min(a.verify, 3), a.cpu_seconds)


if __name__ == "__main__":
    main()

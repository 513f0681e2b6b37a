#!/usr/bin/env python
"""Drop-in latency: how long does a caller wait for a batch of 1 / 10 / 100 / 1000 alignments through the public Python
API, against the CPU reference (oracle/_ref = the reference's own gotoh.cpp / _gotoh2.c, one core, called pair by pair
exactly like the reference's callers do) - and from which batch size on the GPU path is the faster one.

Shapes = what the reference's callers align:
  aa_pr / aa_rt     84-aa window vs PR (99 aa) / RT (440 aa), align_it_aa(ref, q, 40, 10, 1)        aln2counts.py:213-268
  nt_read           251-nt read vs HXB2 pol (3039 nt), align_it(ref, q, 10, 3, 1)                   C2
  nt_refdist        250-nt read vs an HCV genome (9.6 kb), align_it(ref, q, 10, 10, 0)              reference_distances.py:31-41
  nt_genome         HCV consensus vs HCV genome (9.6 kb x 9.6 kb), align_it(ref, q, 15, 3, 1)       C4 / remap.py:248
  g2_aa / g2_read   the live aligner gotoh2.Aligner.align on the aa-window and read shapes          aln2counts.py:34-37, remap.py:33

    python tools/bench_latency.py [--reps 5] > profiles/r02_latency.json
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


def best_of(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--sizes", default="1,10,100,1000")
    ap.add_argument("--emu", action="store_true", help=argparse.SUPPRESS)
    a = ap.parse_args()
    import gotoh_b200
    from gotoh_b200 import workloads
    from gotoh_b200.api import Aligner
    from gotoh_b200.gotoh2 import Aligner as Aligner2
    from oracle.oracle import Oracle, have_reference
    from oracle.oracle2 import Oracle2, have_reference as have_reference2
    if a.emu:
        sys.path.insert(0, os.path.join(ROOT, "tests", "simt_emu"))
        import build_emu
        from gotoh_b200 import _ffi
        lib = _ffi.Library(build_emu.build())
        al = Aligner(lib)
    else:
        lib = None
        al = Aligner()
    ora = Oracle("reference" if have_reference() else "port")
    ora2 = Oracle2("reference" if have_reference2() else "port")
    sizes = [int(x) for x in a.sizes.split(",")]
    nmax = max(sizes)
    aa_refs, aa_q = workloads.c3_queries(3 * nmax, seed=31)
    pol, reads = workloads.c2_reads(nmax, seed=32)
    hcv = workloads.hcv_seeds()
    seeds, ridx, qb, qo = workloads.c4_pairs_packed(min(nmax, 100), seed=33)
    genomes = workloads.unpacked(qb, qo)
    shapes = [
        ("aa_pr", "align_it_aa(PR 99 aa, 84-aa window, 40, 10, 1)", 1, [aa_refs[0]], aa_q[0::3][:nmax], None, (40, 10, 1)),
        ("aa_rt", "align_it_aa(RT 440 aa, 84-aa window, 40, 10, 1)", 1, [aa_refs[1]], aa_q[1::3][:nmax], None, (40, 10, 1)),
        ("nt_read", "align_it(HXB2 pol 3039 nt, 251-nt read, 10, 3, 1)", 0, [pol], reads, None, (10, 3, 1)),
        ("nt_refdist", "align_it(HCV genome 9.6 kb, 251-nt read, 10, 10, 0)", 0, [hcv[0]], reads, None, (10, 10, 0)),
        ("nt_genome", "align_it(HCV genome, HCV consensus 9.6 kb, 15, 3, 1)", 0, seeds, genomes, [int(x) for x in ridx], (15, 3, 1)),
    ]
    out = {"_how": "tools/bench_latency.py: wall time of one call of the public Python API (list of str in, list of (str, str, int) out; "
                   "best and median of %d after a warm-up) vs the CPU reference called pair by pair on one core" % a.reps, "rows": []}
    for name, what, matrix, refs, qs, rix, (gip, gep, term) in shapes:
        cpu_fn = ora.align_it if matrix == 0 else ora.align_it_aa
        ncpu = min(len(qs), 3 if name == "nt_genome" else 20)
        t0 = time.perf_counter()
        for k in range(ncpu):
            cpu_fn(refs[0 if rix is None else rix[k]], qs[k], gip, gep, term)
        cpu_per = (time.perf_counter() - t0) / ncpu
        for n in sizes:
            if n > len(qs):
                continue
            q = qs[:n]
            ri = [0] * n if rix is None else rix[:n]
            strings = lambda: al.align_batch(refs, q, gip, gep, term, matrix, ref_idx=ri)
            compact = lambda: al.align_batch(refs, q, gip, gep, term, matrix, ref_idx=ri, compact=True)
            tb, tm = best_of(strings, a.reps)
            cb, cm = best_of(compact, a.reps)
            row = {"shape": name, "call": what, "batch": n, "gpu_strings_ms": tb * 1e3, "gpu_strings_median_ms": tm * 1e3,
                   "gpu_compact_ms": cb * 1e3, "cpu_reference_ms": cpu_per * n * 1e3, "cpu_per_pair_ms": cpu_per * 1e3,
                   "speedup_strings": cpu_per * n / tb, "speedup_compact": cpu_per * n / cb}
            if n == 1:
                single = {0: al.align_it, 1: al.align_it_aa}[matrix]
                sb, _ = best_of(lambda: single(refs[ri[0]], q[0], gip, gep, term), a.reps)
                row["gpu_single_call_ms"] = sb * 1e3
            out["rows"].append(row)
            print(json.dumps(row), file=sys.stderr)
    # the live aligner (gotoh2.Aligner): aa windows (aln2counts) and reads (remap)
    for name, what, model, gop, gep, glob, s1, s2 in (
            ("g2_aa", "Aligner(40, 10, local, EmpHIV25).align(RT, 84-aa window)", "EmpHIV25", 40, 10, False, aa_refs[1], aa_q[1::3][:nmax]),
            ("g2_read", "Aligner(10, 3, local, HYPHY_NUC).align(HXB2 pol, 251-nt read)", "HYPHY_NUC", 10, 3, False, pol, reads)):
        kw = {"library": lib} if lib is not None else {}
        g2 = Aligner2(gop, gep, glob, model, **kw)
        ncpu = min(len(s2), 10)
        t0 = time.perf_counter()
        for k in range(ncpu):
            ora2.align(s1, s2[k], gop, gep, glob, model)
        cpu_per = (time.perf_counter() - t0) / ncpu
        for n in sizes:
            pairs = [(s1, b) for b in s2[:n]]
            tb, tm = best_of(lambda: g2.align_batch(pairs), a.reps)
            row = {"shape": name, "call": what, "batch": n, "gpu_strings_ms": tb * 1e3, "gpu_strings_median_ms": tm * 1e3,
                   "cpu_reference_ms": cpu_per * n * 1e3, "cpu_per_pair_ms": cpu_per * 1e3, "speedup_strings": cpu_per * n / tb}
            if n == 1:
                sb, _ = best_of(lambda: g2.align(s1, s2[0]), a.reps)
                row["gpu_single_call_ms"] = sb * 1e3
            out["rows"].append(row)
            print(json.dumps(row), file=sys.stderr)
    # crossover: the smallest measured batch from which the GPU call is faster than the CPU reference loop
    cross = {}
    for r in out["rows"]:
        if r["speedup_strings"] >= 1.0 and r["shape"] not in cross:
            cross[r["shape"]] = r["batch"]
    out["crossover_batch"] = cross
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

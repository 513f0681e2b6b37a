#!/usr/bin/env python
"""C1 substitute: ALL 19,200 reads of the reference's example run
(/root/reference/examples/HIV1C-pol_S1_L001_R{1,2}_001.fastq.gz, 9,600 pairs x ~250 nt; R2 reverse-complemented with the
semantics of micall/utils/translation.py:36-37) through align_it(HIV1B-pol-seed, read, 10, 3, 1).

Writes, next to this script (the GPU box has no /root/reference, so both travel as fixtures):
  c1_reads.txt.xz   the reads, one per line (R1 reads, then reverse-complemented R2 reads)
  c1_golden.npz     per read: score, aligned length and the first 8 bytes of sha256(aligned_standard + b"\\n" +
                    aligned_seq), all from the reference's own gotoh.cpp compiled unmodified (oracle/_ref)

    python tests/golden/make_golden_c1.py        # needs /root/reference; ~1 min on 8 cores
"""
import gzip
import hashlib
import lzma
import multiprocessing
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "micall-lite_b200")):
    sys.path.insert(0, p)

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def revcomp(s):
    return "".join(_COMP.get(c, c) for c in reversed(s))


def digest8(a, b):
    return np.frombuffer(hashlib.sha256(a.encode("latin-1") + b"\n" + b.encode("latin-1")).digest()[:8], dtype=np.uint64)[0]


def _work(args):
    ref, reads = args
    from oracle.oracle import Oracle
    ora = Oracle("reference")
    out = []
    for r in reads:
        a, b, sc = ora.align_it(ref, r, 10, 3, 1)
        out.append((sc, len(a), digest8(a, b)))
    return out


def main():
    from gotoh_b200 import workloads
    from oracle import oracle as om
    om.build()
    assert om.have_reference(), "needs oracle/_ref (the compiled reference)"
    ref = workloads.pol_seed()
    reads = []
    for rno in (1, 2):
        with gzip.open("/root/reference/examples/HIV1C-pol_S1_L001_R%d_001.fastq.gz" % rno, "rt") as f:
            rs = [l.strip() for i, l in enumerate(f) if i % 4 == 1]
        reads += rs if rno == 1 else [revcomp(r) for r in rs]
    assert len(reads) == 19200, len(reads)
    with lzma.open(os.path.join(HERE, "c1_reads.txt.xz"), "wt", preset=9) as f:
        f.write("\n".join(reads) + "\n")
    cores = os.cpu_count() or 1
    chunks = [(ref, reads[i::cores]) for i in range(cores)]
    with multiprocessing.get_context("fork").Pool(cores) as pool:
        res = pool.map(_work, chunks)
    score = np.zeros(len(reads), np.int32)
    ln = np.zeros(len(reads), np.int32)
    dg = np.zeros(len(reads), np.uint64)
    for i, part in enumerate(res):
        for j, (sc, l, d) in enumerate(part):
            score[i + j * cores], ln[i + j * cores], dg[i + j * cores] = sc, l, d
    np.savez_compressed(os.path.join(HERE, "c1_golden.npz"), score=score, out_len=ln, digest8=dg)
    print("wrote %d reads; score range %d..%d" % (len(reads), score.min(), score.max()))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Golden answers for the C4 stress test (tests/test_gpu_scale.py): every 10th of the 2,000 consensus-vs-genome pairs
of workloads.c4_pairs_packed(2000, seed=20260144) (~9.6 kb x ~9.6 kb on the HCV seeds) through align_it(ref, q, 15, 3, 1)
of the reference's own gotoh.cpp compiled unmodified (oracle/_ref).  Writes c4_golden.npz: pair index, score, aligned
length, first 8 bytes of sha256(aligned_standard + b"\\n" + aligned_seq).

    python tests/golden/make_golden_c4.py        # needs /root/reference; ~1.7 s per pair per core
"""
import multiprocessing
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "micall-lite_b200"), HERE):
    sys.path.insert(0, p)

from make_golden_c1 import digest8  # noqa: E402

N_PAIRS, SEED, EVERY = 2000, 20260144, 10


def _work(args):
    refs, ridx, qb, qo, ks = args
    from oracle.oracle import Oracle
    ora = Oracle("reference")
    out = []
    for k in ks:
        a, b, sc = ora.align_it(refs[int(ridx[k])], qb[qo[k]:qo[k + 1]].tobytes().decode(), 15, 3, 1)
        out.append((k, sc, len(a), digest8(a, b)))
    return out


def main():
    from gotoh_b200 import workloads
    from oracle import oracle as om
    om.build()
    assert om.have_reference()
    refs, ridx, qb, qo = workloads.c4_pairs_packed(N_PAIRS, seed=SEED)
    ks = list(range(0, N_PAIRS, EVERY))
    cores = os.cpu_count() or 1
    with multiprocessing.get_context("fork").Pool(cores) as pool:
        res = pool.map(_work, [(refs, ridx, qb, qo, ks[i::cores]) for i in range(cores)])
    rows = sorted(r for part in res for r in part)
    np.savez_compressed(os.path.join(HERE, "c4_golden.npz"), pair=np.array([r[0] for r in rows], np.int32),
                        score=np.array([r[1] for r in rows], np.int32), out_len=np.array([r[2] for r in rows], np.int32),
                        digest8=np.array([r[3] for r in rows], np.uint64),
                        qry_sha=np.frombuffer(__import__("hashlib").sha256(qb.tobytes()).digest()[:8], np.uint64))
    print("wrote %d pairs" % len(rows))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Short reads, aa windows (device-built and host-built plans, all three result forms), a few multi-strip pairs in the three
long-pair modes, gotoh2 and the edit distance; every result is compared with the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "micall-lite_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402


def main():
    import gotoh_b200
    from gotoh_b200 import packing, workloads, remap_filter
    from gotoh_b200.api import Aligner
    from gotoh_b200.gotoh2 import Aligner as Aligner2
    from oracle.oracle import Oracle
    from oracle.oracle2 import Oracle2, levenshtein
    al, ora = Aligner(), Oracle("port")
    n = int(os.environ.get("SANITIZE_PAIRS", "600"))
    ref, reads = workloads.c2_reads(n, seed=3)
    arefs, aq = workloads.c3_queries(3 * n, seed=4)
    for prep in ("1", "0"):
        os.environ["GOTOH_B200_DEVICE_PREP"] = prep
        for matrix, refs, qs, ridx, g in ((0, [ref[:900]], reads, [0] * n, (10, 3, 1)), (1, arefs, aq, [k % 3 for k in range(3 * n)], (40, 10, 1))):
            rb, ro = packing.pack(refs)
            qb, qo = packing.pack(qs)
            r = np.asarray(ridx, np.int32)
            s = al.align_packed(rb, ro, r, qb, qo, g[0], g[1], g[2], matrix)
            t = al.align_packed_tight(rb, ro, r, qb, qo, g[0], g[1], g[2], matrix)
            c = al.align_packed_compact(rb, ro, r, qb, qo, g[0], g[1], g[2], matrix)
            assert (t[4] == s[4]).all() and (c.scores == s[4]).all()
            fn = ora.align_it if matrix == 0 else ora.align_it_aa
            for k in range(0, len(qs), 37):
                o, ln = int(s[2][k]), int(s[3][k])
                exp = fn(refs[ridx[k]], qs[k], *g)
                assert (s[0][o:o + ln].tobytes().decode(), s[1][o:o + ln].tobytes().decode(), int(s[4][k])) == exp
                assert c[k] == exp
    os.environ.pop("GOTOH_B200_DEVICE_PREP")
    seeds, ridx, qb, qo = workloads.c4_pairs_packed(3, seed=5)
    qs = [q[:1500] for q in workloads.unpacked(qb, qo)]
    for mode in ("flow", "cta", "warp"):
        os.environ["GOTOH_B200_LONG"] = mode
        got = al.align_batch([s[:1400] for s in seeds], qs, 15, 3, 1, 0, ref_idx=[int(x) for x in ridx])
        for k, g in enumerate(got):
            assert g == ora.align_it(seeds[int(ridx[k])][:1400], qs[k], 15, 3, 1), (mode, k)
    os.environ.pop("GOTOH_B200_LONG")
    ora2 = Oracle2("port")
    pairs = [(ref[:700], r) for r in reads[:40]]
    for (gop, gep, glob) in ((15, 3, True), (10, 3, False)):
        got = Aligner2(gop, gep, glob, "HYPHY_NUC").align_batch(pairs)
        for (a, b), g in zip(pairs, got):
            assert g == ora2.align(a, b, gop, gep, glob, "HYPHY_NUC")
    assert remap_filter.distance_batch(pairs) == [levenshtein(a, b) for a, b in pairs]
    print("sanitize smoke ok")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Randomised differential check on the CPU SIMT emulator: host vs device plan builder x half-warp vs 32-lane wavefronts
(x the int32 kernels), random tables, references (incl. "$$$", IUPAC, lower case, surrounding whitespace, 1-5 rows), query
widths 1..600 (multi-strip), gap models incl. zero penalties, the three result forms - every pair also against the oracle.
    python tools/fuzz_emu.py [seed] [trials]
FUZZ_LONG=1: references of 200-1500 rows with tiny gap-open penalties (many rebase rows); FUZZ_TALL=1: 3 000-12 000 rows.  FUZZ_STRIPS=1: 24 pairs per
trial with queries of 257-1800 columns (2-8 strips) through the three multi-strip forward kernels (strip dataflow, CTA per
pair, warp per pair: GOTOH_B200_LONG=flow|cta|warp) and the int16x2-free host/device builders.
(seed 11 of the first version found the column-0 seed bug of DESIGN.md 3.8.)"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "micall-lite_b200"), os.path.join(ROOT, "tests", "simt_emu")]
import numpy as np  # noqa: E402


def main():
    import build_emu
    from gotoh_b200 import _ffi, packing
    from gotoh_b200.api import Aligner
    from oracle.oracle import Oracle
    al = Aligner(_ffi.Library(build_emu.build()))
    ora = Oracle("port")
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    trials = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    rng = random.Random(seed)
    bad = total = 0
    strips_mode = os.environ.get("FUZZ_STRIPS") == "1"
    tall_mode = os.environ.get("FUZZ_TALL") == "1"      # 40 pairs per trial on references of 3 000-12 000 rows: hundreds of rebase rows
    long_mode = os.environ.get("FUZZ_LONG") == "1" or strips_mode or tall_mode      # references up to 1500 rows, tiny gap-open: many rebase rows
    for trial in range(trials):
        matrix = rng.choice([0, 0, 1, 2])
        if matrix == 0:
            alpha = rng.choice(["ACGT", "ACGTNRY-acgt", "ACGTNRYKMSWBDHVacgtnXx*.-Uu$"])
        else:
            alpha = rng.choice(["ARNDCQEGHILKMFPSTWYV", "ARNDCQEGHILKMFPSTWYVBZX*-?Jj_"])
        nrefs = rng.randint(1, 5)
        refs = []
        for _ in range(nrefs):
            a = "".join(rng.choice(alpha) for _ in range(rng.choice([rng.randint(1, 5), rng.randint(6, 260), rng.randint(6, 260)]) if not long_mode else rng.randint(3000, 12000) if tall_mode else rng.randint(200, 1500)))
            if matrix == 0 and rng.random() < 0.3:
                p = rng.randrange(len(a) + 1)
                a = a[:p] + rng.choice(["$$$", "$$$$", "$$"]) + a[p:]
            if matrix == 2 and not a.replace("-", ""):
                a += "K"
            refs.append(a)
        wide = rng.random() < 0.25
        qs, ridx = [], []
        for k in range(24 if strips_mode else 40 if tall_mode else 160):
            r = rng.randrange(nrefs)
            a = refs[r]
            nmax = 1800 if strips_mode else 600 if (wide and k % 20 == 0) else 256
            if strips_mode and rng.random() < 0.7:
                lo = rng.randrange(max(1, len(a) - 300))
                q = list(a[lo:lo + rng.randint(257, 1500)].replace("$", "A"))
                for _ in range(rng.randint(0, 6)):
                    q[rng.randrange(len(q))] = rng.choice(alpha.replace("$", "A"))
                for _ in range(rng.randint(0, 3)):                       # deletions and insertions of 1-40 characters
                    p0 = rng.randrange(len(q))
                    del q[p0:p0 + rng.randint(1, 40)]
                for _ in range(rng.randint(0, 3)):
                    p0 = rng.randrange(len(q) + 1)
                    q[p0:p0] = [rng.choice(alpha.replace("$", "T")) for _ in range(rng.randint(1, 40))]
                while len(q) < 257:
                    q.append(rng.choice(alpha.replace("$", "T")))
                q = "".join(q)
            elif strips_mode:
                q = "".join(rng.choice(alpha.replace("$", "T")) for _ in range(rng.randint(257, nmax)))
            elif rng.random() < 0.6:
                lo = rng.randrange(len(a))
                q = list(a[lo:lo + rng.randint(1, 200)].replace("$$$", rng.choice(["TAG", "TAA", "TGA", "TGG"])).replace("$", "A") or "A")
                for _ in range(rng.randint(0, 4)):
                    q[rng.randrange(len(q))] = rng.choice(alpha.replace("$", "A"))
                q = "".join(q)
            else:
                q = "".join(rng.choice(alpha.replace("$", "T")) for _ in range(rng.randint(1, nmax)))
            q = q[:nmax]
            if matrix == 2 and not q.replace("-", ""):
                q += "R"
            if rng.random() < 0.1:
                q = rng.choice([" ", "\t", ""]) + q + rng.choice(["\n", " \r\n", ""])
            qs.append(q)
            ridx.append(r)
        gip = rng.choice([0, 0, 1, 3, 10, 15, 40]) if not long_mode else rng.choice([0, 0, 1, 2, 5])
        gep = rng.choice([0, 0, 1, 3, 10]) if not long_mode else rng.choice([1, 3, 10, 20, 40])
        if tall_mode:
            gep = rng.choice([1, 2, 3, 5, 8])          # 2*gip + (M+1)*gep must stay below the reference's sentinel (gotoh.cpp:284)
        if strips_mode and rng.random() < 0.5:
            gip, gep = rng.choice([(15, 3), (10, 3), (6, 1), (40, 10), (3, 0), (0, 0)])
        term = 0 if matrix == 2 else rng.choice([0, 1])
        rb, ro = packing.pack(refs)
        qb, qo = packing.pack(qs)
        ri = np.asarray(ridx, np.int32)
        res = {}
        paths = (("host", {"GOTOH_B200_DEVICE_PREP": "0"}), ("device", {"GOTOH_B200_DEVICE_PREP": "1"}),
                 ("full", {"GOTOH_B200_DEVICE_PREP": "1", "GOTOH_B200_HALF": "0"}), ("int32", {"GOTOH_B200_FORCE_PATH": "32"}))
        if strips_mode:
            paths = (("host", {"GOTOH_B200_DEVICE_PREP": "0"}), ("device", {"GOTOH_B200_DEVICE_PREP": "1"}),
                     ("cta", {"GOTOH_B200_LONG": "cta"}), ("warp", {"GOTOH_B200_LONG": "warp"}))
        for key, env in paths:
            for v in ("GOTOH_B200_DEVICE_PREP", "GOTOH_B200_HALF", "GOTOH_B200_FORCE_PATH", "GOTOH_B200_LONG"):
                os.environ.pop(v, None)
            os.environ.update(env)
            res[key] = al.align_packed(rb, ro, ri, qb, qo, gip, gep, term, matrix)
        comp = al.align_packed_compact(rb, ro, ri, qb, qo, gip, gep, term, matrix)
        base = res["host"]
        for key, v in res.items():
            for x in (0, 1, 3, 4):
                if not (v[x] == base[x]).all():
                    bad += 1
                    print("MISMATCH between paths", key, "trial", trial, "array", x, gip, gep, term, matrix)
        fn = {0: ora.align_it, 1: ora.align_it_aa}.get(matrix)
        for k in range(len(qs)):
            o, ln = int(base[2][k]), int(base[3][k])
            got = (base[0][o:o + ln].tobytes().decode("latin-1"), base[1][o:o + ln].tobytes().decode("latin-1"), int(base[4][k]))
            exp = fn(refs[ridx[k]], qs[k], gip, gep, term) if fn else ora.align(2, refs[ridx[k]], qs[k], gip, gep, 0)
            total += 1
            if got != exp or comp[k] != exp:
                bad += 1
                print("ORACLE MISMATCH trial", trial, "pair", k, (gip, gep, term, matrix), repr(refs[ridx[k]][:40]), repr(qs[k][:40]), got[2], exp[2])
    print("fuzz seed %d done: %d pairs, %d bad" % (seed, total, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

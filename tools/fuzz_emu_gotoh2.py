#!/usr/bin/env python
"""Randomised check of the live aligner (gotoh2.Aligner on the kernels under the CPU SIMT emulator) against the oracle:
random models, global / local, penalties incl. zero, sequence lengths 1..600, shared and distinct first sequences; and
of the edit-distance entry point against Wagner-Fischer.
    python tools/fuzz_emu_gotoh2.py [seed] [trials]
FUZZ_LONG=1: 10 pairs per trial of 600-2 500 x 300-2 500 characters with indels (the remap caller's shape, scaled down:
several strips and arena chunks, the tiled traceback)."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "micall-lite_b200"), os.path.join(ROOT, "tests", "simt_emu")]


def main():
    import build_emu
    from gotoh_b200 import _ffi, remap_filter
    from gotoh_b200.gotoh2 import Aligner
    from oracle.oracle2 import Oracle2, levenshtein
    lib = _ffi.Library(build_emu.build())
    ora = Oracle2("port")
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    trials = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    rng = random.Random(seed)
    bad = total = 0
    models = sorted(ora.models)
    for trial in range(trials):
        model = rng.choice(models)
        alphabet = ora.models[model][1]
        letters = alphabet.replace("?", "") + rng.choice(["", "xn-", "acgt"])
        gop, gep, glob = rng.choice([0, 1, 5, 10, 15, 40]), rng.choice([0, 1, 3, 10]), rng.random() < 0.5
        long_mode = os.environ.get("FUZZ_LONG") == "1"
        shared = "".join(rng.choice(letters) for _ in range(rng.randint(600, 2500) if long_mode else rng.randint(1, 400)))
        pairs = []
        for k in range(10 if long_mode else 0):
            a = shared if rng.random() < 0.6 else "".join(rng.choice(letters) for _ in range(rng.randint(600, 2500)))
            if rng.random() < 0.75:
                lo = rng.randrange(len(a) // 2)
                b = list(a[lo:lo + rng.randint(300, 2500)])
                for _ in range(rng.randint(0, 30)):
                    b[rng.randrange(len(b))] = rng.choice(letters)
                for _ in range(rng.randint(0, 4)):
                    p0 = rng.randrange(len(b))
                    del b[p0:p0 + rng.randint(1, 60)]
                for _ in range(rng.randint(0, 4)):
                    p0 = rng.randrange(len(b) + 1)
                    b[p0:p0] = [rng.choice(letters) for _ in range(rng.randint(1, 60))]
                b = "".join(b) or "A"
            else:
                b = "".join(rng.choice(letters) for _ in range(rng.randint(300, 2500)))
            pairs.append((a, b))
        for k in range(0 if long_mode else 70):
            a = shared if rng.random() < 0.6 else "".join(rng.choice(letters) for _ in range(rng.randint(1, 300)))
            if rng.random() < 0.6:
                lo = rng.randrange(len(a))
                b = list(a[lo:lo + rng.randint(1, 260)] or "A")
                for _ in range(rng.randint(0, 5)):
                    b[rng.randrange(len(b))] = rng.choice(letters)
                b = "".join(b)
            else:
                b = "".join(rng.choice(letters) for _ in range(rng.randint(1, 600 if k % 17 == 0 else 250)))
            pairs.append((a, b))
        al = Aligner(gop, gep, glob, model, library=lib)
        try:
            got = al.align_batch(pairs)
        except RuntimeError as e:                      # "Traceback failed, try local alignment": the oracle must fail too
            got = None
            msg = str(e)
        failed_in_oracle = False
        for k, (a, b) in enumerate(pairs):
            total += 1
            try:
                exp = ora.align(a, b, gop, gep, glob, model)
            except RuntimeError:
                exp = "traceback failed"
            if got is None:
                failed_in_oracle = failed_in_oracle or exp == "traceback failed"
                continue
            if got[k] != exp:
                bad += 1
                print("GOTOH2 MISMATCH trial", trial, "pair", k, (gop, gep, glob, model), repr(a[:30]), repr(b[:30]))
        if got is None and not failed_in_oracle:
            bad += 1
            print("GOTOH2 raised", msg, "but no pair of trial", trial, "fails in the oracle", (gop, gep, glob, model))
        dist_pairs = [(a[:200], b[:200]) for a, b in pairs[:25]]
        try:
            d = remap_filter.distance_batch(dist_pairs, library=lib)
            for (a, b), x in zip(dist_pairs, d):
                total += 1
                if x != levenshtein(a, b):
                    bad += 1
                    print("DISTANCE MISMATCH trial", trial, repr(a[:30]), repr(b[:30]), x)
        except ValueError:
            pass                                        # a pair sharing more than 28 distinct symbols: documented limit
    print("gotoh2 fuzz seed %d done: %d checks, %d bad" % (seed, total, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""Stress of the one-shot call's host pipeline (slab cuts, builder threads, ordered enqueue, two-phase collector, static
sharding over devices) on the CPU SIMT emulator: random batches under random pipeline settings, in every result form,
alone and from several caller threads at once - each result compared byte for byte with a single-slab, single-device,
strided run of the same batch.  One round in seven injects a failing result copy at a random slab (GOTOH_B200_TEST_FAIL_FETCH):
those calls must come back with that error (or complete correctly when the call has fewer slabs), never hang, and the rounds
after them must be clean.
    python tools/stress_emu_pipeline.py [seed] [rounds]"""
import os
import random
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "micall-lite_b200"), os.path.join(ROOT, "tests", "simt_emu")]
os.environ.setdefault("SIMT_EMU_DEVICES", "3")
import numpy as np  # noqa: E402

KNOBS = {"GOTOH_B200_SLAB_MB": ["1", "1", "2", None], "GOTOH_B200_RAMP": ["0", "1", None], "GOTOH_B200_BUILDERS": ["1", "2", "3", None],
         "GOTOH_B200_WORKSPACES": ["2", "3", "8", None], "GOTOH_B200_FWD_DEPTH": ["1", "2", "4", None],
         "GOTOH_B200_DEVICE_PREP": ["0", "1", None], "GOTOH_B200_BLOCKING_SYNC": ["1", None], "GOTOH_B200_HOST_THREADS": ["1", "3", None]}


def batch(rng, n):
    refs, qs = [], []
    shared = "".join(rng.choice("ACGT") for _ in range(rng.randint(50, 500)))
    for k in range(n):
        a = shared if rng.random() < 0.5 else "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 420)))
        if rng.random() < 0.6:
            lo = rng.randrange(len(a))
            b = list(a[lo:lo + rng.randint(1, 330)])
            for _ in range(rng.randint(0, 4)):
                b[rng.randrange(len(b))] = rng.choice("ACGTN-")
            b = "".join(b)
        else:
            b = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 600 if k % 9 == 0 else 250)))
        refs.append(a)
        qs.append(b)
    return refs, qs


def strings(res):
    from gotoh_b200 import packing
    return list(zip(packing.unpack(res[0], res[2], res[3]), packing.unpack(res[1], res[2], res[3]), res[4].tolist()))


def main():
    import build_emu
    from gotoh_b200 import _ffi, packing
    from gotoh_b200.api import Aligner
    al = Aligner(_ffi.Library(build_emu.build()))
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    rng = random.Random(seed)
    bad = calls = injected = 0
    for rnd in range(rounds):
        for k in list(KNOBS) + ["GOTOH_B200_TEST_FAIL_FETCH"]:
            os.environ.pop(k, None)
        jobs = []
        for _ in range(rng.randint(1, 3)):                       # caller threads of this round
            refs, qs = batch(rng, rng.randint(1, 500))
            gip, gep, term = rng.choice([(10, 3, 1), (5, 1, 0), (0, 10, 1), (15, 3, 1)])
            rb, ro = packing.pack(refs)
            qb, qo = packing.pack(qs)
            want = strings(al.align_packed(rb, ro, None, qb, qo, gip, gep, term, 0))
            jobs.append((rb, ro, qb, qo, gip, gep, term, want, rng.choice(["strided", "tight", "compact"]), rng.choice([1, 2, 3, 5, 7])))
        setting = {}
        for k, choices in KNOBS.items():
            v = rng.choice(choices)
            if v is not None:
                os.environ[k] = v
                setting[k] = v
        inject = rng.random() < 1 / 7
        if inject:
            os.environ["GOTOH_B200_TEST_FAIL_FETCH"] = setting["GOTOH_B200_TEST_FAIL_FETCH"] = rng.choice(["0", "1", "2", "3", "5"])
            os.environ["GOTOH_B200_SLAB_MB"] = setting["GOTOH_B200_SLAB_MB"] = "1"
        out = [None] * len(jobs)

        def call(i):
            rb, ro, qb, qo, gip, gep, term, _, form, mask = jobs[i]
            try:
                if form == "strided":
                    out[i] = strings(al.align_packed(rb, ro, None, qb, qo, gip, gep, term, 0, device_mask=mask))
                elif form == "tight":
                    out[i] = strings(al.align_packed_tight(rb, ro, None, qb, qo, gip, gep, term, 0, device_mask=mask))
                else:
                    c = al.align_packed_compact(rb, ro, None, qb, qo, gip, gep, term, 0, device_mask=mask)
                    out[i] = [c[k] for k in range(len(c))]
            except Exception as e:                                # noqa: BLE001 - reported below
                out[i] = e

        ths = [threading.Thread(target=call, args=(i,), daemon=True) for i in range(len(jobs))]
        for t in ths:
            t.start()
        for t in ths:
            t.join(600)
        for i, t in enumerate(ths):
            calls += 1
            if t.is_alive():
                print("HANG round", rnd, "job", i, jobs[i][8], "mask", jobs[i][9], setting, flush=True)
                return 2
            if inject and isinstance(out[i], _ffi.GotohError) and "injected" in str(out[i]):
                injected += 1
                continue
            if out[i] != jobs[i][7]:
                bad += 1
                what = out[i] if isinstance(out[i], Exception) else "%d of %d pairs differ" % (
                    sum(x != y for x, y in zip(out[i], jobs[i][7])), len(jobs[i][7]))
                print("MISMATCH round", rnd, "job", i, jobs[i][8], "mask", jobs[i][9], setting, what, flush=True)
    print("pipeline stress seed %d done: %d calls (%d ended by an injected failure), %d bad" % (seed, calls, injected, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

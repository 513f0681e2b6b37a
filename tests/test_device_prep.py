"""The device-side plan builder (csrc/gotoh_prep.cuh: trim, validation, int16x2 admission, grouping and HBM layout as
kernels) against the host builder (plan_build) and the oracle.  Both builders must produce the same bytes for every
result form; whatever the device builder does not cover must fall back to the host builder, which reports input errors
exactly as before.  CPU: kernel sources under the SIMT emulator; -m gpu: the product library."""
import random

import numpy as np
import pytest

from gotoh_b200 import _ffi, packing, workloads


def _ragged(seed, n, nrefs=4):
    """Queries of every width 1..256 (all K classes of the 32-lane and half-warp kernels) against a few references of
    different and equal lengths, some with surrounding whitespace."""
    rng = random.Random(seed)
    refs = ["".join(rng.choice("ACGT") for _ in range(m)) for m in ([300, 120, 300, 77, 1000, 33] * 20)[:nrefs]]
    qs, ridx = [], []
    for k in range(n):
        r = rng.randrange(nrefs)
        width = rng.choice([rng.randint(1, 256), rng.randint(60, 100), 251])
        lo = rng.randrange(len(refs[r]))
        q = refs[r][lo:lo + width] or "A"
        q = list(q + "".join(rng.choice("ACGTN") for _ in range(rng.randint(0, 4))))[:256]
        for _ in range(rng.randint(0, 3)):
            q[rng.randrange(len(q))] = rng.choice("ACGTRY")
        q = "".join(q)
        if k % 9 == 0:
            q = "  " + q + "\n"
        qs.append(q)
        ridx.append(r)
    return refs, qs, ridx


def _both_builders(aligner, oracle, monkeypatch, matrix, refs, qs, ridx, gip, gep, term, n_oracle=60, expect_device=True):
    rb, ro = packing.pack(refs)
    qb, qo = packing.pack(qs)
    ridx = np.asarray(ridx, np.int32)
    monkeypatch.setenv("GOTOH_B200_DEVICE_PREP", "0")
    h = aligner.align_packed(rb, ro, ridx, qb, qo, gip, gep, term, matrix)
    monkeypatch.setenv("GOTOH_B200_DEVICE_PREP", "1")
    plan = aligner.plan(rb, ro, ridx, qb, qo, gip, gep, term, matrix)
    assert plan.stat(8) == (1 if expect_device else 0), "device builder %s" % ("declined" if expect_device else "accepted")
    plan.run()
    p = plan.fetch()
    assert plan.cells == int((np.diff(ro)[ridx].astype(np.int64) * np.array([len(q.strip()) for q in qs])).sum()) or not expect_device
    plan.close()
    d = aligner.align_packed(rb, ro, ridx, qb, qo, gip, gep, term, matrix)
    c = aligner.align_packed_compact(rb, ro, ridx, qb, qo, gip, gep, term, matrix)
    t = aligner.align_packed_tight(rb, ro, ridx, qb, qo, gip, gep, term, matrix)
    for x in (0, 1, 3, 4):
        assert (h[x] == d[x]).all(), x
    assert (p[0] == h[0]).all() and (p[1] == h[1]).all() and (p[2] == h[3]).all() and (p[3] == h[4]).all()
    assert (c.scores == h[4]).all() and (c.out_len == h[3]).all() and (t[3] == h[3]).all() and (t[4] == h[4]).all()
    a = packing.unpack(h[0], h[2], h[3])
    b = packing.unpack(h[1], h[2], h[3])
    ta = packing.unpack(t[0], t[2], t[3])
    tb = packing.unpack(t[1], t[2], t[3])
    assert ta == a and tb == b
    fn = oracle.align_it if matrix == 0 else oracle.align_it_aa
    for k in range(len(qs)):
        assert c.strings(k) == (a[k], b[k]), k
    for k in range(0, len(qs), max(1, len(qs) // n_oracle)):
        assert (a[k], b[k], int(h[4][k])) == fn(refs[ridx[k]], qs[k], gip, gep, term), k
    return h


@pytest.mark.parametrize("half", ["1", "0"])
def test_emu_device_builder_matches_host_builder(emu_aligner, oracle_port, monkeypatch, half):
    monkeypatch.setenv("GOTOH_B200_HALF", half)
    refs, qs, ridx = _ragged(3, 320)
    _both_builders(emu_aligner, oracle_port, monkeypatch, 0, refs, qs, ridx, 10, 3, 1)
    _both_builders(emu_aligner, oracle_port, monkeypatch, 0, refs, qs[:160], ridx[:160], 10, 10, 0)
    arefs, aq = workloads.c3_queries(301, seed=5)
    _both_builders(emu_aligner, oracle_port, monkeypatch, 1, arefs, aq, [k % 3 for k in range(301)], 40, 10, 1)


def test_emu_device_builder_stop_codon_classes_and_slabs(emu_aligner, oracle_port, monkeypatch):
    """References with "$$$" (rule-mask classes live in the query profile) and a call cut into several slabs."""
    rng = random.Random(8)
    refs = ["ACGT$$$ACGTTAGCA" * 6, "TTGA$$$$CCATAGA" * 5]
    qs = ["".join(rng.choice("ACGTTAGTAA") for _ in range(rng.randint(5, 120))) for _ in range(200)]
    _both_builders(emu_aligner, oracle_port, monkeypatch, 0, refs, qs, [k % 2 for k in range(200)], 10, 3, 1)
    monkeypatch.setenv("GOTOH_B200_SLAB_MB", "1")
    refs, qs, ridx = _ragged(4, 300, nrefs=3)
    _both_builders(emu_aligner, oracle_port, monkeypatch, 0, refs, qs, ridx, 5, 1, 1)


def test_emu_device_builder_falls_back_and_errors_are_reported(emu_aligner, oracle_port, monkeypatch):
    monkeypatch.setenv("GOTOH_B200_DEVICE_PREP", "1")
    refs, qs, ridx = _ragged(5, 120, nrefs=2)
    # a multi-strip query, a pair beyond the int16 proof (huge penalties), > 64 references, degapping: host builder, same results
    _both_builders(emu_aligner, oracle_port, monkeypatch, 0, refs, qs + ["ACGT" * 100], ridx + [0], 10, 3, 1, expect_device=False)
    _both_builders(emu_aligner, oracle_port, monkeypatch, 0, refs, qs, ridx, 10, 90, 1, expect_device=False)
    many = ["".join(random.Random(k).choice("ACGT") for _ in range(40 + k)) for k in range(70)]
    _both_builders(emu_aligner, oracle_port, monkeypatch, 0, many, [m[3:30] for m in many], list(range(70)), 10, 3, 1, expect_device=False)
    got = emu_aligner.align_batch(["K-F-RWW"] * 3, ["KF--GR", "K-FR", "RWW"], 4, 2, 0, _ffi.AA_RB, ref_idx=[0, 1, 2])
    assert [g[:2] for g in got] == [oracle_port.align_it_aa_rb("K-F-RWW", q, 4, 2) for q in ["KF--GR", "K-FR", "RWW"]]
    # input errors come from the host builder with the same codes and messages as before
    for bad, code in (("AC\x7fGT", _ffi.EDOMAIN), ("  \n", _ffi.EEMPTY)):
        with pytest.raises(_ffi.GotohInputError) as e:
            emu_aligner.align_batch(refs, qs[:50] + [bad], 10, 3, 1, 0, ref_idx=ridx[:50] + [0])
        assert e.value.code == code and "query 50" in str(e.value)
    with pytest.raises(_ffi.GotohInputError) as e:
        emu_aligner.align_batch(refs, qs[:50], 10, 2000, 1, 0, ref_idx=ridx[:50])
    assert e.value.code == _ffi.ESENTINEL


@pytest.mark.gpu
def test_gpu_device_builder_matches_host_builder(gpu_aligner, oracle_port, monkeypatch):
    refs, qs, ridx = _ragged(13, 60000, nrefs=6)
    _both_builders(gpu_aligner, oracle_port, monkeypatch, 0, refs, qs, ridx, 10, 3, 1, n_oracle=1500)
    monkeypatch.setenv("GOTOH_B200_HALF", "0")
    _both_builders(gpu_aligner, oracle_port, monkeypatch, 0, refs, qs, ridx, 10, 10, 0, n_oracle=300)
    monkeypatch.delenv("GOTOH_B200_HALF")
    arefs, aq = workloads.c3_queries(250000, seed=15)                       # three slabs of the one-shot pipeline
    _both_builders(gpu_aligner, oracle_port, monkeypatch, 1, arefs, aq, [k % 3 for k in range(len(aq))], 40, 10, 1, n_oracle=3000)
    ref, reads = workloads.c2_reads(30000, seed=16)
    _both_builders(gpu_aligner, oracle_port, monkeypatch, 0, [ref], reads, [0] * len(reads), 10, 3, 1, n_oracle=300)
    # the most references the device builder takes (64), every query width, several slabs
    refs, qs, ridx = _ragged(17, 120000, nrefs=64)
    _both_builders(gpu_aligner, oracle_port, monkeypatch, 0, refs, qs, ridx, 10, 3, 1, n_oracle=800)


@pytest.mark.gpu
def test_gpu_concurrent_callers_share_a_device(gpu_aligner, oracle_port):
    """Two host threads call the one-shot entry points at the same time (ctypes releases the GIL; the per-device context
    serialises them): both get their own results."""
    import threading
    jobs = []
    for seed, form in ((21, "strings"), (22, "compact"), (23, "tight")):
        refs, qs, ridx = _ragged(seed, 20000, nrefs=3)
        jobs.append((form, refs, qs, ridx))
    got = {}

    def run(form, refs, qs, ridx):
        rb, ro = packing.pack(refs)
        qb, qo = packing.pack(qs)
        r = np.asarray(ridx, np.int32)
        fn = {"strings": gpu_aligner.align_packed, "compact": gpu_aligner.align_packed_compact, "tight": gpu_aligner.align_packed_tight}[form]
        got[form] = fn(rb, ro, r, qb, qo, 10, 3, 1, 0)
    th = [threading.Thread(target=run, args=j) for j in jobs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for form, refs, qs, ridx in jobs:
        g = got[form]
        for k in range(0, len(qs), 97):
            exp = oracle_port.align_it(refs[ridx[k]], qs[k], 10, 3, 1)
            if form == "compact":
                assert g[k] == exp
            else:
                o, ln = int(g[2][k]), int(g[3][k])
                assert (g[0][o:o + ln].tobytes().decode(), g[1][o:o + ln].tobytes().decode(), int(g[4][k])) == exp

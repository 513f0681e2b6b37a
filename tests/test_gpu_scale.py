"""GPU parity at the sizes and occupancies the bench runs at (VERDICT r1 "parity holes"): all 19,200 real reads of the
reference's example run, 2,000 full-size HCV genome pairs in one plan (the strip-dataflow kernel at full occupancy,
several arena chunks), the int16x2 admit/fallback boundary of the main aligner on the GPU's own arithmetic, and the
library's multi-device sharding on a box that has more than one device."""
import hashlib
import lzma
import os

import numpy as np
import pytest

from conftest import GOLDEN
from gotoh_b200 import packing, workloads

pytestmark = pytest.mark.gpu


def _digest8(a, b):
    return np.frombuffer(hashlib.sha256(a.tobytes() + b"\n" + b.tobytes()).digest()[:8], dtype=np.uint64)[0]


def test_gpu_c1_all_example_reads_vs_reference_hashes(gpu_aligner):
    """C1 substitute: every read of examples/HIV1C-pol_S1_L001_R{1,2} (R2 reverse-complemented) vs
    HIV1B-pol-seed, align_it(ref, read, 10, 3, 1); golden = the reference's own gotoh.cpp (tests/golden/make_golden_c1.py)."""
    with lzma.open(os.path.join(GOLDEN, "c1_reads.txt.xz"), "rt") as f:
        reads = f.read().split()
    g = np.load(os.path.join(GOLDEN, "c1_golden.npz"))
    assert len(reads) == 19200 == len(g["score"])
    ref = workloads.pol_seed()
    rb, ro = packing.pack([ref])
    qb, qo = packing.pack(reads)
    out_a, out_b, off, ln, sc = gpu_aligner.align_packed(rb, ro, np.zeros(len(reads), np.int32), qb, qo, 10, 3, 1, 0)
    assert (sc == g["score"]).all(), np.nonzero(sc != g["score"])[0][:5]
    assert (ln == g["out_len"]).all()
    for k in range(len(reads)):
        o = int(off[k])
        assert _digest8(out_a[o:o + ln[k]], out_b[o:o + ln[k]]) == g["digest8"][k], k
    c = gpu_aligner.align_packed_compact(rb, ro, np.zeros(len(reads), np.int32), qb, qo, 10, 3, 1, 0)
    assert (c.scores == g["score"]).all() and (c.out_len == g["out_len"]).all()
    for k in range(0, len(reads), 7):
        assert _digest8(*c.arrays(k)) == g["digest8"][k], k


@pytest.mark.parametrize("repeat", [0, 1, 2])
def test_gpu_c4_two_thousand_long_pairs_one_plan(gpu_aligner, monkeypatch, repeat):
    """2,000 ~9.6 kb x ~9.6 kb pairs in ONE plan under GOTOH_B200_LONG=flow: 76,000 strip tasks, every SM full, the arena cut
    into chunks.  200 pairs are compared with the reference's own answers (tests/golden/make_golden_c4.py), all 2,000
    with the size-independent properties; three repeats because the kernel's cross-SM hand-over is timing dependent."""
    monkeypatch.setenv("GOTOH_B200_LONG", "flow")
    monkeypatch.setenv("GOTOH_B200_ARENA_MB", "30000")      # 46 GB of directions -> two chunks of ~1,300 pairs
    g = np.load(os.path.join(GOLDEN, "c4_golden.npz"))
    refs, ridx, qb, qo = workloads.c4_pairs_packed(2000, seed=20260144)
    assert np.frombuffer(hashlib.sha256(qb.tobytes()).digest()[:8], np.uint64)[0] == g["qry_sha"][0], "generator drifted from the golden inputs"
    rb, ro = packing.pack(refs)
    plan = gpu_aligner.plan(rb, ro, ridx, qb, qo, 15, 3, 1, 0)
    plan.run()
    assert plan.stat(6) == 2000 and plan.stat(7) >= 2, (plan.stat(6), plan.stat(7))
    out_a, out_b, ln, sc = plan.fetch()
    off = plan.out_off
    plan.close()
    for k, s, l, d in zip(g["pair"], g["score"], g["out_len"], g["digest8"]):
        o = int(off[k])
        assert sc[k] == s and ln[k] == l, (k, sc[k], s)
        assert _digest8(out_a[o:o + l], out_b[o:o + l]) == d, k
    M = np.diff(ro)[ridx]
    N = np.diff(qo)
    assert (ln >= np.maximum(M, N)).all() and (ln <= M + N).all()
    for k in range(2000):
        o, l = int(off[k]), int(ln[k])
        a, b = out_a[o:o + l], out_b[o:o + l]
        assert not ((a == 45) & (b == 45)).any(), k
        assert (a != 45).sum() == M[k] and (b[b != 45] == qb[qo[k]:qo[k + 1]]).all(), k
        assert not out_a[o + l:int(off[k + 1])].any()
    # the one-shot call (slab pipeline, another launch shape) gives the same bytes on a subset
    sub = 300
    one = gpu_aligner.align_packed(rb, ro, ridx[:sub], qb[:qo[sub]], qo[:sub + 1], 15, 3, 1, 0)
    assert (one[4] == sc[:sub]).all() and (one[3] == ln[:sub]).all()
    assert (one[0] == out_a[:int(off[sub])]).all() and (one[1] == out_b[:int(off[sub])]).all()


def test_gpu_int16x2_admission_boundary_main_aligner(gpu_aligner, oracle_port):
    """The range proof of the int16x2 path (fits_int16, csrc/gotoh_b200.cu) on the GPU's own 16-bit arithmetic: for a
    fixed batch the gap-extension penalty is raised until the host stops admitting pairs to int16x2; the batches exactly
    at the last admitted value and one past it (and a mixed one in between) must all equal the oracle, and both kernels
    must really have run."""
    import random
    rng = random.Random(77)
    ref = "".join(rng.choice("ACGT") for _ in range(900))
    qs = []
    for _ in range(600):
        lo = rng.randrange(len(ref) - 260)
        q = list(ref[lo:lo + rng.choice([60, 128, 200, 251, 256])])
        for _ in range(4):
            q[rng.randrange(len(q))] = rng.choice("ACGT")
        qs.append("".join(q))
    rb, ro = packing.pack([ref])
    qb, qo = packing.pack(qs)
    ridx = np.zeros(len(qs), np.int32)

    def paths(gep):
        plan = gpu_aligner.plan(rb, ro, ridx, qb, qo, 10, gep, 1, 0)
        r = (plan.stat(5), plan.stat(6))
        plan.close()
        return r

    first_mixed = next(g for g in range(1, 200) if paths(g)[1] > 0)          # some pair no longer fits
    first_none = next(g for g in range(first_mixed, 400) if paths(g)[0] == 0)   # no pair fits
    assert paths(first_mixed - 1) == (len(qs), 0) and first_none > first_mixed
    for gep in (first_mixed - 1, first_mixed, (first_mixed + first_none) // 2, first_none - 1, first_none):
        x2, x1 = paths(gep)
        got = gpu_aligner.align_batch([ref], qs, 10, gep, 1, 0, ref_idx=ridx)
        for k, q in enumerate(qs):
            assert got[k] == oracle_port.align_it(ref, q, 10, gep, 1), (gep, k, x2, x1)
    x2, x1 = paths((first_mixed + first_none) // 2)
    assert x2 > 0 and x1 > 0, "the mixed batch must exercise both kernels in one plan"


def test_gpu_mixed_int16x2_and_int32_launches_share_a_plan(gpu_aligner, oracle_port):
    """GPU twin of test_emu_mixed_int16x2_and_int32_launches_share_a_plan, asserting that both paths ran."""
    import random
    rng = random.Random(2026)
    alpha = "ACGTNRYKMSWBDHVacgtnXx*.-Uu"
    refs, qs = [], []
    for _ in range(3000):
        a = "".join(rng.choice(alpha if rng.random() < 0.3 else "ACGT") for _ in range(rng.randint(1, 300)))
        if rng.random() < 0.6:
            lo = rng.randrange(len(a))
            b = list(a[lo:lo + rng.randint(1, 300)])
            for _ in range(rng.randint(0, 5)):
                b[rng.randrange(len(b))] = rng.choice(alpha)
            b = "".join(b)
        else:
            b = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 300)))
        refs.append(a)
        qs.append(b)
    rb, ro = packing.pack(refs)
    qb, qo = packing.pack(qs)
    for gip, gep, term in [(40, 10, 1), (25, 12, 0)]:
        plan = gpu_aligner.plan(rb, ro, None, qb, qo, gip, gep, term, 0)
        assert plan.stat(5) > 0 and plan.stat(6) > 0, (plan.stat(5), plan.stat(6))
        plan.run()
        out = plan.fetch()
        exp = oracle_port.align_batch(0, rb, ro, None, qb, qo, gip, gep, term)
        assert (out[2] == exp[3]).all() and (out[3] == exp[4]).all()
        assert (out[0] == exp[0]).all() and (out[1] == exp[1]).all()
        plan.close()


def test_gpu_multi_device_sharding_on_two_or_more_devices(gpu_aligner, oracle_port):
    """device_mask over every visible GPU (the library's own static sharding, one host thread per device) gives the
    same bytes as device 0 alone, for all three result forms.  Needs a multi-GPU box."""
    nd = gpu_aligner.device_count()
    if nd < 2:
        pytest.skip("only %d CUDA device visible: in-library sharding needs >= 2 (run on a multi-GPU lease; log in profiles/)" % nd)
    ref, qb, qo = workloads.c2_reads_packed(40000, seed=10)
    rb, ro = packing.pack([ref])
    ridx = np.zeros(40000, np.int32)
    allm = (1 << nd) - 1
    one = gpu_aligner.align_packed(rb, ro, ridx, qb, qo, 10, 3, 1, 0, device_mask=1)
    many = gpu_aligner.align_packed(rb, ro, ridx, qb, qo, 10, 3, 1, 0, device_mask=allm)
    assert (one[3] == many[3]).all() and (one[4] == many[4]).all()
    assert (one[0] == many[0]).all() and (one[1] == many[1]).all()
    t = gpu_aligner.align_packed_tight(rb, ro, ridx, qb, qo, 10, 3, 1, 0, device_mask=allm)
    c = gpu_aligner.align_packed_compact(rb, ro, ridx, qb, qo, 10, 3, 1, 0, device_mask=allm)
    assert (t[3] == one[3]).all() and (t[4] == one[4]).all() and (c.scores == one[4]).all() and (np.diff(t[2]) >= one[3][:-1]).all()
    for k in range(0, 40000, 37):
        o, l, to = int(one[2][k]), int(one[3][k]), int(t[2][k])
        assert (t[0][to:to + l] == one[0][o:o + l]).all() and (t[1][to:to + l] == one[1][o:o + l]).all()
        oa, ob = c.arrays(k)
        assert (oa == one[0][o:o + l]).all() and (ob == one[1][o:o + l]).all()
    for k in range(0, 40000, 4000):
        assert c[k] == oracle_port.align_it(ref, qb[qo[k]:qo[k + 1]].tobytes().decode(), 10, 3, 1)

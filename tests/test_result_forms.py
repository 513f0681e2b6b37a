"""The three result forms of the batched call (include/gotoh_b200.h): strided strings (the reference's two strings at
caller-chosen offsets), tight strings (no stride tails cross PCIe) and compact (record + 2-bit op script, rendered on
demand by gotoh_b200/compact.py).  All three must agree byte for byte with each other and with the oracle
(gotoh.cpp:418-513: end cell, traceback, overhangs, reversal).  CPU: the kernel sources under the SIMT emulator;
-m gpu: the product library on C2 / C3 / C4 shapes."""
import random
import threading

import numpy as np
import pytest

from gotoh_b200 import _ffi, packing, workloads


def _mixed_batch(seed, n):
    """nt pairs of every flavour: prefixes/suffixes/infixes of the reference (both overhang cases, gotoh.cpp:429-449),
    unrelated pairs, multi-strip queries, whitespace to trim."""
    rng = random.Random(seed)
    refs, qs = [], []
    for k in range(n):
        a = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 420)))
        r = rng.random()
        if r < 0.5:
            lo = rng.randrange(len(a))
            b = list(a[lo:lo + rng.randint(1, 330)])
            for _ in range(rng.randint(0, 4)):
                b[rng.randrange(len(b))] = rng.choice("ACGTN-")
            b = "".join(b)
        elif r < 0.7:
            b = a + "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 60)))     # query overhangs the reference
        else:
            b = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 300)))
        if k % 7 == 0:
            a, b = "  " + a + "\n", "\t" + b + " \r\n"
        refs.append(a)
        qs.append(b)
    return refs, qs


def _check_forms(aligner, oracle, matrix, refs, qs, gip, gep, term, ref_idx=None, n_oracle=None, device_mask=1):
    rb, ro = packing.pack(refs)
    qb, qo = packing.pack(qs)
    ridx = None if ref_idx is None else np.asarray(ref_idx, np.int32)
    s = aligner.align_packed(rb, ro, ridx, qb, qo, gip, gep, term, matrix, device_mask=device_mask)
    t = aligner.align_packed_tight(rb, ro, ridx, qb, qo, gip, gep, term, matrix, device_mask=device_mask)
    c = aligner.align_packed_compact(rb, ro, ridx, qb, qo, gip, gep, term, matrix, device_mask=device_mask)
    n = len(qs)
    assert (t[3] == s[3]).all() and (t[4] == s[4]).all()
    assert (c.scores == s[4]).all() and (c.out_len == s[3]).all()
    # tight: offsets are the running sum of the lengths (one device) and the strings equal the strided ones
    if device_mask == 1:
        assert t[2][0] == 0 and (np.diff(t[2]) == s[3][:-1]).all()
    a = packing.unpack(s[0], s[2], s[3])
    b = packing.unpack(s[1], s[2], s[3])
    ta = packing.unpack(t[0], t[2], t[3])
    tb = packing.unpack(t[1], t[2], t[3])
    assert ta == a and tb == b
    for k in range(n):
        assert c.strings(k) == (a[k], b[k]), k
    fn = {0: oracle.align_it, 1: oracle.align_it_aa}[matrix]
    step = max(1, n // (n_oracle or n))
    for k in range(0, n, step):
        r = refs[k if ref_idx is None else ref_idx[k]]
        assert c[k] == fn(r, qs[k], gip, gep, term), k
    return s, t, c


@pytest.mark.parametrize("gip,gep,term", [(10, 3, 1), (5, 1, 0)])
def test_emu_result_forms_agree(emu_aligner, oracle_port, monkeypatch, gip, gep, term):
    monkeypatch.setenv("GOTOH_B200_SLAB_MB", "1")          # several slabs: the collector assigns positions slab by slab
    refs, qs = _mixed_batch(11, 120)
    _, _, c = _check_forms(emu_aligner, oracle_port, 0, refs, qs, gip, gep, term)
    assert c.nbytes() < 0.2 * sum(len(a) + len(b) for a, b in zip(refs, qs)) * 2


def test_emu_result_forms_amino_and_degap(emu_aligner, oracle_port):
    refs, qs = workloads.c3_queries(90, seed=5)
    _check_forms(emu_aligner, oracle_port, 1, refs, qs, 40, 10, 0, ref_idx=[k % 3 for k in range(90)])
    # align_it_aa_rb: inputs are degapped before the alignment, the rendering must degap too
    got = emu_aligner.align_batch(["K-F-RWW", "AC-D"], ["KF--GR", "A-CD"], 4, 2, 0, _ffi.AA_RB, compact=True)
    assert got.strings(0) == oracle_port.align_it_aa_rb("K-F-RWW", "KF--GR", 4, 2)
    assert got.strings(1) == oracle_port.align_it_aa_rb("AC-D", "A-CD", 4, 2)


def test_emu_compact_capacity_error_and_retry(emu_aligner, monkeypatch):
    monkeypatch.setenv("GOTOH_B200_SLAB_MB", "1")
    refs, qs = _mixed_batch(3, 60)
    rb, ro = packing.pack(refs)
    qb, qo = packing.pack(qs)
    n = len(qs)
    rec, off = np.zeros(8 * n, np.int32), np.zeros(n, np.int64)
    with pytest.raises(_ffi.GotohCapacityError):
        emu_aligner.align_packed_compact(rb, ro, None, qb, qo, 10, 3, 1, 0, out=(rec, np.zeros(40, np.uint32), off))
    with pytest.raises(_ffi.GotohCapacityError):
        emu_aligner.align_packed_tight(rb, ro, None, qb, qo, 10, 3, 1, 0,
                                       out=(np.zeros(900, np.uint8), np.zeros(900, np.uint8), off, np.zeros(n, np.int32), np.zeros(n, np.int32)))
    # the default call sizes the op buffer for typical alignments and falls back to the worst-case bound
    c = emu_aligner.align_packed_compact(rb, ro, None, qb, qo, 10, 3, 1, 0)
    s = emu_aligner.align_packed(rb, ro, None, qb, qo, 10, 3, 1, 0)
    assert (c.scores == s[4]).all()


@pytest.mark.parametrize("mode", ["strided", "tight", "compact"])
def test_emu_failed_result_copy_does_not_hang(emu_aligner, monkeypatch, mode):
    """a builder whose result copy failed used to return without waking the other builder (lost wakeup ->
    gotoh_b200_align_batch hung in join()).  An injected failure on slab 1 of a many-slab call must come back as an
    error from every result form."""
    monkeypatch.setenv("GOTOH_B200_SLAB_MB", "1")
    monkeypatch.setenv("GOTOH_B200_TEST_FAIL_FETCH", "1")
    refs, qs = _mixed_batch(4, 200)
    rb, ro = packing.pack(refs)
    qb, qo = packing.pack(qs)
    fn = {"strided": emu_aligner.align_packed, "tight": emu_aligner.align_packed_tight, "compact": emu_aligner.align_packed_compact}[mode]
    res = {}

    def call():
        try:
            fn(rb, ro, None, qb, qo, 10, 3, 1, 0)
            res["r"] = "returned"
        except _ffi.GotohError as e:
            res["r"] = e

    th = threading.Thread(target=call, daemon=True)
    th.start()
    th.join(120)
    assert not th.is_alive(), "the call hangs after a failed result copy"
    assert isinstance(res["r"], _ffi.GotohError) and "injected" in str(res["r"])


def test_emu_multi_device_sharding_all_forms(emu_aligner, oracle_port, monkeypatch):
    """The library's own static sharding (device_mask, one host thread per device, capacity slices for the tight and
    compact forms) on three emulated devices: same bytes as one device, for every result form."""
    monkeypatch.setenv("SIMT_EMU_DEVICES", "3")
    monkeypatch.setenv("GOTOH_B200_SLAB_MB", "1")
    assert emu_aligner.device_count() == 3
    refs, qs = _mixed_batch(17, 150)
    s, t, c = _check_forms(emu_aligner, oracle_port, 0, refs, qs, 10, 3, 1, n_oracle=30, device_mask=0b111)
    one = emu_aligner.align_packed(*packing.pack(refs), None, *packing.pack(qs), 10, 3, 1, 0, device_mask=1)
    assert (one[0] == s[0]).all() and (one[1] == s[1]).all() and (one[3] == s[3]).all() and (one[4] == s[4]).all()
    assert (np.diff(t[2]) >= s[3][:-1]).all()                     # increasing, gaps only at device boundaries
    # fewer pairs than devices, and a device index that is not there
    got = emu_aligner.align_batch(refs[:2], qs[:2], 10, 3, 1, 0, devices=[0, 1, 2], compact=True)
    assert [got[k] for k in range(2)] == [oracle_port.align_it(refs[k], qs[k], 10, 3, 1) for k in range(2)]
    with pytest.raises(_ffi.GotohError):
        emu_aligner.align_batch(refs[:2], qs[:2], 10, 3, 1, 0, devices=[5])


def test_emu_device_index_out_of_range_is_rejected(emu_aligner):
    with pytest.raises(ValueError):
        emu_aligner.align_batch(["ACGT"], ["ACG"], 5, 1, 1, 0, devices=[32])


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_result_forms_c2_c3_c4_shapes(gpu_aligner, oracle_port, monkeypatch):
    """compact -> strings and tight strings equal the strided strings on every pair, and the oracle on a sample, for the
    three benchmark shapes (C2 reads vs HXB2 pol, C3 aa windows, C4 9.6 kb genome pairs) and a mixed fuzz batch."""
    refs, qs = _mixed_batch(21, 4000)
    _check_forms(gpu_aligner, oracle_port, 0, refs, qs, 10, 3, 1, n_oracle=1500)
    _check_forms(gpu_aligner, oracle_port, 0, refs, qs, 10, 10, 0, n_oracle=500)
    ref, reads = workloads.c2_reads(30000, seed=12)            # 4 slabs of the one-shot pipeline
    _check_forms(gpu_aligner, oracle_port, 0, [ref], reads, 10, 3, 1, ref_idx=[0] * len(reads), n_oracle=300)
    arefs, aq = workloads.c3_queries(60000, seed=13)
    _check_forms(gpu_aligner, oracle_port, 1, arefs, aq, 40, 10, 1, ref_idx=[k % 3 for k in range(len(aq))], n_oracle=3000)
    seeds, ridx, qb, qo = workloads.c4_pairs_packed(10, seed=14)
    _check_forms(gpu_aligner, oracle_port, 0, seeds, workloads.unpacked(qb, qo), 15, 3, 1, ref_idx=[int(x) for x in ridx])


@pytest.mark.gpu
def test_gpu_compact_full_size_c2(gpu_aligner, oracle_port):
    """BASELINE configs[1] at full size through the compact call: scores and lengths equal the string call's on all
    1,000,000 reads, rendered strings equal it on a sample, and the bytes that cross PCIe are < 2 % of the strings'."""
    n = 1000000
    ref, qb, qo = workloads.c2_reads_packed(n, seed=77)
    rb, ro = packing.pack([ref])
    ridx = np.zeros(n, np.int32)
    s = gpu_aligner.align_packed(rb, ro, ridx, qb, qo, 10, 3, 1, 0)
    c = gpu_aligner.align_packed_compact(rb, ro, ridx, qb, qo, 10, 3, 1, 0)
    assert (c.scores == s[4]).all() and (c.out_len == s[3]).all()
    assert c.nbytes() < 0.02 * 2 * int(s[2][-1])
    for k in range(0, n, 499):
        o, ln = int(s[2][k]), int(s[3][k])
        oa, ob = c.arrays(k)
        assert (oa == s[0][o:o + ln]).all() and (ob == s[1][o:o + ln]).all(), k
    for k in range(0, n, n // 200):
        assert c[k] == oracle_port.align_it(ref, qb[qo[k]:qo[k + 1]].tobytes().decode(), 10, 3, 1)
    t = gpu_aligner.align_packed_tight(rb, ro, ridx, qb, qo, 10, 3, 1, 0)
    assert (t[3] == s[3]).all() and (t[4] == s[4]).all() and t[2][0] == 0 and (np.diff(t[2]) == s[3][:-1]).all()
    for k in range(0, n, 997):
        o, ln, to = int(s[2][k]), int(s[3][k]), int(t[2][k])
        assert (t[0][to:to + ln] == s[0][o:o + ln]).all() and (t[1][to:to + ln] == s[1][o:o + ln]).all(), k

"""NEXT #1: the live aligner gotoh2.Aligner.align on the GPU, checked against golden
vectors produced by the reference's own _gotoh2.c (incl. the reference's unit-test answers) and
against the oracle.  The emu_* tests run the kernel sources under the CPU SIMT emulator (not gpu);
the gpu_* tests run the product library on a B200."""
import hashlib
import random

import pytest

from conftest import load_golden


def _check_golden(make_aligner, max_cells):
    doc = load_golden("gotoh2")
    groups = {}
    for c in doc["kats"] + doc["fuzz"]:
        if len(c["a"]) * len(c["b"]) > max_cells:
            continue
        groups.setdefault((c["gop"], c["gep"], c["is_global"], c["model"]), []).append(c)
    n, bad = 0, []
    for (gop, gep, glob, model), lst in groups.items():
        al = make_aligner(gop, gep, glob, model)
        try:
            out = al.align_batch([(c["a"], c["b"]) for c in lst])
        except RuntimeError:
            out = []
            for c in lst:
                try:
                    out.append(al.align(c["a"], c["b"]))
                except RuntimeError:
                    out.append("traceback")
        for c, r in zip(lst, out):
            n += 1
            if "error" in c:
                ok = r == "traceback"
            elif r == "traceback":
                ok = False
            elif "out_a" in c:
                ok = r == (c["out_a"], c["out_b"], c["score"])
            else:
                ok = (r[2] == c["score"] and hashlib.sha256(r[0].encode()).hexdigest() == c["sha_a"] and
                      hashlib.sha256(r[1].encode()).hexdigest() == c["sha_b"])
            if not ok:
                bad.append(c["name"])
    return n, bad


def _long_pairs(seed):
    rng = random.Random(seed)
    pairs = []
    shapes = [(300, 600), (100, 1000), (700, 255), (64, 256), (513, 257), (40, 1300), (900, 300), (255, 512)]
    # full-width single strips (8 columns per lane) and short, wide grids: the byte-packed planes of the reverse sweep
    shapes += [(rng.randint(20, 90), rng.randint(200, 256)) for _ in range(24)] + [(rng.randint(30, 60), rng.randint(257, 700)) for _ in range(8)]
    for l1, l2 in shapes:
        a = "".join(rng.choice("ACGT") for _ in range(l1))
        b = list((a * (l2 // l1 + 2))[:l2])
        for _ in range(l2 // 12):
            b[rng.randrange(l2)] = rng.choice("ACGTN")
        for _ in range(3):
            p = rng.randrange(len(b))
            del b[p:p + rng.randint(1, 9)]
        pairs.append((a, "".join(b)))
    return pairs


@pytest.fixture(scope="module")
def oracle2_port():
    from oracle.oracle2 import Oracle2
    return Oracle2("port")


def test_emu_gotoh2_golden(emu_aligner):
    from gotoh_b200.gotoh2 import Aligner
    n, bad = _check_golden(lambda *a: Aligner(*a, library=emu_aligner._libobj), max_cells=3e5)
    assert n >= 1500 and not bad, bad[:5]


def test_emu_gotoh2_multi_strip_and_interface(emu_aligner, oracle2_port):
    from gotoh_b200.gotoh2 import Aligner
    pairs = _long_pairs(4)
    for gop, gep, glob in [(15, 3, True), (5, 1, False), (0, 0, True)]:
        al = Aligner(gop, gep, glob, "HYPHY_NUC", library=emu_aligner._libobj)
        out = al.align_batch(pairs)
        for (a, b), o in zip(pairs, out):
            assert o == oracle2_port.align(a, b, gop, gep, glob, "HYPHY_NUC")
    al = Aligner(library=emu_aligner._libobj)                 # defaults of gotoh2.py:8
    assert (al.gap_open_penalty, al.gap_extend_penalty, al.is_global, al.alphabet) == (10, 1, False, "ACGT?")
    assert al.clean_sequence("acgtn-x?") == "ACGT????"        # gotoh2.py:70-72
    al.gap_open_penalty = 5
    al.is_global = True
    assert al.align("ACGT", "ACT") == ("ACGT", "AC-T", 9)     # alignment/tests/test.py:174-182
    al.is_global = False
    assert al.align("TACGTA", "ACGT") == ("TACGTA", "-ACGT-", 20)   # test.py:193-203
    with pytest.raises(AssertionError):
        al.align("", "ACGT")                                  # gotoh2.py:84
    with pytest.raises(AssertionError):
        al.align(b"ACGT", "ACGT")                             # gotoh2.py:82


def test_emu_gotoh2_general_kernels_still_match(emu_aligner, monkeypatch):
    """GOTOH_B200_GOTOH2=general pins the un-tuned kernels (the path negative penalties take); same answers."""
    from gotoh_b200.gotoh2 import Aligner
    monkeypatch.setenv("GOTOH_B200_GOTOH2", "general")
    n, bad = _check_golden(lambda *a: Aligner(*a, library=emu_aligner._libobj), max_cells=2e3)
    assert n >= 300 and not bad, bad[:5]


def _couple_cases(n_reads, seed):
    """Many second sequences of ragged length against a shared first sequence: the int16x2 forward kernel couples
    them two per warp (k2f_x2); widths cover every K group, the last read of a group stays single."""
    from gotoh_b200 import workloads
    rng = random.Random(seed)
    ref, reads = workloads.c2_reads(n_reads, seed=seed)
    sub = ref[500:1100]
    nt = [(sub, r[:rng.choice([251, 250, 200, 130, 97, 64, 33, 17, 5, 1])]) for r in reads]
    nt += [(ref[1000:1300], r[:rng.randint(1, 251)]) for r in reads[:n_reads // 3]]
    refs, qs = workloads.c3_queries(n_reads, seed=seed + 1)
    aa = [(refs[k % 3], q[:rng.randint(1, len(q))]) for k, q in enumerate(qs)]
    return ref, reads, nt, aa


def _x2_tasks(lib):
    import ctypes
    stats = (ctypes.c_double * 11)()
    assert lib.lib.gotoh_b200_gotoh2_last_stats(stats, 11) == 11
    return int(stats[10])


def _check_couples(lib, oracle2, monkeypatch, n_reads, n_full):
    from gotoh_b200.gotoh2 import Aligner
    ref, reads, nt, aa = _couple_cases(n_reads, 3)
    jobs = [(nt, g, e, glob, "HYPHY_NUC") for g, e, glob in [(10, 3, False), (15, 3, True), (0, 0, True), (2, 1, False), (0, 1, True)]]
    jobs += [(aa, g, e, glob, "EmpHIV25") for g, e, glob in [(40, 10, False), (40, 10, True), (3, 1, False)]]
    # full-length first sequence (3039 rows): ordinary read settings, penalties close to the 16-bit range limit
    # (admitted), and just past it (must fall back to the int32 forward kernel)
    full = [(ref, r) for r in reads[:n_full]]
    jobs += [(full, 10, 3, False, "HYPHY_NUC"), (full, 10, 9, True, "HYPHY_NUC"), (full, 300, 8, False, "HYPHY_NUC")]
    for pairs, gop, gep, glob, model in jobs:
        al = Aligner(gop, gep, glob, model, library=lib)
        out = al.align_batch(pairs)
        assert _x2_tasks(lib) >= len(pairs) // 2 - 8, (gop, gep, glob, model)
        for (a, b), o in zip(pairs, out):
            assert o == oracle2.align(a, b, gop, gep, glob, model), (gop, gep, glob, model, b)
    al = Aligner(10, 10, False, "HYPHY_NUC", library=lib)
    out = al.align_batch(full)
    assert _x2_tasks(lib) == 0
    for (a, b), o in zip(full, out):
        assert o == oracle2.align(a, b, 10, 10, False, "HYPHY_NUC")
    # GOTOH_B200_GOTOH2=r1 pins the one-pair-per-warp reverse sweep (K <= 3 couples otherwise share a warp, k2r_x2)
    monkeypatch.setenv("GOTOH_B200_GOTOH2", "r1")
    for pairs, gop, gep, glob, model in [(nt, 10, 3, False, "HYPHY_NUC"), (aa, 40, 10, False, "EmpHIV25"), (aa, 40, 10, True, "EmpHIV25")]:
        al = Aligner(gop, gep, glob, model, library=lib)
        for (a, b), o in zip(pairs, al.align_batch(pairs)):
            assert o == oracle2.align(a, b, gop, gep, glob, model), (gop, gep, glob, model, b)
    # GOTOH_B200_GOTOH2=x1 pins the int32 forward kernel: same answers, no couples
    monkeypatch.setenv("GOTOH_B200_GOTOH2", "x1")
    al = Aligner(10, 3, False, "HYPHY_NUC", library=lib)
    out = al.align_batch(nt)
    assert _x2_tasks(lib) == 0
    for (a, b), o in zip(nt, out):
        assert o == oracle2.align(a, b, 10, 3, False, "HYPHY_NUC")


def test_emu_gotoh2_int16x2_couples(emu_aligner, oracle2_port, monkeypatch):
    _check_couples(emu_aligner._libobj, oracle2_port, monkeypatch, n_reads=36, n_full=4)


@pytest.mark.gpu
def test_gpu_gotoh2_int16x2_couples(gpu_aligner, oracle2_port, monkeypatch):
    _check_couples(gpu_aligner._libobj, oracle2_port, monkeypatch, n_reads=900, n_full=40)


@pytest.mark.gpu
def test_gpu_gotoh2_golden_incl_reference_unit_tests(gpu_aligner):
    from gotoh_b200.gotoh2 import Aligner
    n, bad = _check_golden(lambda *a: Aligner(*a, library=gpu_aligner._libobj), max_cells=1e12)
    assert n >= 1515 and not bad, bad[:5]


@pytest.mark.gpu
def test_gpu_gotoh2_shapes_vs_oracle(gpu_aligner, oracle2_port):
    """remap.py:33 settings (15,3,global,HYPHY_NUC) on ~2 kb pairs, aln2counts.py:34-37 settings
    (40,10,local,EmpHIV25) on amino-acid windows, reads vs the HXB2 pol seed in local mode."""
    from gotoh_b200 import workloads
    from gotoh_b200.gotoh2 import Aligner
    rng = random.Random(6)
    lib = gpu_aligner._libobj
    pairs = _long_pairs(9)
    seeds = workloads.hcv_seeds()
    pairs += [(seeds[0][:2200], seeds[7][100:2300]), (seeds[3][4000:6100], seeds[3][4050:6000])]
    for gop, gep, glob in [(15, 3, True), (5, 1, False)]:
        al = Aligner(gop, gep, glob, "HYPHY_NUC", library=lib)
        for (a, b), o in zip(pairs, al.align_batch(pairs)):
            assert o == oracle2_port.align(a, b, gop, gep, glob, "HYPHY_NUC")
    refs, qs = workloads.c3_queries(600, seed=5)
    aa = [(refs[k % 3], q) for k, q in enumerate(qs)]
    al = Aligner(40, 10, False, "EmpHIV25", library=lib)
    for (a, b), o in zip(aa, al.align_batch(aa)):
        assert o == oracle2_port.align(a, b, 40, 10, False, "EmpHIV25")
    ref, reads = workloads.c2_reads(150, seed=8)
    rp = [(ref, r) for r in reads]
    al = Aligner(10, 3, False, "HYPHY_NUC", library=lib)
    for (a, b), o in zip(rp, al.align_batch(rp)):
        assert o == oracle2_port.align(a, b, 10, 3, False, "HYPHY_NUC")
    del rng


@pytest.mark.gpu
def test_gpu_gotoh2_large_batch_properties(gpu_aligner, oracle2_port):
    """50,000 reads against the HXB2 pol seed through Aligner.align_batch (local mode, the aln2counts/remap style of
    call): on every pair the two aligned strings have equal length, no column pairs two gaps and removing the gaps
    gives back the cleaned inputs; 250 pairs are compared with the oracle."""
    from gotoh_b200 import workloads
    from gotoh_b200.gotoh2 import Aligner
    ref, reads = workloads.c2_reads(50000, seed=21)
    al = Aligner(10, 3, False, "HYPHY_NUC", library=gpu_aligner._libobj)
    out = al.align_batch([(ref, r) for r in reads])
    cref = al.clean_sequence(ref)
    for k, (a, b, score) in enumerate(out):
        assert len(a) == len(b)
        assert a.replace("-", "") == cref and b.replace("-", "") == al.clean_sequence(reads[k])
        assert not any(x == "-" and y == "-" for x, y in zip(a[:40], b[:40]))
        if k % 200 == 0:
            assert (a, b, score) == oracle2_port.align(ref, reads[k], 10, 3, False, "HYPHY_NUC")

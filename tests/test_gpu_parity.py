"""Parity tests proper: the product library (nvcc, sm_100a) on a real B200, through the C ABI,
against the golden vectors generated from the reference and against the oracle on seeded
inputs.  Bit-exact: score and both aligned strings.  Nothing here reads /root/reference."""
import random

import numpy as np
import pytest

from conftest import load_golden, run_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["auto", "32", "full"])
def forced_path(request, monkeypatch):
    """auto: int16x2 path wherever the range proof allows (two 16-lane wavefronts per warp for queries <= 128 wide);
    32: everything on the int32 path; full: int16x2 on 32-lane wavefronts only (GOTOH_B200_HALF=0)."""
    monkeypatch.delenv("GOTOH_B200_FORCE_PATH", raising=False)
    monkeypatch.delenv("GOTOH_B200_HALF", raising=False)
    if request.param == "32":
        monkeypatch.setenv("GOTOH_B200_FORCE_PATH", "32")
    elif request.param == "full":
        monkeypatch.setenv("GOTOH_B200_HALF", "0")
    return request.param


def test_gpu_golden_vectors(gpu_aligner, forced_path):
    """Appendix B KATs, 1500 fuzz cases, benchmark shapes incl. 9.6 kb HCV genome pairs."""
    for name, least in (("appendix_b", 35), ("fuzz_small", 1400), ("shapes", 120)):
        n, bad = run_cases(gpu_aligner, load_golden(name)["cases"])
        assert n >= least and not bad, "%s: %d/%d differ, first %r" % (name, len(bad), n, bad[0][0])


def test_gpu_drop_in_functions(gpu_aligner, oracle_port):
    assert gpu_aligner.align_it("ACGT", "ACT", 5, 1, 1) == ("ACGT", "AC-T", 9)
    assert gpu_aligner.align_it("TACGTA", "ACGT", 5, 1, 0) == ("TACGTA", "-ACGT-", 26)
    assert gpu_aligner.align_it_aa("WWWWKFR", "KFR", 40, 10, 0) == ("WWWWKFR", "----KFR", 104)
    assert gpu_aligner.align_it_aa_rb("K-F-R", "KF--GR", 4, 2) == oracle_port.align_it_aa_rb("K-F-R", "KF--GR", 4, 2)
    import gotoh_b200
    import gotoh
    assert gotoh.align_it("  ACGT\n", "\tACT \r\n", 5, 1, 1) == ("ACGT", "AC-T", 9)
    assert gotoh_b200.align_it_aa("ERM", "ERM", 40, 10, 1) == ("ERM", "ERM", 24)


def _check_packed(aligner, oracle, matrix, rb, ro, ridx, qb, qo, gip, gep, term, **kw):
    got = aligner.align_packed(rb, ro, ridx, qb, qo, gip, gep, term, matrix, **kw)
    exp = oracle.align_batch(matrix, rb, ro, ridx, qb, qo, gip, gep, term)
    assert (got[3] == exp[3]).all(), "aligned lengths differ at %r" % np.nonzero(got[3] != exp[3])[0][:5]
    assert (got[4] == exp[4]).all(), "scores differ at %r" % np.nonzero(got[4] != exp[4])[0][:5]
    off, ln = exp[2], exp[3]
    mask = (np.arange(len(got[0]))[None, :] < 0)  # placeholder to keep numpy import used
    del mask
    for k in range(len(ln)):
        s, e = int(off[k]), int(off[k]) + int(ln[k])
        assert (got[0][s:e] == exp[0][s:e]).all() and (got[1][s:e] == exp[1][s:e]).all(), "pair %d strings differ" % k
    return got


@pytest.mark.parametrize("gip,gep,term", [(10, 3, 1), (10, 10, 0)])
def test_gpu_c2_reads_vs_oracle(gpu_aligner, oracle_port, gip, gep, term, forced_path):
    """C2 shape: 251-nt reads vs the 3039-nt HXB2 pol seed; both parameter sweeps of SURVEY 8d."""
    from gotoh_b200 import packing, workloads
    ref, qb, qo = workloads.c2_reads_packed(1500, seed=42)
    rb, ro = packing.pack([ref])
    _check_packed(gpu_aligner, oracle_port, 0, rb, ro, np.zeros(1500, np.int32), qb, qo, gip, gep, term)


@pytest.mark.parametrize("term", [0, 1])
def test_gpu_c3_amino_vs_oracle(gpu_aligner, oracle_port, term, forced_path):
    from gotoh_b200 import packing, workloads
    refs, ridx, qb, qo = workloads.c3_queries_packed(30000, seed=43)
    rb, ro = packing.pack(refs)
    _check_packed(gpu_aligner, oracle_port, 1, rb, ro, ridx, qb, qo, 40, 10, term)


@pytest.mark.parametrize("long_mode", ["flow", "cta", "warp"])
def test_gpu_c4_long_pairs_vs_oracle(gpu_aligner, oracle_port, monkeypatch, long_mode):
    """C4 shape: ~9.6 kb x ~9.6 kb, 38 strips of 256 columns per pair, IUPAC codes in the refs; both
    long-pair kernels (CTA wavefront / warp-serial strips)."""
    from gotoh_b200 import packing, workloads
    monkeypatch.setenv("GOTOH_B200_LONG", long_mode)
    refs, ridx, qb, qo = workloads.c4_pairs_packed(12, seed=44)
    rb, ro = packing.pack(refs)
    _check_packed(gpu_aligner, oracle_port, 0, rb, ro, ridx, qb, qo, 15, 3, 1)
    # odd shapes: few rows / many strips, strip counts around multiples of the 4 warps
    rng = random.Random(5)
    refs, qs = [], []
    for M, N in [(40, 5000), (3000, 257), (1500, 1300), (700, 1025), (2049, 2049), (257, 4097), (5000, 300)]:
        a = "".join(rng.choice("ACGT") for _ in range(M))
        b = list((a * (N // M + 2))[:N])
        for _ in range(N // 15):
            b[rng.randrange(N)] = rng.choice("ACGTN")
        refs.append(a)
        qs.append("".join(b))
    rb, ro = packing.pack(refs)
    qb, qo = packing.pack(qs)
    _check_packed(gpu_aligner, oracle_port, 0, rb, ro, None, qb, qo, 15, 3, 1)


def test_gpu_random_fuzz_batch(gpu_aligner, oracle_port, forced_path):
    """20k unrelated/related pairs, full byte alphabet of the nt table, all penalty combinations."""
    from gotoh_b200 import packing
    rng = random.Random(2026)
    alpha = "ACGTNRYKMSWBDHVacgtnXx*.-Uu"
    for gip, gep, term in [(0, 0, 1), (0, 1, 0), (1, 0, 0), (5, 1, 1), (10, 3, 0), (40, 10, 1), (15, 3, 1)]:
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
        _check_packed(gpu_aligner, oracle_port, 0, rb, ro, None, qb, qo, gip, gep, term)


def test_gpu_half_warp_wavefronts(gpu_aligner, oracle_port, monkeypatch):
    """Queries <= 128 wide run as two 16-lane wavefronts per warp (four alignments): ragged widths, uneven reference
    shares (couple + filler warps), chunked arenas; checked against the oracle, and 200k C3 windows against the
    32-lane kernels (GOTOH_B200_HALF=0), an independent implementation of the same cells."""
    from test_emu_parity import _half_warp_batch
    from gotoh_b200 import packing, workloads
    pairs = _half_warp_batch(4001, 33)
    for arena_mb in (None, "2"):
        if arena_mb:
            monkeypatch.setenv("GOTOH_B200_ARENA_MB", arena_mb)
        for gip, gep, term in [(40, 10, 1), (3, 1, 0)]:
            got = gpu_aligner.align_batch([a for a, _ in pairs], [b for _, b in pairs], gip, gep, term, 1)
            for (a, b), g in zip(pairs, got):
                assert g == oracle_port.align_it_aa(a, b, gip, gep, term), (a, b, gip, gep, term)
    monkeypatch.delenv("GOTOH_B200_ARENA_MB", raising=False)
    refs, ridx, qb, qo = workloads.c3_queries_packed(200000, seed=9)
    rb, ro = packing.pack(refs)
    a = gpu_aligner.align_packed(rb, ro, ridx, qb, qo, 40, 10, 1, 1)
    monkeypatch.setenv("GOTOH_B200_HALF", "0")
    b = gpu_aligner.align_packed(rb, ro, ridx, qb, qo, 40, 10, 1, 1)
    assert (a[3] == b[3]).all() and (a[4] == b[4]).all() and (a[0] == b[0]).all() and (a[1] == b[1]).all()


def test_gpu_paths_agree_and_order_invariance_at_scale(gpu_aligner, monkeypatch):
    """Size-independent properties on 200k C2 reads (too many for the CPU oracle in a test):
    the int16x2 and int32 kernels are independent implementations and must agree bit for bit;
    results do not depend on batch order / warp pairing; degapped outputs reproduce the inputs."""
    from gotoh_b200 import packing, workloads
    n = 200000
    ref, qb, qo = workloads.c2_reads_packed(n, seed=77)
    rb, ro = packing.pack([ref])
    ridx = np.zeros(n, np.int32)
    monkeypatch.delenv("GOTOH_B200_FORCE_PATH", raising=False)
    a = gpu_aligner.align_packed(rb, ro, ridx, qb, qo, 10, 3, 1, 0)
    monkeypatch.setenv("GOTOH_B200_FORCE_PATH", "32")
    b = gpu_aligner.align_packed(rb, ro, ridx, qb, qo, 10, 3, 1, 0)
    monkeypatch.delenv("GOTOH_B200_FORCE_PATH", raising=False)
    assert (a[3] == b[3]).all() and (a[4] == b[4]).all()
    off, ln = a[2], a[3]
    valid = (np.arange(int(off[-1])) - np.repeat(off[:-1], np.diff(off))) < np.repeat(ln, np.diff(off))
    assert (a[0][valid] == b[0][valid]).all() and (a[1][valid] == b[1][valid]).all()
    # reversed batch order -> same per-pair results
    qs = workloads.unpacked(qb[:qo[2000]], qo[:2001])
    rq, rqo = packing.pack(qs[::-1])
    c = gpu_aligner.align_packed(rb, ro, np.zeros(2000, np.int32), rq, rqo, 10, 3, 1, 0)
    assert (c[4][::-1] == a[4][:2000]).all() and (c[3][::-1] == a[3][:2000]).all()
    # degapped aligned strings are the (trimmed) inputs; no column is gap/gap
    for k in range(0, n, 997):
        s, e = int(off[k]), int(off[k]) + int(ln[k])
        ar, aq = a[0][s:e], a[1][s:e]
        assert ar[ar != 45].tobytes().decode() == ref
        assert (aq[aq != 45] == qb[qo[k]:qo[k + 1]]).all()
        assert not ((ar == 45) & (aq == 45)).any()


def test_gpu_stop_codon_bonus_and_many_classes(gpu_aligner, oracle_port, forced_path):
    rng = random.Random(32)
    refs, qs = [], []
    for _ in range(2000):
        a = "".join(rng.choice("ACGT") for _ in range(rng.randint(3, 200)))
        for _ in range(rng.randint(1, 3)):
            p = rng.randrange(len(a) + 1)
            a = a[:p] + rng.choice(["$$$", "$$$$", "$$", "$$$$$$"]) + a[p:]
        b = list(a.replace("$$$", rng.choice(["TAG", "TAA", "TGA", "TGG"])).replace("$", "A"))
        for _ in range(rng.randint(0, 3)):
            b[rng.randrange(len(b))] = rng.choice("ACGTTAG")
        refs.append(a)
        qs.append("".join(b))
    from gotoh_b200 import packing
    rb, ro = packing.pack(refs)
    qb, qo = packing.pack(qs)
    _check_packed(gpu_aligner, oracle_port, 0, rb, ro, None, qb, qo, 10, 3, 1)
    alpha = [chr(c) for c in range(1, 127) if chr(c) not in " \t\n\r$"]
    ref = "".join(alpha) + "".join(rng.choice(alpha) for _ in range(300))
    qs = ["".join(rng.choice(alpha) for _ in range(rng.randint(5, 400))) for _ in range(200)]
    got = gpu_aligner.align_batch(ref, qs, 10, 3, 1, 0)
    for q, g in zip(qs, got):
        assert g == oracle_port.align_it(ref, q, 10, 3, 1)


def test_gpu_chunked_arena(gpu_aligner, oracle_port, monkeypatch):
    from gotoh_b200 import packing, workloads
    monkeypatch.setenv("GOTOH_B200_ARENA_MB", "8")
    ref, qb, qo = workloads.c2_reads_packed(400, seed=9)
    rb, ro = packing.pack([ref])
    _check_packed(gpu_aligner, oracle_port, 0, rb, ro, np.zeros(400, np.int32), qb, qo, 10, 3, 1)


def test_gpu_int_peak_microbenchmarks_run(gpu_aligner):
    v = gpu_aligner.int_peak(2)
    assert v > 100.0   # G thread-instructions/s; a B200 does thousands


def _check_chunk_properties(out_a, out_b, out_off, out_len, M, N, c0, c1):
    """Pairs c0..c1: gaps removed, the aligned strings hold M reference and N query characters; no column pairs two
    gaps; the bytes between out_len and the stride are zero.  Vectorised over one chunk of the packed outputs."""
    base = int(out_off[c0])
    a = out_a[base:int(out_off[c1])]
    b = out_b[base:int(out_off[c1])]
    rel = (out_off[c0:c1 + 1] - base).astype(np.int64)
    ln = out_len[c0:c1].astype(np.int64)
    inside = np.zeros(len(a) + 1, np.int32)
    np.add.at(inside, rel[:-1], 1)
    np.add.at(inside, rel[:-1] + ln, -1)
    inside = np.cumsum(inside[:-1]) > 0                       # positions below out_len of their pair
    assert not a[~inside].any() and not b[~inside].any(), "stride tail is not zero"
    assert not ((a == ord("-")) & (b == ord("-")) & inside).any(), "a column pairs two gaps"
    ca = np.add.reduceat(((a != ord("-")) & inside).astype(np.int32), rel[:-1])
    cb = np.add.reduceat(((b != ord("-")) & inside).astype(np.int32), rel[:-1])
    assert (ca == M).all(), "aligned reference does not hold the reference's characters"
    assert (cb == N[c0:c1]).all(), "aligned query does not hold the query's characters"


def test_gpu_full_size_c2_properties(gpu_aligner, oracle_port):
    """BASELINE.json configs[1] at FULL size (1,000,000 reads vs HXB2 pol) through the one-shot C-ABI call, checked by
    size-independent properties on every pair - removing the gaps from the two aligned strings gives back M reference
    and N query characters, no column pairs two gaps, lengths within M+N, stride tails are zero - plus a batch-order
    check (a reversed batch gives the reversed outputs) and 500 pairs against the oracle."""
    from gotoh_b200 import packing, workloads
    n = 1000000
    ref, qb, qo = workloads.c2_reads_packed(n, seed=77)
    rb, ro = packing.pack([ref])
    ridx = np.zeros(n, np.int32)
    out_off = packing.out_offsets(ro, ridx, qo)
    out_a, out_b, _, out_len, score = gpu_aligner.align_packed(rb, ro, ridx, qb, qo, 10, 3, 1, 0, out_off=out_off)
    M, N = len(ref), np.diff(qo)
    assert (out_len >= np.maximum(M, N)).all() and (out_len <= M + N).all()
    for c0 in range(0, n, 50000):
        _check_chunk_properties(out_a, out_b, out_off, out_len, M, N, c0, min(n, c0 + 50000))
    for k in range(0, n, n // 500):
        q = qb[qo[k]:qo[k + 1]].tobytes().decode()
        s, ln = int(out_off[k]), int(out_len[k])
        exp = oracle_port.align_it(ref, q, 10, 3, 1)
        assert (out_a[s:s + ln].tobytes().decode(), out_b[s:s + ln].tobytes().decode(), int(score[k])) == exp
    # batch-order independence on the first 100,000 pairs: reversed order in, reversed results out
    m = 100000
    rq_off = np.zeros(m + 1, np.int64)
    np.cumsum(N[:m][::-1], out=rq_off[1:])
    rq = np.concatenate([qb[qo[k]:qo[k + 1]] for k in range(m - 1, -1, -1)])
    r_off = packing.out_offsets(ro, ridx[:m], rq_off)
    g2 = gpu_aligner.align_packed(rb, ro, ridx[:m], rq, rq_off, 10, 3, 1, 0, out_off=r_off)
    assert (g2[4][::-1] == score[:m]).all() and (g2[3][::-1] == out_len[:m]).all()
    for k in range(0, m, m // 1000):
        s1, s2, ln = int(out_off[k]), int(r_off[m - 1 - k]), int(out_len[k])
        assert (out_b[s1:s1 + ln] == g2[1][s2:s2 + ln]).all()
        assert (out_a[s1:s1 + ln] == g2[0][s2:s2 + ln]).all()


def test_gpu_rebase_rows_move_the_column0_seed(gpu_aligner, oracle_port, forced_path):
    """GPU twin of tests/test_emu_parity.py::test_emu_rebase_rows_move_the_column0_seed, plus a larger random batch with
    tiny gap-open penalties and a rebase period of 32 rows."""
    import test_emu_parity
    test_emu_parity.test_emu_rebase_rows_move_the_column0_seed(gpu_aligner, oracle_port, forced_path)
    rng = random.Random(199)
    alpha = "ARNDCQEGHILKMFPSTWYVBZX*-"
    refs = ["".join(rng.choice(alpha) for _ in range(rng.randint(60, 400))) for _ in range(5)]
    qs, ridx = [], []
    for _ in range(6000):
        qs.append("".join(rng.choice(alpha) for _ in range(rng.randint(1, 256))))
        ridx.append(rng.randrange(5))
    for gip, gep in ((0, 10), (1, 4), (0, 1)):
        got = gpu_aligner.align_batch(refs, qs, gip, gep, 1, 1, ref_idx=ridx)
        for k in range(len(qs)):
            assert got[k] == oracle_port.align_it_aa(refs[ridx[k]], qs[k], gip, gep, 1), (gip, gep, k)

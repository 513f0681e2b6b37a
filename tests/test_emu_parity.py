"""The exact kernel sources (micall-lite_b200/csrc), compiled for the TEST-ONLY CPU SIMT
emulator (tests/simt_emu), against the golden vectors and the oracle.  This is how kernel
logic is checked in the GPU-less container; the -m gpu tests repeat it on the real library."""
import os
import random

import numpy as np
import pytest

from conftest import load_golden, run_cases


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


def test_emu_appendix_b_and_fuzz(emu_aligner, forced_path):
    for name in ("appendix_b", "fuzz_small"):
        n, bad = run_cases(emu_aligner, load_golden(name)["cases"])
        assert n > 30 and not bad, "%s: %d/%d differ, first %r" % (name, len(bad), n, bad[0])


def test_emu_benchmark_shapes(emu_aligner, forced_path):
    """reads vs HXB2 pol (rebased int16 frames), multi-strip queries, aa windows."""
    n, bad = run_cases(emu_aligner, load_golden("shapes")["cases"], max_cells=1.2e6)
    assert n >= 90 and not bad, "%d/%d differ, first %r" % (len(bad), n, bad[0][0])


def test_emu_chunked_arena_and_mixed_batch(emu_aligner, oracle_port, monkeypatch):
    """One batch mixing short and multi-strip queries, several references, forced into many arena
    chunks; results must come back in caller order."""
    monkeypatch.setenv("GOTOH_B200_ARENA_MB", "1")
    rng = random.Random(5)
    refs = ["".join(rng.choice("ACGT") for _ in range(n)) for n in (700, 90, 333)]
    queries, ridx = [], []
    for k in range(60):
        r = rng.randrange(3)
        lo = rng.randrange(len(refs[r]) - 40)
        q = list(refs[r][lo:lo + rng.choice([30, 64, 65, 120, 257, 300])])
        for _ in range(3):
            q[rng.randrange(len(q))] = rng.choice("ACGTN")
        queries.append("".join(q))
        ridx.append(r)
    got = emu_aligner.align_batch(refs, queries, 10, 3, 1, 0, ref_idx=ridx)
    for k in range(60):
        assert got[k] == oracle_port.align_it(refs[ridx[k]], queries[k], 10, 3, 1), k


def test_emu_stop_codon_bonus_rule(emu_aligner, oracle_port, forced_path):
    """the "$$$" stop-codon bonus (gotoh.cpp:324-344) lives in the query profile."""
    rng = random.Random(31)
    refs, qs = [], []
    for _ in range(300):
        a = "".join(rng.choice("ACGT") for _ in range(rng.randint(3, 60)))
        for _ in range(rng.randint(1, 3)):
            p = rng.randrange(len(a) + 1)
            a = a[:p] + rng.choice(["$$$", "$$$$", "$$", "$$$$$$"]) + a[p:]
        b = a.replace("$$$", rng.choice(["TAG", "TAA", "TGA", "TGG"]))
        b = list(b.replace("$", rng.choice("ACGT")))
        for _ in range(rng.randint(0, 3)):
            b[rng.randrange(len(b))] = rng.choice("ACGTTAG")
        refs.append(a)
        qs.append("".join(b))
    for gip, gep, term in [(10, 3, 1), (0, 0, 0), (5, 1, 0)]:
        got = emu_aligner.align_batch(refs, qs, gip, gep, term, 0)
        for k in range(len(refs)):
            assert got[k] == oracle_port.align_it(refs[k], qs[k], gip, gep, term), (refs[k], qs[k], gip, gep, term)
    # the rule sits in align() itself, so it also fires under the amino-acid tables
    assert emu_aligner.align_it_aa("KF$$$R", "KFTAGR", 40, 10, 1) == oracle_port.align_it_aa("KF$$$R", "KFTAGR", 40, 10, 1)


def test_emu_mixed_int16x2_and_int32_launches_share_a_plan(emu_aligner, oracle_port):
    """Large penalties and a -20 score ('.') push the longer pairs of a batch past the int16 range proof while
    the shorter ones stay in int16x2: both kernels then run in ONE plan, the shifted int16 frame
    must not leak into the int32 launches, and the plan-wide shift must suit every admitted pair."""
    rng = random.Random(2026)
    alpha = "ACGTNRYKMSWBDHVacgtnXx*.-Uu"
    refs, qs = [], []
    for _ in range(260):
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
    for gip, gep, term in [(40, 10, 1), (25, 12, 0)]:
        got = emu_aligner.align_batch(refs, qs, gip, gep, term, 0)
        for k in range(len(refs)):
            assert got[k] == oracle_port.align_it(refs[k], qs[k], gip, gep, term), (refs[k], qs[k], gip, gep, term)
    # the batch really is mixed: a plan over the same pairs reports both paths
    from gotoh_b200 import packing
    rb, ro = packing.pack(refs)
    qb, qo = packing.pack(qs)
    plan = emu_aligner.plan(rb, ro, None, qb, qo, 40, 10, 1, 0)
    assert plan.stat(5) > 0 and plan.stat(6) > 0, (plan.stat(5), plan.stat(6))
    plan.close()


def _half_warp_batch(n, seed):
    """aa windows of ragged width (1..128: every K of the half-warp kernels) against PR/RT/INT in uneven proportions, so
    that warps get two couples, one couple + filler, and single pairs; plus nt reads cut to <= 128 columns."""
    from gotoh_b200 import workloads
    rng = random.Random(seed)
    refs, qs = workloads.c3_queries(n, seed=seed)
    pairs = []
    for k, q in enumerate(qs):
        r = refs[0] if k % 7 == 0 else refs[1] if k % 3 else refs[2]
        w = rng.choice([1, 2, 16, 17, 31, 32, 33, 47, 48, 49, 64, 65, 84, 95, 96, 97, 120, 128])
        pairs.append((r, (q * 2)[:w]))
    return pairs


def test_emu_half_warp_wavefronts(emu_aligner, oracle_port, monkeypatch):
    pairs = _half_warp_batch(151, 31)
    for arena_mb in (None, "1"):                      # "1": the arena is cut into many chunks between the warps' slabs
        if arena_mb:
            monkeypatch.setenv("GOTOH_B200_ARENA_MB", arena_mb)
        for gip, gep, term in [(40, 10, 1), (40, 10, 0), (3, 1, 1)]:
            got = emu_aligner.align_batch([a for a, _ in pairs], [b for _, b in pairs], gip, gep, term, 1)
            for (a, b), g in zip(pairs, got):
                assert g == oracle_port.align_it_aa(a, b, gip, gep, term), (a, b, gip, gep, term)
    monkeypatch.delenv("GOTOH_B200_ARENA_MB", raising=False)
    from gotoh_b200 import workloads
    ref, reads = workloads.c2_reads(40, seed=5)
    rng = random.Random(2)
    qs = [r[:rng.randint(1, 128)] for r in reads]
    got = emu_aligner.align_batch(ref[200:900], qs, 10, 3, 1, 0)
    for q, g in zip(qs, got):
        assert g == oracle_port.align_it(ref[200:900], q, 10, 3, 1)


def test_emu_many_tiny_pairs_take_the_radix_sorted_task_order(emu_aligner, oracle_port):
    """More than 2048 int16x2-eligible pairs in one plan: the host orders the task keys with its radix sort (smaller
    plans use std::sort), couples partners that share a reference and forms half-warp quads; tiny grids keep the
    emulator fast."""
    rng = random.Random(77)
    refs = ["".join(rng.choice("ACGT") for _ in range(rng.randint(4, 14))) for _ in range(9)]
    ridx = [rng.randrange(len(refs)) for _ in range(2600)]
    qs = ["".join(rng.choice("ACGTN") for _ in range(rng.randint(1, 12))) for _ in ridx]
    got = emu_aligner.align_batch(refs, qs, 10, 3, 1, 0, ref_idx=ridx)
    for k, (r, q) in enumerate(zip(ridx, qs)):
        assert got[k] == oracle_port.align_it(refs[r], q, 10, 3, 1), (refs[r], q)


def test_emu_many_reference_byte_classes(emu_aligner, oracle_port):
    """A reference using ~120 distinct bytes needs a 120-class query profile: the launcher drops to
    fewer warps per CTA instead of failing."""
    rng = random.Random(8)
    alpha = [chr(c) for c in range(1, 127) if chr(c) not in " \t\n\r$"]
    ref = "".join(alpha) + "".join(rng.choice(alpha) for _ in range(100))
    qs = ["".join(rng.choice(alpha) for _ in range(rng.randint(5, 90))) for _ in range(12)]
    got = emu_aligner.align_batch(ref, qs, 10, 3, 1, 0)
    for q, g in zip(qs, got):
        assert g == oracle_port.align_it(ref, q, 10, 3, 1)


@pytest.mark.parametrize("long_mode", ["flow", "cta", "warp"])
def test_emu_long_pairs_cta_wavefront_and_warp_strips(emu_aligner, oracle_port, monkeypatch, long_mode):
    """K2: queries wider than 256 columns.  'cta' = one CTA per pair, 4 warps pipelined over adjacent
    strips (shared-memory rings + the cross-round column); 'warp' = one warp walks the strips."""
    monkeypatch.setenv("GOTOH_B200_LONG", long_mode)
    rng = random.Random(12)
    refs, qs = [], []
    for M, N in [(600, 1100), (520, 2100), (900, 257), (1500, 1300), (40, 3000), (700, 1025), (512, 1024), (300, 2600)]:
        a = "".join(rng.choice("ACGT") for _ in range(M))
        b = list((a * (N // M + 2))[:N])
        for _ in range(N // 15):
            b[rng.randrange(N)] = rng.choice("ACGTN")
        refs.append(a)
        qs.append("".join(b))
    got = emu_aligner.align_batch(refs, qs, 15, 3, 1, 0)
    for k in range(len(refs)):
        assert got[k] == oracle_port.align_it(refs[k], qs[k], 15, 3, 1), (long_mode, len(refs[k]), len(qs[k]))


def test_emu_slab_pipeline_many_slabs(emu_aligner, oracle_port, monkeypatch):
    """The one-shot call cuts the batch into slabs that ping-pong between two workspaces; force
    one slab per handful of pairs and check order, contents and the zeroed stride tails."""
    from gotoh_b200 import packing, workloads
    monkeypatch.setenv("GOTOH_B200_SLAB_MB", "1")
    ref, qb, qo = workloads.c2_reads_packed(23, seed=21)
    rb, ro = packing.pack([ref])
    ridx = np.zeros(23, np.int32)
    got = emu_aligner.align_packed(rb, ro, ridx, qb, qo, 10, 3, 1, 0)
    exp = oracle_port.align_batch(0, rb, ro, ridx, qb, qo, 10, 3, 1)
    assert (got[3] == exp[3]).all() and (got[4] == exp[4]).all()
    assert (got[0] == exp[0]).all() and (got[1] == exp[1]).all()      # incl. zero tails: oracle buffers start zeroed


def test_emu_plan_interface_and_stats(emu_aligner, oracle_port):
    from gotoh_b200 import packing, workloads
    ref, qb, qo = workloads.c2_reads_packed(10, seed=3)
    rb, ro = packing.pack([ref])
    plan = emu_aligner.plan(rb, ro, np.zeros(10, np.int32), qb, qo, 10, 3, 1, 0)
    assert plan.cells == int((np.diff(qo) * 3039).sum())
    assert plan.stat(5) == 10 and plan.stat(6) == 0          # all on the int16x2 path
    plan.run()
    plan.run()                                               # re-runnable
    o_ref, o_qry, o_len, o_score = plan.fetch()
    a = packing.unpack(o_ref, plan.out_off, o_len)
    b = packing.unpack(o_qry, plan.out_off, o_len)
    for k, q in enumerate(workloads.unpacked(qb, qo)):
        assert (a[k], b[k], int(o_score[k])) == oracle_port.align_it(ref, q, 10, 3, 1)
    plan.close()


def test_emu_input_domain_errors(emu_aligner):
    from gotoh_b200 import GotohInputError
    with pytest.raises(GotohInputError):
        emu_aligner.align_it("  \n", "ACGT", 10, 3, 1)           # empty after trim (gotoh.cpp:555 UB)
    with pytest.raises(GotohInputError):
        emu_aligner.align_it("ACGT", "AC\x7fT", 10, 3, 1)        # byte 127 (gotoh.cpp:216-219 OOB)
    with pytest.raises(GotohInputError):
        emu_aligner.align_it("ACGT", "ACéT", 10, 3, 1)      # non-ASCII -> bytes >= 128
    with pytest.raises(GotohInputError):
        emu_aligner.align_it("A" * 40, "ACGT", 10, 5000, 1)      # -100000 sentinel domain (gotoh.cpp:284)
    with pytest.raises(ValueError):
        emu_aligner.align_it("ACGT", "AC\0T", 10, 3, 1)          # "s" rejects embedded NUL
    assert emu_aligner.align_batch(["ACGT"], [], 10, 3) == []


def test_emu_wrapper_semantics(emu_aligner, oracle_port):
    """trim on both inputs, degap + forced term=0 for align_it_aa_rb (gotoh.cpp:641-642,711-718)."""
    assert emu_aligner.align_it("  ACGT\n", "\tACT \r\n", 5, 1, 1) == ("ACGT", "AC-T", 9)
    assert emu_aligner.align_it_aa_rb("K-F-R", "KF--GR", 4, 2) == oracle_port.align_it_aa_rb("K-F-R", "KF--GR", 4, 2)
    assert emu_aligner.align_it_aa("WWWWKFR", "KFR", 40, 10, 0) == ("WWWWKFR", "----KFR", 104)


def test_emu_rebase_rows_move_the_column0_seed(emu_aligner, oracle_port, forced_path):
    """Regression (found by tools/fuzz_emu.py; present since round 1): in the int16x2 frame the column-0 seed of the Q
    recurrence is 4(u - base*g), so it has to move at every rebase row like all other stored values.  The stale seed only
    wins with a small gip, right after a rebase row, against a first-column mismatch: gip = 0, gep = 10, and a second pair
    whose width forces the plan's rebase period down to 32 rows."""
    refs = ['HNQRXSTW-CLBCRYNXYBXQFYXKRABSQKNEGPPMZTPNPFCEGZLHYNVVXMFK-**-EDPFCWXTZF*R*DXAXQXVPTRWXKYXGNBMQXMDGYBYEVFSCZRMKRZD*XXDKLECMTLCEFNRDKBRFYTYFRMCSADKZQXWVEZXFEA-AAYC', 'MVTZN*HKZS-MHEWVHSEAMSNZ-ZKDMXCLGNMPEEZVZM-WCDMEWMNCNZ*XFFQBFN-RB-WGGSHNIF-NT-XRRPXFXQYHXZQNBYCLQFXLFINBPRLQB-LQ*QCLAZWNVWQWWLABRNCSBLL*YI-TZVTGI-TARXTCTHXKA*GHH*SMEGV*YXWRG-HN-*SHYM*ZSDTXEKEDKHHSDEDT-NLXWKIEMPAMM*RDGZPCEGHSQVGMENLATHWFKVRFS-*-WKQKMCP']
    qs = ['FKHHSDEDT-NLXWKIEMPAMM*RDGZPCEGHSQVGMENLAKHWFKVRFS-*-WKQKMCP', 'VSTXWXKLKQSM--WFRRWKA*FH*STNGY*XBGHQEFMNTKM**ZKKMFBTAKMPGFXYCGYYBN*WQRN-QBB-FLNYMKQWM-KHKAZGDIDIWTPQBIYKLQCTDBBVGVGHMQNKKAW*TFVZANITTSQNQQTPAXIYZH-DXVHVGTXQSVQXDMZQ-SKNLGKGTXBWSBTYI-KSCGZBLK*QHBAZGVZLKTDSRTDNRZNAWHMSHYFZII*QYLMFWTGLBDZG']
    ridx = [1, 0]
    got = emu_aligner.align_batch(refs, qs, 0, 10, 1, 1, ref_idx=ridx)
    for k in range(len(qs)):
        assert got[k] == oracle_port.align_it_aa(refs[ridx[k]], qs[k], 0, 10, 1), k
    # the same rule at scale: random amino-acid pairs, tiny gap-open penalties, one wide pair per batch
    rng = random.Random(99)
    alpha = "ARNDCQEGHILKMFPSTWYVBZX*-"
    for gip, gep in ((0, 10), (1, 7), (0, 3)):
        refs = ["".join(rng.choice(alpha) for _ in range(rng.randint(100, 260))) for _ in range(3)]
        qs, ridx = [], []
        for _ in range(90):
            r = rng.randrange(3)
            qs.append("".join(rng.choice(alpha) for _ in range(rng.randint(1, 250))))
            ridx.append(r)
        got = emu_aligner.align_batch(refs, qs, gip, gep, 1, 1, ref_idx=ridx)
        for k in range(len(qs)):
            assert got[k] == oracle_port.align_it_aa(refs[ridx[k]], qs[k], gip, gep, 1), (gip, gep, k)


def test_emu_many_rebase_rows_on_the_c2_shape(emu_aligner, oracle_port):
    """251-nt reads against the 3039-nt pol seed with gep = 10 rebase every 256 rows: the column-0 seeds reach their floor
    after four rebase rows and must stay there (a 16-bit wrap-around there put 37 of the 123 golden shapes off on the GPU
    before the floor was computed in 32 bits)."""
    from gotoh_b200 import workloads
    ref, reads = workloads.c2_reads(8, seed=5)
    reads = ["T" + r[1:] for r in reads]
    for gip, gep, term in ((0, 10, 1), (10, 10, 0), (1, 3, 1)):
        got = emu_aligner.align_batch([ref], reads, gip, gep, term, 0, ref_idx=[0] * len(reads))
        for g, r in zip(got, reads):
            assert g == oracle_port.align_it(ref, r, gip, gep, term), (gip, gep, term)

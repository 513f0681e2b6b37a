"""the two callers of the live aligner, batched on the GPU, against golden vectors produced
by the reference's own Python (tests/golden/make_golden_callers.py: micall.core.remap.sam_to_conseqs and
micall.core.aln2counts.SequenceReport imported from /root/reference, incl. the reference tests' own inputs and the
distances they assert) and against the oracle.  emu_* tests run the kernel sources under the CPU SIMT emulator."""
import random

import pytest

from conftest import load_golden


def _check_levenshtein(lib, max_cells):
    from gotoh_b200 import remap_filter
    cases = [c for c in load_golden("callers")["levenshtein"] if len(c["a"]) * len(c["b"]) <= max_cells]
    got = remap_filter.distance_batch([(c["a"], c["b"]) for c in cases], library=lib)
    assert got == [c["d"] for c in cases]
    assert remap_filter.distance("kitten", "sitting", library=lib) == 3
    return len(cases)


def _check_remap_filter(lib, max_len):
    from collections import Counter
    from gotoh_b200 import remap_filter
    n = 0
    for c in load_golden("callers")["remap_filter"]:
        if max(len(s) for s in c["seeds"].values()) > max_len:
            continue
        report = {}
        kept = remap_filter.filter_conseqs(c["new_conseqs"], c["relevant"], c["seeds"], read_counts=Counter(c["read_counts"]),
                                           distance_report=report, library=lib)
        assert kept == c["expected_conseqs"], c["name"]
        assert report == c["expected_distances"], c["name"]
        n += 1
    return n


def _check_coordinate_map(lib, max_len):
    from gotoh_b200 import coordinate_map
    n = 0
    al = coordinate_map.default_aligner(library=lib)
    for c in load_golden("callers")["coordinate_map"]:
        if max(len(s) for s in c["seed_amino_seqs"]) > max_len:
            continue
        frames = {int(f): s for f, s in c["frame_consensus"].items()}
        reqs = [(co["coordinate_ref"], frames, c["consensus_length"], c["seed_amino_seqs"]) for co in c["coordinates"]]
        maps = coordinate_map.map_coordinate_refs(reqs, aligner=al)
        for co, m in zip(c["coordinates"], maps):
            assert m.reading_frame == co["reading_frame"], (c["name"], co["coordinate_name"])
            assert m.consensus == co["consensus"], (c["name"], co["coordinate_name"])
            assert [list(x) for x in m.conseq_indexes()] == co["conseq_indexes"], (c["name"], co["coordinate_name"])
            if co["inserts"] is not None:
                assert sorted(m.inserts()) == co["inserts"], (c["name"], co["coordinate_name"])
            n += 1
    return n


def test_extract_relevant_seed_reference_expectations():
    """micall/tests/remap_test.py:547-566 (testExtractRelevantSeeds), same expectations."""
    from gotoh_b200.remap_filter import extract_relevant_seed, relevant_conseq
    for aligned_conseq, aligned_seed, expected in [("ACTG", "ATTG", "ATTG"), ("-ACTG-", "CATTGT", "ATTG"), ("-AC-TG--", "CATATGT", "ATATG"),
                                                   ("-AC-TG-AT-", "CATATGTATC", "ATATGTAT"), ("--T--", "CATAT", "T"), ("TACG----", "----GGCC", "")]:
        assert extract_relevant_seed(aligned_conseq, aligned_seed) == expected
    counts = {1: {"A": 2}, 2: {"C": 1}, 3: {"G": 1, "T": 1}, 4: {}}
    assert relevant_conseq("ACGT", counts, 2) == "AG"                      # remap.py:236-240


def test_levenshtein_oracle_matches_golden():
    from oracle.oracle2 import levenshtein
    for c in load_golden("callers")["levenshtein"]:
        assert levenshtein(c["a"], c["b"]) == c["d"]


def test_emu_edit_distance(emu_aligner):
    assert _check_levenshtein(emu_aligner._libobj, 4e5) >= 100


def test_emu_remap_filter_reference_tests(emu_aligner):
    assert _check_remap_filter(emu_aligner._libobj, 1300) >= 8


def test_emu_coordinate_map(emu_aligner):
    assert _check_coordinate_map(emu_aligner._libobj, 1100) >= 10


def test_emu_edit_distance_many_byte_values(emu_aligner):
    """Bytes that occur on one side only share a class; more than 30 common bytes is outside the supported domain."""
    from gotoh_b200 import _ffi, remap_filter
    from oracle.oracle2 import levenshtein
    rng = random.Random(2)
    a = bytes(rng.randrange(1, 120) for _ in range(150))
    b = bytes(rng.choice(b"ACGT") if rng.random() < 0.5 else a[k % len(a)] for k in range(140))
    common = len(set(a) & set(b))
    if common <= 30:
        assert remap_filter.distance(a, b, library=emu_aligner._libobj) == levenshtein(a, b)
    # a pair that shares more than 28 distinct symbols is outside what the kernel's class table holds: a clear error
    wide = bytes(range(1, 80))
    with pytest.raises(ValueError):
        remap_filter.distance(wide, wide, library=emu_aligner._libobj)


@pytest.mark.gpu
def test_gpu_edit_distance(gpu_aligner):
    from gotoh_b200 import remap_filter, workloads
    from oracle.oracle2 import levenshtein
    lib = gpu_aligner._libobj
    assert _check_levenshtein(lib, 1e12) >= 129
    seeds = workloads.hcv_seeds()
    pairs = [(seeds[0], seeds[5]), (seeds[3][:5000], seeds[3][40:5100]), (seeds[9], seeds[9])]
    assert remap_filter.distance_batch(pairs, library=lib) == [levenshtein(a, b) for a, b in pairs]


@pytest.mark.gpu
def test_gpu_remap_filter_incl_hcv_genomes(gpu_aligner):
    assert _check_remap_filter(gpu_aligner._libobj, 10 ** 9) >= 9


@pytest.mark.gpu
def test_gpu_coordinate_map_incl_hcv(gpu_aligner):
    assert _check_coordinate_map(gpu_aligner._libobj, 10 ** 9) >= 20


def test_emu_distance_is_a_drop_in_for_arbitrary_text(emu_aligner):
    """Levenshtein.distance works on characters and on any alphabet; the device kernel takes at most 30 byte
    classes per batch.  Non-ASCII text and batches that mix alphabets are re-coded pair by pair."""
    import string
    from oracle.oracle2 import levenshtein
    from gotoh_b200 import remap_filter
    lib = emu_aligner._libobj
    assert remap_filter.distance("héllo", "hello", library=lib) == 1
    wide = string.ascii_letters[:40]               # 40 distinct symbols in the batch, 26 shared by this pair
    pairs = [(wide, wide[:26][::-1] + "0123456789"), ("ACGTN-ACGT", "ACGTTACGN"), ("abcdefghijklmnopqrstuvwxyz", "abcdefghijklmnopqrstuvwxy"),
             ("", "abc"), ("αβγδ", "αγδε"), (b"ACGT", b"AGT")]
    got = remap_filter.distance_batch(pairs, library=lib)
    def wagner_fischer(a, b):                       # characters, not bytes (small cases only)
        prev = list(range(len(b) + 1))
        for i, ca in enumerate(a, 1):
            cur = [i]
            for j, cb in enumerate(b, 1):
                cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
            prev = cur
        return prev[-1]
    exp = [wagner_fischer(a if isinstance(a, str) else a.decode(), b if isinstance(b, str) else b.decode()) for a, b in pairs]
    assert got == exp, (got, exp)
    assert exp[1] == levenshtein("ACGTN-ACGT", "ACGTTACGN")

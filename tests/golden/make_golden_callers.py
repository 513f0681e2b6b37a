#!/usr/bin/env python
"""Golden vectors for the two CALLERS of the live aligner, produced by THE REFERENCE'S OWN
PYTHON, imported from /root/reference and run unmodified:

  remap_filter     micall.core.remap.sam_to_conseqs(..., is_filtered=True) (remap.py:228-263) on the inputs of the
                   reference's own tests (micall/tests/remap_test.py:415-545, incl. the distances they assert) and on
                   synthetic HCV consensuses; `Levenshtein` (third-party, not installed here) is provided by
                   oracle/levenshtein_oracle.c - the reference tests' asserted distances pin that restatement.
  coordinate_map   micall.core.aln2counts.SequenceReport.read -> _map_to_coordinate_ref (aln2counts.py:191-304) on
                   the tiny projects of micall/tests/aln2counts_test.py and on real seeds of projects.json.
  levenshtein      distances of the oracle on seeded strings (cross-checked against a pure-Python DP here).

The reference's `_gotoh2.c` is the compiled oracle/_ref/_gotoh2*.so.  Run only where /root/reference exists:
    make -C oracle && python tests/golden/make_golden_callers.py        -> tests/golden/callers.json
"""
import glob
import importlib.util
import json
import os
import random
import sys
import types
import warnings
from collections import Counter
from io import StringIO

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.oracle2 import levenshtein  # noqa: E402

REF = "/root/reference"


def import_reference():
    sys.path.insert(0, REF)
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "_gotoh2*.so"))[0]
    spec = importlib.util.spec_from_file_location("micall.alignment._gotoh2", so)
    ext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ext)
    import micall.alignment
    sys.modules["micall.alignment._gotoh2"] = ext
    micall.alignment._gotoh2 = ext
    lev = types.ModuleType("Levenshtein")
    lev.distance = levenshtein
    sys.modules["Levenshtein"] = lev
    from micall.core import aln2counts, project_config, remap
    from micall.utils import translation
    return remap, aln2counts, project_config, translation


def mutate(rng, seq, sub=0.03, indels=4, alphabet="ACGT"):
    s = list(seq)
    for k in range(len(s)):
        if rng.random() < sub:
            s[k] = rng.choice(alphabet)
    for _ in range(indels):
        p = rng.randrange(len(s))
        if rng.random() < 0.5:
            del s[p:p + rng.randint(1, 6)]
        else:
            s[p:p] = [rng.choice(alphabet) for _ in range(rng.randint(1, 6))]
    return "".join(s)


# ---------------------------------------------------------------------------------------------------------
def remap_cases(remap):
    J = "J"
    tests = [   # (name, sam body, seeds, filter_coverage) - inputs of remap_test.py:415-545
        ("testSeedsConverged remap_test.py:415",
         [("test1", "test", 1, "10M", "ATGAGGAGTA"), ("other1", "other", 1, "10M", "ATGACCAGTA"), ("wayoff1", "wayoff", 1, "10M", "ATGAGGGTAC")],
         {"test": "ATGAAGTA", "other": "AAGCCGAA", "wayoff": "TCATGTAC"}, 1),
        ("testSeedsConvergedWithDifferentAlignment remap_test.py:445",
         [("test1", "test", 1, "10M", "ATGAGGAGTA"), ("other1", "other", 11, "10M", "ATGACCAGTA")],
         {"test": "ATGAAGTA", "other": "TCTCTCTCTCAAGCCGAA"}, 1),
        ("testSeedsConvergedWithDifferentAlignmentAndGap remap_test.py:463",
         [("test1", "test", 1, "10M", "ATGAGGAGTA"), ("other1", "other", 11, "5M", "ATGAC"), ("other2", "other", 26, "5M", "CAGTA")],
         {"test": "ATGAAGTA", "other": "TCTCTCTCTCAAGCTATATATATACGAA"}, 1),
        ("testSeedsConvergedWithConfusingGap remap_test.py:482",
         [("test1", "test", 1, "8M", "ATGTCGTA"), ("other1", "other", 14, "9M", "AAGCTATAT")],
         {"test": "ATGAAGTA", "other": "ATGTCTCTCTCTCAAGCTATATATATACGAAGTA"}, 1),
        ("testSeedsConvergedPlusOtherLowCoverage remap_test.py:501",
         [("test1", "test", 1, "10M", "ATGAGGAGTA"), ("test2", "test", 1, "10M", "ATGAGGAGTA"), ("other1", "other", 1, "10M", "ATGACCAGTA"),
          ("other2", "other", 1, "10M", "ATGACCAGTA"), ("other3", "other", 11, "6M", "GTGTGT")],
         {"test": "ATGAAGTACTCTCT", "other": "AAGCCGAAGTGTGT"}, 2),
        ("testAllSeedsLowCoverage remap_test.py:524",
         [("test1", "test", 1, "10M", "ATGAGGAGTA"), ("test2", "test", 11, "6M", "CTCTCT"), ("other1", "other", 1, "10M", "ATGACCAGTA")],
         {"test": "ATGAAGTACTCTCT", "other": "AAGCCGAAGTGTGT"}, 2),
    ]
    # synthetic: consensuses that drifted from their seeds (short, medium, HCV genome length)
    rng = random.Random(20260105)
    nt = json.load(open(os.path.join(ROOT, "micall-lite_b200", "gotoh_b200", "data", "references.json")))["nucleotide"]
    hcv = [k for k in sorted(nt) if k.startswith("HCV-")]
    for size, nseeds, label in ((300, 4, "synthetic 300 nt x 4 seeds"), (1200, 3, "synthetic 1.2 kb x 3 seeds"),
                                (None, 3, "HCV genomes x 3 seeds (remap's real shape)")):
        names = rng.sample(hcv, nseeds)
        start = rng.randrange(0, 5000)
        seeds = {n: (nt[n] if size is None else nt[n][start:start + size]) for n in names}
        reads = []
        for k, n in enumerate(names):
            src = seeds[names[(k + 1) % nseeds]] if k == 0 else seeds[n]        # the first consensus drifted to another seed
            seq = mutate(rng, src, 0.02, 3)[:len(seeds[n])]
            reads.append((n + "_r", n, 1, "%dM" % len(seq), seq))
        tests.append((label, reads, seeds, 1))

    out = []
    for name, reads, seeds, cov in tests:
        sam = "@SQ\t" + "\t".join("SN:" + s for s in seeds) + "\n"
        for qname, rname, pos, cigar, seq in reads:
            sam += "\t".join([qname, "99", rname, str(pos), "44", cigar, "=", "1", str(len(seq)), seq, J * len(seq)]) + "\n"
        captured = {}
        real_c2c, real_align = remap.counts_to_conseqs, remap.aligner.align
        calls = []

        def c2c(refmap):
            captured["new_conseqs"] = real_c2c(refmap)
            return captured["new_conseqs"]

        def align(a, b):
            r = real_align(a, b)
            calls.append((a, b))
            return r

        remap.counts_to_conseqs = c2c
        remap.aligner.align = align
        try:
            dist = {}
            kept = remap.sam_to_conseqs(StringIO(sam), seeds=seeds, is_filtered=True, filter_coverage=cov, distance_report=dist)
        finally:
            remap.counts_to_conseqs = real_c2c
            remap.aligner.align = real_align
        new_conseqs = captured["new_conseqs"]
        # the relevant consensus of each name is the query of its alignments (remap.py:248); names without calls had none
        relevant = {n: "" for n in new_conseqs}
        order = sorted(new_conseqs)
        per_name = len(order)
        idx = 0
        names_with_calls = [n for n in order if n in dist]
        for n in names_with_calls:
            relevant[n] = calls[idx][1]
            idx += per_name
        read_counts = Counter(r[1] for r in reads)
        out.append({"name": name, "new_conseqs": new_conseqs, "relevant": relevant, "seeds": seeds,
                    "read_counts": dict(read_counts), "expected_conseqs": kept, "expected_distances": dist})
        print("remap_filter", name, {k: len(v) for k, v in new_conseqs.items()}, dist)
    return out


# ---------------------------------------------------------------------------------------------------------
def coordinate_cases(aln2counts, project_config, translation):
    rng = random.Random(20260106)
    out = []

    def run(label, projects, seed, rows):
        report = aln2counts.SequenceReport(aln2counts.InsertionWriter(StringIO()), projects, [0.1])
        report.read(rows)
        seed_nuc = projects.getReference(seed)
        if type(seed_nuc) == bytes:
            seed_nuc = seed_nuc.decode("utf-8")
        frames = {str(f): "".join(a.get_consensus() for a in aminos) for f, aminos in report.seed_aminos.items()}
        clen = len([a for a in report.seed_aminos[0] if a.counts])
        seeds_aa = [translation.translate(seed_nuc, offset=f, ambig_char="-") for f in range(3)]
        coords = []
        for cname, cref in report.coordinate_refs.items():
            if type(cref) == bytes:
                cref = cref.decode("utf-8")
            coords.append({"coordinate_name": cname, "coordinate_ref": cref,
                           "reading_frame": report.reading_frames.get(cname), "consensus": report.consensus[cname],
                           "conseq_indexes": [[ra.position, ra.seed_amino.consensus_index] for ra in report.reports[cname]],
                           "inserts": sorted(report.inserts[cname]) if cname in report.inserts else None})
        out.append({"name": label, "seed": seed, "frame_consensus": frames, "consensus_length": clen,
                    "seed_amino_seqs": seeds_aa, "coordinates": coords})
        print("coordinate_map", label, seed, clen, [(c["coordinate_name"], c["reading_frame"], len(c["conseq_indexes"])) for c in coords])

    def rows_for(seed, seqs):
        return [dict(refname=seed, qcut="15", rank=str(k), count=str(c), offset=str(o), seq=s) for k, (c, o, s) in enumerate(seqs)]

    # tiny projects of micall/tests/aln2counts_test.py:45-150 (R1..R3) with reads in the style of its tests
    class NamedIO(StringIO):
        name = "tiny_projects.json"        # ProjectConfig.load records json_file.name (project_config.py:40)

    tiny = project_config.ProjectConfig()
    tiny.load(NamedIO(json.dumps({
        "projects": {r: {"max_variants": 10, "regions": [{"coordinate_region": r, "seed_region_names": [r + "-seed"]}]} for r in ("R1", "R2", "R3")},
        "regions": {"R1-seed": {"is_nucleotide": True, "reference": ["AAATTTAGG"]}, "R1": {"is_nucleotide": False, "reference": ["KFR"]},
                    "R2-seed": {"is_nucleotide": True, "reference": ["AAATTTGGCCCGAGA"]}, "R2": {"is_nucleotide": False, "reference": ["KFGPR"]},
                    "R3-seed": {"is_nucleotide": True, "reference": ["AAATTTCAGACCCCACGAGAGCAT"]}, "R3": {"is_nucleotide": False, "reference": ["KFQTPREH"]}}})))
    run("aln2counts_test style: R1 exact", tiny, "R1-seed", rows_for("R1-seed", [(9, 0, "AAATTTAGG")]))
    run("aln2counts_test style: R1 offset frame", tiny, "R1-seed", rows_for("R1-seed", [(9, 1, "AATTTAGG")]))
    run("aln2counts_test style: R2 deletion", tiny, "R2-seed", rows_for("R2-seed", [(9, 0, "AAATTTCCGAGA")]))
    run("aln2counts_test style: R2 insertion", tiny, "R2-seed", rows_for("R2-seed", [(9, 0, "AAATTTGGCAACCCGAGA")]))
    run("aln2counts_test style: R3 partial + mixture", tiny, "R3-seed",
        rows_for("R3-seed", [(5, 0, "AAATTTCAGACCCCACGAGAGCAT"), (4, 3, "TTTCAGACCCCACGA"), (1, 0, "AAATTTCAGACTCCACGAGAGCAT")]))
    # real seeds of projects.json with mutated full-length and partial reads
    # ProjectConfig.loadDefault opens with mode 'rU' (project_config.py:15), which Python >= 3.11 rejects: load it directly
    real = project_config.ProjectConfig()
    with open(os.path.join(REF, "micall", "projects.json")) as f:
        real.load(f)
    for seed, label in (("HIV1B-vpr-seed", "HIV1B vpr (small)"), ("HIV1B-nef-seed", "HIV1B nef"), ("HIV1B-pol-seed", "HIV1B pol: PR/RT/INT"),
                        ("HCV-1a", "HCV-1a genome: 10 coordinate references")):
        try:
            ref = real.getReference(seed)
        except KeyError:
            print("no seed", seed)
            continue
        if type(ref) == bytes:
            ref = ref.decode("utf-8")
        full = mutate(rng, ref, 0.02, 3)
        part_start = (len(ref) // 4) // 3 * 3
        part = mutate(rng, ref[part_start:part_start + len(ref) // 2], 0.02, 1)
        run(label, real, seed, rows_for(seed, [(7, 0, full), (3, part_start, part)]))
    return out


def levenshtein_cases():
    rng = random.Random(20260107)
    out = []

    def slow(a, b):
        prev = list(range(len(b) + 1))
        for i, ca in enumerate(a, 1):
            cur = [i]
            for j, cb in enumerate(b, 1):
                cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
            prev = cur
        return prev[-1]

    fixed = [("kitten", "sitting"), ("", "abc"), ("abc", ""), ("", ""), ("ATGAAGTA", "ATGAGGAGTA"), ("A", "A"), ("A", "C"),
             ("flaw", "lawn"), ("ACGT" * 70, "ACGT" * 70)]
    for a, b in fixed:
        out.append({"a": a, "b": b, "d": levenshtein(a, b)})
    for _ in range(120):
        la, lb = rng.randint(1, 600), rng.randint(1, 600)
        a = "".join(rng.choice("ACGTN-acgtRYK") for _ in range(la))
        b = list(a[:lb]) if rng.random() < 0.6 else [rng.choice("ACGT") for _ in range(lb)]
        for _ in range(len(b) // 10):
            b[rng.randrange(len(b))] = rng.choice("ACGTXYZ*")
        b = "".join(b)
        d = levenshtein(a, b)
        if la * lb < 40000:
            assert d == slow(a, b)
        out.append({"a": a, "b": b, "d": d})
    return out


def main():
    remap, aln2counts, project_config, translation = import_reference()
    doc = {"_how": "tests/golden/make_golden_callers.py (reference Python imported from /root/reference, _gotoh2.c compiled unmodified)",
           "remap_filter": remap_cases(remap), "coordinate_map": coordinate_cases(aln2counts, project_config, translation),
           "levenshtein": levenshtein_cases()}
    path = os.path.join(HERE, "callers.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=0)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

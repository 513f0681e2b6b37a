#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running THE REFERENCE ITSELF.

Run only in the build container, where /root/reference exists:

    make -C oracle            # compiles /root/reference/micall/alignment/gotoh.cpp unmodified
    python tests/golden/make_golden.py

Every expected output below comes from oracle/_ref/libgotoh_ref.so, i.e. from the
reference's own align()/init_pairscore*() (gotoh.cpp:26-527) driven exactly like its
Python wrappers (gotoh.cpp:624-727).  The reference's test-suite has no vectors for
this path, so these fixtures are what pins parity on the
GPU box, where /root/reference does not exist.
"""
import gzip
import hashlib
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle  # noqa: E402

REFS = json.load(open(os.path.join(ROOT, "micall-lite_b200", "gotoh_b200", "data", "references.json")))


def sha(s):
    return hashlib.sha256(s.encode("latin-1")).hexdigest()


def revcomp(s):
    return s[::-1].translate(str.maketrans("ACGTacgt", "TGCAtgca"))


# SURVEY.md Appendix B inputs (outputs are re-derived from the reference here).
APPENDIX_B = [
    (0, "ACGT", "ACT", 5, 1, 1), (0, "TACGTA", "ACGT", 5, 1, 1), (0, "TACGTA", "ACGT", 5, 1, 0),
    (0, "ACGT", "TTACGTTT", 5, 1, 1), (0, "ACGT", "TTACGTTT", 5, 1, 0),
    (0, "ACGTACGT", "ACGTACGT", 10, 3, 1), (0, "AAAA", "TTTT", 10, 3, 1), (0, "AAAA", "TTTT", 10, 3, 0),
    (0, "ACGTTTACGT", "ACGTACGT", 10, 3, 1), (0, "ACGTACGT", "ACGTTTACGT", 10, 3, 1),
    (0, "AACCGGTT", "CCGG", 10, 3, 1), (0, "AACCGGTT", "CCGG", 10, 3, 0), (0, "CCGG", "AACCGGTT", 10, 3, 0),
    (0, "ACGT", "acgt", 10, 3, 1), (0, "ACNT", "ACGT", 10, 3, 1), (0, "ARGT", "AAGT", 10, 3, 1),
    (0, "AYGT", "AAGT", 10, 3, 1), (0, "GATTACA", "GCATGCT", 1, 1, 1), (0, "GATTACA", "GCATGCT", 0, 1, 1),
    (0, "AAAAAAAAAA", "AAAAA", 10, 3, 1), (0, "AAAAA", "AAAAAAAAAA", 10, 3, 1), (0, "ATATATAT", "ATAT", 2, 1, 1),
    (0, "A", "A", 10, 3, 1), (0, "A", "C", 10, 3, 1), (0, "A", "C", 10, 3, 0),
    (0, "ACG$$$ACG", "ACGTAGACG", 10, 3, 1), (0, "ACG$$$ACG", "ACGCCCACG", 10, 3, 1), (0, "ACG$$$", "ACGTAA", 10, 3, 1),
    (0, "  ACGT\n", "\tACT \r\n", 5, 1, 1),
    (1, "ERM", "ERM", 40, 10, 1), (1, "KFR", "KFGR", 40, 10, 1), (1, "KFGPR", "KFPR", 40, 10, 1),
    (1, "KFGPR", "KFPR", 40, 10, 0), (1, "KF*R", "KFJR", 40, 10, 1), (1, "KFXR", "KFXR", 40, 10, 1),
    (1, "kfr", "KFR", 40, 10, 1), (1, "WWWWKFR", "KFR", 40, 10, 1), (1, "WWWWKFR", "KFR", 40, 10, 0),
    (1, "PQITLWQRPLVTIKIGGQLKEALLDTGADDTVLEEMSLPGRWKPKMIGGIGGFIKVRQYDQILIEICGHKAIGTVLVGPTPVNIIGRNLLTQIGCTLNF",
     "PQITLWQRPLVTIKIGGQLKEALLDTGADDTVLEEMNLPGKWKPKMIGGIGGFIKVRQYDQIPIEICGHKAIGTVLVGPTPVNIIGRNLLTQIGCTLNF", 40, 10, 1),
    (2, "KFR", "KFGR", 40, 10, 0), (2, "K-F-R", "KF--GR", 4, 2, 0), (2, "KXR", "K-ZR", 4, 2, 0),
]

NT_ALPHA = "ACGTNRYKMSWBDHVacgtnXx*.-Uu"
AA_ALPHA = "ARNDCQEGHILKMFPSTWYVBZ?*XJ_-akl"


def mutate(rng, s, alpha, sub=0.03, nindel=2, maxindel=3):
    out = []
    for c in s:
        out.append(rng.choice(alpha) if rng.random() < sub else c)
    for _ in range(rng.randint(0, nindel)):
        p = rng.randrange(len(out) + 1)
        k = rng.randint(1, maxindel)
        if rng.random() < 0.5:
            del out[p:p + k]
        else:
            out[p:p] = [rng.choice(alpha) for _ in range(k)]
    return "".join(out) or alpha[0]


def fuzz_cases(seed, n):
    rng = random.Random(seed)
    cases = []
    while len(cases) < n:
        mode = rng.choice([0, 0, 0, 1, 1, 2])
        if mode == 0:
            alpha = "ACGT" if rng.random() < 0.5 else NT_ALPHA
        else:
            alpha = AA_ALPHA
        M = rng.randint(1, 70)
        a = "".join(rng.choice(alpha) for _ in range(M))
        if rng.random() < 0.6:
            lo = rng.randrange(len(a))
            b = mutate(rng, a[lo:lo + rng.randint(1, 70)], alpha)
        else:
            b = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 70)))
        if mode == 0 and rng.random() < 0.08:
            p = rng.randrange(len(a)); a = a[:p] + "$$$" + a[p:]
            p = rng.randrange(len(b)); b = b[:p] + rng.choice(["TAG", "TAA", "TGA"]) + b[p:]
        if mode == 2 and (not a.replace("-", "") or not b.replace("-", "")):
            continue
        gip = rng.choice([0, 1, 2, 5, 10, 15, 40])
        gep = rng.choice([0, 1, 3, 10])
        term = 0 if mode == 2 else rng.choice([0, 1])
        cases.append((mode, a, b, gip, gep, term))
    return cases


def shape_cases():
    """Benchmark-shaped pairs: reads vs HXB2 pol, aa vs PR/RT/INT, multi-strip, HCV."""
    rng = random.Random(20260101)
    pol = REFS["nucleotide"]["HIV1B-pol-seed"]
    cases = []
    # C1 substitute: real reads from the reference's example FASTQs (R2 reverse-complemented)
    ex = "/root/reference/examples/HIV1C-pol_S1_L001_R%d_001.fastq.gz"
    for rno in (1, 2):
        with gzip.open(ex % rno, "rt") as f:
            reads = [l.strip() for i, l in enumerate(f) if i % 4 == 1]
        for k in rng.sample(range(len(reads)), 12):
            r = reads[k]
            # the example run mixes orientations; keep both so unrelated (random-like) pairs are covered too
            cases.append((0, {"ref": "HIV1B-pol-seed"}, r, 10, 3, 1))
            cases.append((0, {"ref": "HIV1B-pol-seed"}, revcomp(r), 10, 10, 0))
    # C2: synthetic 251-nt reads
    for k in range(24):
        lo = rng.randrange(len(pol) - 251)
        q = mutate(rng, pol[lo:lo + 251], "ACGT", sub=0.02, nindel=1)
        if rng.random() < 0.3:
            p = rng.randrange(len(q)); q = q[:p] + "N" + q[p + 1:]
        cases.append((0, {"ref": "HIV1B-pol-seed"}, q, 10, 3, 1) if k % 2 else (0, {"ref": "HIV1B-pol-seed"}, q, 10, 10, 0))
    # edge widths around the 256-column strip boundary and multi-strip queries
    for n in (255, 256, 257, 300, 511, 512, 513, 700, 1031):
        lo = rng.randrange(len(pol) - n)
        cases.append((0, {"ref": "HIV1B-pol-seed"}, mutate(rng, pol[lo:lo + n], "ACGT", sub=0.03, nindel=3, maxindel=9), 10, 3, 1))
        cases.append((0, mutate(rng, pol[lo:lo + n], "ACGTN", sub=0.05, nindel=3), pol[lo:lo + 90], 15, 3, n % 2))
    # C3: amino acid windows vs PR / RT / INT
    for k in range(30):
        name = ("PR", "RT", "INT")[k % 3]
        ref = REFS["amino"][name]
        lo = rng.randrange(max(1, len(ref) - 84))
        q = mutate(rng, ref[lo:lo + 84], "ARNDCQEGHILKMFPSTWYV*?", sub=0.03, nindel=1, maxindel=1)
        cases.append((1, {"aa": name}, q, 40, 10, k % 2))
    # C4: HCV genome vs genome / mutated copy (hash-only outputs)
    names = sorted(k for k in REFS["nucleotide"] if k.startswith("HCV-"))
    a, b = names[0], names[7]
    cases.append((0, {"ref": a}, {"ref": b}, 15, 3, 1))
    src = REFS["nucleotide"][names[20]]
    cases.append((0, {"ref": names[20]}, mutate(rng, src, "ACGT", sub=0.05, nindel=10, maxindel=30), 15, 3, 1))
    cases.append((0, {"ref": names[3]}, src[4000:4250], 10, 10, 0))  # reference_distances.py:31-41 shape
    return cases


def resolve(x):
    if isinstance(x, dict):
        return REFS["nucleotide"][x["ref"]] if "ref" in x else REFS["amino"][x["aa"]]
    return x


def run(R, cases, hash_over=2000):
    out = []
    for mode, a, b, gip, gep, term in cases:
        ra, rb, sc = R.align(mode, resolve(a), resolve(b), gip, gep, term)
        rec = {"mode": mode, "a": a, "b": b, "gip": gip, "gep": gep, "term": term, "score": sc, "len": len(ra)}
        if len(ra) > hash_over:
            rec["sha_a"], rec["sha_b"] = sha(ra), sha(rb)
        else:
            rec["out_a"], rec["out_b"] = ra, rb
        out.append(rec)
    return out


def main():
    R = Oracle("reference")
    doc = {"_generated_by": "tests/golden/make_golden.py from oracle/_ref/libgotoh_ref.so "
                            "(= /root/reference/micall/alignment/gotoh.cpp, unmodified)"}
    json.dump(dict(doc, cases=run(R, APPENDIX_B)), open(os.path.join(HERE, "appendix_b.json"), "w"), indent=0)
    json.dump(dict(doc, cases=run(R, fuzz_cases(7, 1500))), open(os.path.join(HERE, "fuzz_small.json"), "w"), indent=0)
    json.dump(dict(doc, cases=run(R, shape_cases())), open(os.path.join(HERE, "shapes.json"), "w"), indent=0)
    import numpy as np
    tabs = {str(m): R.table(m).astype(np.int8).tolist() for m in range(3)}
    assert all((R.table(m) == np.array(tabs[str(m)])).all() for m in range(3))
    json.dump(dict(doc, tables=tabs), open(os.path.join(HERE, "pairscore_tables.json"), "w"))
    for f in ("appendix_b", "fuzz_small", "shapes", "pairscore_tables"):
        print(f, os.path.getsize(os.path.join(HERE, f + ".json")))


if __name__ == "__main__":
    main()

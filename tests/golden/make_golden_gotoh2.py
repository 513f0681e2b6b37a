#!/usr/bin/env python
"""Golden vectors for the live aligner gotoh2.Aligner.align, produced by THE
REFERENCE ITSELF: oracle/_ref/_gotoh2*.so is /root/reference/micall/alignment/src/_gotoh2.c compiled
unmodified, driven exactly like gotoh2.py:74-96.  Run only where /root/reference exists.

    make -C oracle && python tests/golden/make_golden_gotoh2.py

Contents of tests/golden/gotoh2.json:
  kats   the reference's own unit tests (micall/alignment/tests/test.py:174-298) with the answers the
         test file asserts, re-derived here from the compiled reference (inputs incl. the NL4-3 /
         HXB2 sequences are read out of that test file's setUp)
  fuzz   seeded random cases over all four models, global and local, penalties incl. 0
Cases whose alignment has length l1+l2 (no column pairs two characters) are skipped: there the
reference writes its NUL terminator one past `char aligned1[l1+l2]` (_gotoh2.c:481,424) and returns
corrupted strings - undefined behaviour, not a behaviour to match.
"""
import hashlib
import json
import os
import random
import re
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.oracle2 import Oracle2  # noqa: E402

TEST_PY = "/root/reference/micall/alignment/tests/test.py"


def reference_test_sequences():
    src = open(TEST_PY).read()
    body = src[src.index("def setUp(self):"):src.index("class TestAlignerSimpleGlobal")]
    body = "\n".join(l[8:] for l in body.splitlines()[1:] if "Aligner()" not in l and "gap_open_penalty" not in l)
    ns = {"self": types.SimpleNamespace()}
    exec(body, ns)
    return vars(ns["self"])


def sha(s):
    return hashlib.sha256(s.encode()).hexdigest()


def record(R, name, a, b, gop, gep, is_global, model, hash_over=1500):
    try:
        r = R.align(a, b, gop, gep, is_global, model)
    except RuntimeError:
        return {"name": name, "a": a, "b": b, "gop": gop, "gep": gep, "is_global": is_global, "model": model, "error": "traceback"}
    rec = {"name": name, "a": a, "b": b, "gop": gop, "gep": gep, "is_global": is_global, "model": model,
           "score": r[2], "len": len(r[0])}
    if len(r[0]) > hash_over:
        rec["sha_a"], rec["sha_b"] = sha(r[0]), sha(r[1])
    else:
        rec["out_a"], rec["out_b"] = r[0], r[1]
    return rec


def main():
    R = Oracle2("reference")
    P = Oracle2("port")
    seqs = reference_test_sequences()
    kats = [
        record(R, "TestAlignerSimpleGlobal test.py:174", "ACGT", "ACT", 5, 1, True, "HYPHY_NUC"),
        record(R, "TestAlignerLongerGlobal test.py:185", "ACGTACGTACGTACGT", "ACGTACGTACTACGT", 5, 1, True, "HYPHY_NUC"),
        record(R, "TestAlignerSimpleLocal test.py:193", "TACGTA", "ACGT", 5, 1, False, "HYPHY_NUC"),
        record(R, "TestHIV test.py:205", seqs["nl43"], seqs["hxb2_rt"], 5, 1, False, "HYPHY_NUC"),
        record(R, "TestFlouri.test_NWalign_example test.py:218", "GGTGTGA", "TCGCGT", 10, 1, True, "NWALIGN"),
        record(R, "TestFlouri.test_Biopp_example1 test.py:228", "AAAGGG", "TTAAAAGGGGTT", 5, 1, True, "Biopp"),
        record(R, "TestIssues.test_issue6 test.py:248", seqs["hxb2_integrase"][:100], seqs["u54771"][:100], 2, 1, True, "HYPHY_NUC"),
        record(R, "TestIssues.test_issue16a test.py:255", "AT", "ATTTTTT", 5, 1, True, "HYPHY_NUC"),
        record(R, "TestIssues.test_issue16b", "AT", "ATTTTT", 5, 1, False, "HYPHY_NUC"),
        record(R, "TestIssues.test_issue16c", "A", "ATTTTT", 5, 1, False, "HYPHY_NUC"),
        record(R, "TestIssues.test_issue14 test.py:276", "GCA", "CA", 10, 1, True, "HYPHY_NUC"),
        record(R, "TestIssues.test_issue15 test.py:289 (skipped upstream)", "ERM", "ERM", 40, 10, False, "EmpHIV25"),
        record(R, "clean_sequence gotoh2.py:70-72", "acgtnRY-x", "ACGTACGT", 5, 1, True, "HYPHY_NUC"),
        record(R, "aln2counts.py:34-37 settings", "PQITLWQRPLVTIKIGGQLKEALLDTGADDTVLEEMSLPGRWKPKMIGGIGGFIKVRQYDQILIEICGHKAIGTVLVGPTPVNIIGRNLLTQIGCTLNF",
               "QRPLVTIKIGGQLKEALLDTGADDTVLEEMNLPGKWKPKMIGGIGGFIKVRQYDQIPIEICGHK", 40, 10, False, "EmpHIV25"),
        record(R, "remap.py:33 settings", seqs["hxb2_integrase"], seqs["hxb2_integrase"][30:500].replace("GGAA", "GGA"), 15, 3, True, "HYPHY_NUC"),
    ]
    # the answers the reference's test file asserts
    assert (kats[0]["out_a"], kats[0]["out_b"], kats[0]["score"]) == ("ACGT", "AC-T", 9)
    assert kats[1]["out_b"] == "ACGTACGTAC-TACGT"
    assert (kats[2]["out_a"], kats[2]["out_b"], kats[2]["score"]) == ("TACGTA", "-ACGT-", 20)
    assert kats[4]["score"] == -3 and kats[5]["score"] == -15
    assert (kats[7]["out_a"], kats[7]["out_b"], kats[7]["score"]) == ("AT-----", "ATTTTTT", 0)
    assert (kats[8]["out_a"], kats[8]["out_b"], kats[8]["score"]) == ("AT----", "ATTTTT", 10)
    assert (kats[9]["out_a"], kats[9]["out_b"], kats[9]["score"]) == ("A-----", "ATTTTT", 5)
    assert (kats[10]["out_a"], kats[10]["out_b"], kats[10]["score"]) == ("GCA", "-CA", -1)

    rng = random.Random(20260202)
    fuzz = []
    while len(fuzz) < 1500:
        model = rng.choice(["HYPHY_NUC", "HYPHY_NUC", "NWALIGN", "Biopp", "EmpHIV25"])
        alpha = "ACGTNacgt-" if model != "EmpHIV25" else "ARNDCQEGHILKMFPSTWYVBZX*?ak"
        a = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 70)))
        if rng.random() < 0.65:
            b = list(a[rng.randrange(len(a)):][:rng.randint(1, 70)] or a)
            for _ in range(rng.randint(0, 4)):
                if rng.random() < 0.5 and len(b) > 1:
                    del b[rng.randrange(len(b))]
                else:
                    b.insert(rng.randrange(len(b) + 1), rng.choice(alpha))
            b = "".join(b)
        else:
            b = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 70)))
        gop, gep, glob = rng.choice([0, 1, 2, 5, 10, 15, 40]), rng.choice([0, 1, 3, 10]), rng.random() < 0.5
        try:
            port = P.align(a, b, gop, gep, glob, model)
        except RuntimeError:
            port = None
        if port is not None and len(port[0]) == len(a) + len(b):
            continue   # reference UB, see module docstring
        fuzz.append(record(R, "fuzz", a, b, gop, gep, glob, model))
    doc = {"_generated_by": "tests/golden/make_golden_gotoh2.py from oracle/_ref/_gotoh2*.so "
                            "(= /root/reference/micall/alignment/src/_gotoh2.c, unmodified)",
           "kats": kats, "fuzz": fuzz}
    out = os.path.join(HERE, "gotoh2.json")
    json.dump(doc, open(out, "w"), indent=0)
    print(out, os.path.getsize(out), "kats", len(kats), "fuzz", len(fuzz))


if __name__ == "__main__":
    main()

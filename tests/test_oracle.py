"""The oracle (CPU restatement, test infrastructure) against the vectors generated from the
reference itself (tests/golden/make_golden.py) and - where it is built - against the compiled
reference on a fresh seeded fuzz."""
import random

import numpy as np
import pytest

from conftest import load_golden, matches, resolve


@pytest.mark.parametrize("name", ["appendix_b", "fuzz_small", "shapes"])
def test_port_matches_golden(oracle_port, name):
    cases = load_golden(name)["cases"]
    bad = []
    for c in cases:
        got = oracle_port.align(c["mode"], resolve(c["a"]), resolve(c["b"]), c["gip"], c["gep"], c["term"])
        if not matches(c, got):
            bad.append(c)
    assert not bad, "%d of %d golden cases differ, first: %r" % (len(bad), len(cases), bad[0])


def test_port_tables_match_golden(oracle_port):
    tabs = load_golden("pairscore_tables")["tables"]
    for m in range(3):
        assert (oracle_port.table(m) == np.array(tabs[str(m)])).all()


def test_port_matches_compiled_reference_fuzz(oracle_port, oracle_ref):
    rng = random.Random(99)
    alphas = {0: "ACGTNRYKMSWBDHVacgtnXx*.-Uu$", 1: "ARNDCQEGHILKMFPSTWYVBZ?*XJ_-akl", 2: "ARNDCQEGHILKMFPSTWYVXZ-z"}
    for _ in range(4000):
        mode = rng.choice([0, 0, 1, 2])
        al = alphas[mode]
        a = "".join(rng.choice(al) for _ in range(rng.randint(1, 50)))
        b = "".join(rng.choice(al) for _ in range(rng.randint(1, 50)))
        if mode == 2 and (not a.replace("-", "") or not b.replace("-", "")):
            continue
        gip, gep, term = rng.choice([0, 1, 5, 10, 40]), rng.choice([0, 1, 3, 10]), rng.choice([0, 1])
        assert oracle_port.align(mode, a, b, gip, gep, term) == oracle_ref.align(mode, a, b, gip, gep, term)


def test_port_tables_match_compiled_reference(oracle_port, oracle_ref):
    for m in range(3):
        assert (oracle_port.table(m) == oracle_ref.table(m)).all()


def test_port_rejects_undefined_domain(oracle_port):
    with pytest.raises(ValueError):
        oracle_port.align(0, "   ", "ACGT", 10, 3, 1)      # empty after trim: UB in the reference
    with pytest.raises(ValueError):
        oracle_port.align(0, "AC\x7fT", "ACGT", 10, 3, 1)  # byte 127: out-of-bounds table index


def test_port_batch_equals_single(oracle_port):
    from gotoh_b200 import packing, workloads
    ref, reads = workloads.c2_reads(6, seed=3)
    rb, ro = packing.pack([ref])
    qb, qo = packing.pack(reads)
    oa, ob, off, ln, sc = oracle_port.align_batch(0, rb, ro, np.zeros(6, np.int32), qb, qo, 10, 3, 1)
    for k, q in enumerate(reads):
        a, b, s = oracle_port.align_it(ref, q, 10, 3, 1)
        assert packing.unpack(oa, off, ln)[k] == a and packing.unpack(ob, off, ln)[k] == b and sc[k] == s


# ---- the live aligner gotoh2.Aligner.align ---------------------------------------

def _g2_matches(case, got):
    if "error" in case:
        return got == "traceback"
    if got == "traceback" or got[2] != case["score"] or len(got[0]) != case["len"]:
        return False
    if "out_a" in case:
        return got[0] == case["out_a"] and got[1] == case["out_b"]
    import hashlib
    return (hashlib.sha256(got[0].encode()).hexdigest() == case["sha_a"] and
            hashlib.sha256(got[1].encode()).hexdigest() == case["sha_b"])


def test_gotoh2_port_matches_golden_from_reference_unit_tests_and_fuzz():
    from oracle.oracle2 import Oracle2
    P = Oracle2("port")
    doc = load_golden("gotoh2")
    bad = []
    for c in doc["kats"] + doc["fuzz"]:
        try:
            got = P.align(c["a"], c["b"], c["gop"], c["gep"], c["is_global"], c["model"])
        except RuntimeError:
            got = "traceback"
        if not _g2_matches(c, got):
            bad.append(c["name"])
    assert not bad, bad[:5]
    # the answers micall/alignment/tests/test.py asserts
    k = {c["name"].split(" ")[0]: c for c in doc["kats"]}
    assert (k["TestAlignerSimpleGlobal"]["out_a"], k["TestAlignerSimpleGlobal"]["out_b"], k["TestAlignerSimpleGlobal"]["score"]) == ("ACGT", "AC-T", 9)
    assert (k["TestAlignerSimpleLocal"]["out_b"], k["TestAlignerSimpleLocal"]["score"]) == ("-ACGT-", 20)
    assert k["TestFlouri.test_NWalign_example"]["score"] == -3 and k["TestFlouri.test_Biopp_example1"]["score"] == -15
    assert k["TestIssues.test_issue14"]["out_b"] == "-CA" and k["TestIssues.test_issue14"]["score"] == -1


def test_gotoh2_port_matches_compiled_reference_fuzz():
    from oracle.oracle2 import Oracle2, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref/_gotoh2*.so not built (needs /root/reference)")
    P, R = Oracle2("port"), Oracle2("reference")
    rng = random.Random(77)
    for _ in range(3000):
        model = rng.choice(["HYPHY_NUC", "NWALIGN", "Biopp", "EmpHIV25"])
        alpha = "ACGTN" if model != "EmpHIV25" else "ARNDCQEGHILKMFPSTWYVBZX*?"
        a = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 45)))
        b = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 45)))
        args = (a, b, rng.choice([0, 1, 5, 10, 40]), rng.choice([0, 1, 3, 10]), rng.random() < 0.5, model)
        p = P.align(*args)
        if len(p[0]) == len(a) + len(b):
            continue   # the reference overruns its output buffer here (_gotoh2.c:481,424): UB
        assert p == R.align(*args), args

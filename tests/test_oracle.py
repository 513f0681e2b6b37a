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

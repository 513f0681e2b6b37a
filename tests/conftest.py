import hashlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "micall-lite_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests", "simt_emu"), os.path.join(PKG, "csrc")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def _refs():
    from gotoh_b200 import workloads
    return workloads.references()


def resolve(x):
    """Golden inputs are literal strings or {'ref': name} / {'aa': name} pointers into references.json."""
    if isinstance(x, dict):
        return _refs()["nucleotide"][x["ref"]] if "ref" in x else _refs()["amino"][x["aa"]]
    return x


def sha(s):
    return hashlib.sha256(s.encode("latin-1")).hexdigest()


def matches(case, got):
    """got = (aligned_ref, aligned_qry, score) vs a golden record (literal or hashed outputs)."""
    oa, ob, sc = got
    if sc != case["score"] or len(oa) != case["len"] or len(ob) != case["len"]:
        return False
    if "out_a" in case:
        return oa == case["out_a"] and ob == case["out_b"]
    return sha(oa) == case["sha_a"] and sha(ob) == case["sha_b"]


def run_cases(aligner, cases, max_cells=None, skip_dollar=False):
    """Align golden cases through an Aligner, batched per parameter set.  Returns mismatching cases."""
    groups = {}
    for c in cases:
        a, b = resolve(c["a"]), resolve(c["b"])
        if max_cells is not None and len(a) * len(b) > max_cells:
            continue
        if skip_dollar and "$$$" in a:
            continue
        groups.setdefault((c["mode"], c["gip"], c["gep"], c["term"]), []).append((a, b, c))
    bad, n = [], 0
    for (mode, gip, gep, term), lst in groups.items():
        out = aligner.align_batch([x[0] for x in lst], [x[1] for x in lst], gip, gep, term, mode)
        for (a, b, c), got in zip(lst, out):
            n += 1
            if not matches(c, got):
                bad.append((c, got))
    return n, bad


@pytest.fixture(scope="session")
def oracle_port():
    from oracle.oracle import Oracle
    return Oracle("port")


@pytest.fixture(scope="session")
def oracle_ref():
    from oracle.oracle import Oracle, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref/libgotoh_ref.so not built (needs /root/reference)")
    return Oracle("reference")


@pytest.fixture(scope="session")
def product_library():
    """The product .so (nvcc-built).  Loading it needs no GPU; computing does."""
    import build as build_cuda
    from gotoh_b200 import _ffi
    return _ffi.Library(build_cuda.build())


@pytest.fixture(scope="session")
def emu_aligner():
    """TEST-ONLY: the kernel sources compiled for the CPU SIMT emulator (tests/simt_emu)."""
    import build_emu
    from gotoh_b200 import _ffi
    from gotoh_b200.api import Aligner
    return Aligner(_ffi.Library(build_emu.build()))


@pytest.fixture(scope="session")
def gpu_aligner(product_library):
    from gotoh_b200.api import Aligner
    if product_library.device_count() < 1:
        pytest.fail("gpu test selected but no CUDA device is visible (the product has no CPU path)")
    return Aligner(product_library)

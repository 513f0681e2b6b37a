"""Host-side logic and the C-ABI surface; nothing here needs a GPU and nothing computes an
alignment with the product library (it cannot: it has no CPU path)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


def test_library_exports_every_declared_symbol(product_library):
    from gotoh_b200 import _ffi
    header = open(os.path.join(ROOT, "include", "gotoh_b200.h")).read()
    declared = set(re.findall(r"\b(gotoh_b200_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_ffi.SYMBOLS), declared ^ set(_ffi.SYMBOLS)
    for sym in declared:
        assert hasattr(product_library.lib, sym), sym
    assert product_library.version() == 200


def test_product_tables_match_reference_dump(product_library):
    """K0 (score tables) against pairscore() dumped from the compiled reference (rows a1-a4)."""
    tabs = load_golden("pairscore_tables")["tables"]
    for m in range(3):
        t = np.zeros(127 * 127, dtype=np.int32)
        assert product_library.lib.gotoh_b200_pairscore_table(m, t.ctypes.data) == 0
        assert (t.reshape(127, 127) == np.array(tabs[str(m)])).all()


def test_no_gpu_means_loud_failure_not_fallback(product_library):
    from gotoh_b200 import _ffi
    from gotoh_b200.api import Aligner
    if product_library.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(_ffi.GotohError) as e:
        Aligner(product_library).align_batch(["ACGT"], ["ACT"], 5, 1, 1, 0)
    assert e.value.code == _ffi.ENODEVICE


def test_missing_library_is_an_import_error(tmp_path):
    from gotoh_b200 import _ffi
    with pytest.raises(ImportError):
        _ffi.Library(str(tmp_path / "libgotoh_b200.so"))


def test_product_package_never_imports_oracle():
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import gotoh_b200, gotoh; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)" %
            os.path.join(ROOT, "micall-lite_b200"))
    subprocess.run([sys.executable, "-c", code], check=True)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "micall-lite_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "gotoh_oracle" not in text, f


def test_packing_roundtrip_and_arg_rules():
    from gotoh_b200 import packing
    data, off = packing.pack(["ACGT", b"AC", ""])
    assert data.tobytes() == b"ACGTAC" and off.tolist() == [0, 4, 6, 6]
    with pytest.raises(TypeError):
        packing.pack([None])
    with pytest.raises(ValueError):
        packing.pack(["AC\0GT"])
    ro = np.array([0, 10], dtype=np.int64)
    qo = np.array([0, 3, 7], dtype=np.int64)
    assert packing.out_offsets(ro, np.array([0, 0]), qo).tolist() == [0, 13, 27]


def test_drop_in_signatures_are_positional_only(emu_aligner):
    """The reference parses "ssiii"/"ssii" positionally (gotoh.cpp:633,669,703)."""
    with pytest.raises(TypeError):
        emu_aligner.align_it("ACGT", "ACT", 5, 1)                       # arity
    with pytest.raises(TypeError):
        emu_aligner.align_it(standard="ACGT", seq="ACT", gap_init_penalty=5, gap_extend_penalty=1,
                             use_terminal_gap_penalty=1)                # keywords
    with pytest.raises(TypeError):
        emu_aligner.align_it("ACGT", 7, 5, 1, 1)                        # "s" wants a string
    with pytest.raises(TypeError):
        emu_aligner.align_it("ACGT", "ACT", 5.0, 1, 1)                  # "i" rejects float
    assert emu_aligner.align_it("ACGT", "ACT", 5, 1, True) == ("ACGT", "AC-T", 9)   # bool ok (reference_distances.py:33)


def test_workload_generators_are_seeded_and_shaped():
    from gotoh_b200 import workloads
    ref, qb, qo = workloads.c2_reads_packed(500, seed=11)
    ref2, qb2, qo2 = workloads.c2_reads_packed(500, seed=11)
    assert len(ref) == 3039 and (qb == qb2).all() and (qo == qo2).all()
    lens = np.diff(qo)
    assert lens.min() >= 245 and lens.max() <= 257 and set(qb.tolist()) <= set(b"ACGTN")
    refs, ridx, qb, qo = workloads.c3_queries_packed(300, seed=5)
    assert [len(r) for r in refs] == [99, 440, 288] and set(np.diff(qo).tolist()) <= {83, 84, 85}
    seeds, ridx, qb, qo = workloads.c4_pairs_packed(3, seed=5)
    assert len(seeds) == 57 and all(9000 < n < 10100 for n in np.diff(qo))
    assert workloads.shard_range(10, 0, 3) == (0, 3) and workloads.shard_range(10, 2, 3) == (6, 10)

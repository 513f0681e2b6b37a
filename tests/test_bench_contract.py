"""bench.py's JSON contract, exercised without a GPU: `--emu` runs the script's whole control flow (resident arm, both
end-to-end arms with their byte-for-byte verification, host-ceiling probe, roofline and CPU-baseline bookkeeping) on the
kernel sources under the SIMT emulator.  The numbers mean nothing here; the keys, the verification counts and the
reference arm's line do."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("config,pairs", [("c2", 24), ("c3", 300), ("c5", 40)])
def test_bench_line_has_the_contract_keys_and_verifies_what_it_times(config, pairs):
    d = _run("--emu", "--config", config, "--pairs", str(pairs), "--steps", "1", "--warmup", "1", "--verify", "4", "--cpu-cells-per-core", "2e6")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "bit_exact_verified_pairs"):
        assert key in d, key
    assert d["metric"] == "GCUPS" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["name"] == config and "workload" in d["config"]
    assert d["scaling"] == ("strong" if config == "c5" else "weak")
    e = d["e2e"]
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "compact", "verified_pairs_vs_resident"):
        assert key in e, key
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > e["compact"]["d2h_bytes_per_step"] > 0
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "instr_per_cell"):
        assert key in d["roofline"], key
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in d["cpu_baseline"], key
    v = d["bit_exact_verified_pairs"]
    assert v["resident_vs_oracle"] >= 3 and v["e2e_strings_vs_resident"] > 0 and v["e2e_compact_vs_strings"] == v["e2e_strings_vs_resident"]
    assert d["gpu_launches"] > 0 and "emu_selftest" in d


def test_reference_arm_line():
    d = _run("--impl", "reference", "--config", "c3", "--steps", "1", "--warmup", "0", "--cpu-cells-per-core", "3e6")
    assert d["impl"] == "reference" and d["metric"] == "GCUPS" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0

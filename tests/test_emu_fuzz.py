"""One short round of each randomised differential tool (tools/fuzz_emu.py, tools/stress_emu_pipeline.py) per test run, on
the CPU SIMT emulator, with a fixed seed so the suite stays deterministic - the long campaigns over many seeds are run by
hand."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = "2026"


def _run(tool, trials, **env):
    e = dict(os.environ, **env)
    for k in [k for k in e if k.startswith("GOTOH_B200_")]:
        del e[k]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), SEED, str(trials)], env=e, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0 and " 0 bad" in r.stdout, "%s seed %s %s:\n%s\n%s" % (tool, SEED, env, r.stdout[-3000:], r.stderr[-3000:])


@pytest.mark.parametrize("mode", ["", "FUZZ_STRIPS"])
def test_emu_fuzz_round(mode):
    _run("fuzz_emu.py", 1, **({mode: "1"} if mode else {}))


def test_emu_gotoh2_fuzz_round():
    _run("fuzz_emu_gotoh2.py", 1)


def test_emu_pipeline_stress_round():
    _run("stress_emu_pipeline.py", 2)

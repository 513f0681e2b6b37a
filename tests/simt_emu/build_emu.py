"""Build the TEST-ONLY CPU emulation of libgotoh_b200 (see simt_emu.h).

Output: tests/_build/libgotoh_b200_emu.so.  It exists so that the exact kernel sources can be
checked against the oracle in the GPU-less container; it is never loaded by the product
package and is not a fallback: gotoh_b200/_ffi.py only opens micall-lite_b200/lib/libgotoh_b200.so.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(ROOT, "tests", "_build", "libgotoh_b200_emu.so")
SRC = os.path.join(ROOT, "micall-lite_b200", "csrc")


def build(force=False):
    deps = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(HERE, f) for f in os.listdir(HERE)]
    deps.append(os.path.join(ROOT, "include", "gotoh_b200.h"))
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-I", HERE,
           "-include", os.path.join(HERE, "simt_emu.h"),
           "-x", "c++", os.path.join(SRC, "gotoh_b200.cu"), os.path.join(HERE, "simt_emu.cpp"),
           "-o", OUT, "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emu build failed:\n" + r.stderr[-4000:])
    return OUT


if __name__ == "__main__":
    print(build(force=True))

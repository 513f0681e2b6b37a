"""Compile libgotoh_b200.so for sm_100a with nvcc (cross-compiles without a GPU).

    python micall-lite_b200/csrc/build.py [--force]

Output: micall-lite_b200/lib/libgotoh_b200.so (git-ignored, travels to the GPU box).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OUT = os.path.join(PKG, "lib", "libgotoh_b200.so")
SOURCES = ["gotoh_b200.cu"]
DEPS = ["gotoh_b200.cu", "gotoh_kernels.cuh", "gotoh_tables.h", "gotoh_intpeak.cuh", "gotoh2_kernels.cuh", "gotoh2_host.cuh", "gotoh2_fast.cuh", "gotoh_prep.cuh", "gotoh_plan_math.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; libgotoh_b200 cannot be built (there is no CPU build of it)")
    return p


def build(force=False, verbose=False):
    deps = [os.path.join(HERE, d) for d in DEPS] + [os.path.join(ROOT, "include", "gotoh_b200.h"), __file__]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + [os.path.join(HERE, s) for s in SOURCES] + ["-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(PKG, "lib", "ptxas.log")      # register / spill report of this build (git-ignored, like the .so)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + (r.stdout + r.stderr)[-6000:])
    if verbose:
        print(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

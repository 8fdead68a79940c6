"""Build tests/hostsim/libhostsim.so: the device routine compiled for the host. TEST INFRASTRUCTURE ONLY."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "libhostsim.so")
SRC = os.path.join(HERE, "hostsim.cpp")
CSRC = os.path.join(ROOT, "raymarchcl_b200", "csrc")
DEPS = [SRC, os.path.join(HERE, "stubs", "cuda_runtime.h")] + [
    os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]


def build(verbose: bool = False) -> str:
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    # -ffp-contract=off == nvcc -fmad=false; no fast-math: IEEE division and sqrt like -prec-div/-prec-sqrt
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", "-w",
           "-x", "c++", "-I", os.path.join(HERE, "stubs"), "-I", CSRC, SRC, "-o", OUT]
    if verbose:
        print("[build_hostsim]", " ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(verbose=True)

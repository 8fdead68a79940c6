#!/usr/bin/env python3
"""Compile the C restatements oracle/rm_oracle.c + oracle/meshvoxel_oracle.c -> oracle/librm_oracle.so -- TEST INFRASTRUCTURE ONLY.

Strict IEEE fp32, no FMA contraction (the pinned semantics, see the header of rm_oracle.c).
The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rm_oracle.c")
SRC2 = os.path.join(HERE, "meshvoxel_oracle.c")  # restatement of meshvoxel.clj (the "next" row 8f-4)
OUT = os.path.join(HERE, "librm_oracle.so")


def build(verbose: bool = True, force: bool = False) -> str:
    if (not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(SRC), os.path.getmtime(SRC2))):
        return OUT
    cmd = ["gcc", "-std=c99", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
           "-Wall", "-Wextra", SRC, SRC2, "-o", OUT, "-lm"]
    if verbose:
        print("[build_oracle]", " ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)

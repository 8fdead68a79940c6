#!/usr/bin/env python3
"""Build libraymarch_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

-fmad=false : the geometry code must evaluate a*b+c as two roundings (DESIGN.md "Numerics").
-lineinfo   : ncu source pages map to these files.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libraymarch_b200.so")
SOURCES = ["rm_api.cu", "rm_kernels.cu", "rm_accel.cu", "rm_render_fast.cu", "rm_render_persist.cu", "rm_render_warp.cu", "rm_render_wave.cu", "rm_generate.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared"]


def _newest_source() -> float:
    t = 0.0
    for root in (HERE, os.path.join(PKG, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def build(verbose: bool = True, force: bool = False, extra=(), out: str = OUT) -> str:
    """Compile the library. `extra` nvcc flags and `out` exist for A/B experiments (load the
    variant by pointing RAYMARCH_B200_LIB at it)."""
    OUT = out
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest_source():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *extra, *[os.path.join(HERE, s) for s in SOURCES], "-o", OUT]
    if verbose:
        print("[build]", " ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    argv = sys.argv[1:]
    out = OUT
    if "-o" in argv:
        i = argv.index("-o")
        out = os.path.abspath(argv[i + 1])
        del argv[i:i + 2]
    build(force=True, extra=[a for a in argv if a != "--force"], out=out)
